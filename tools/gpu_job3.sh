#!/bin/bash
# round-2 GPU job 3: A/B of the walk loop shapes and of the sigma-by-multiplication SHA variants
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TPS=4096 bash tools/ab.sh > gpurun_out/j3_ab_scan.log 2>&1; cat gpurun_out/j3_ab_scan.log
for v in peel sigw1 sigw2; do echo "=== hashpeak $v"; KH_B200_LIB=$PWD/gpurun_variants/libkh_$v.so python tools/hashpeak.py; done > gpurun_out/j3_hashpeak.log 2>&1; cat gpurun_out/j3_hashpeak.log
python -c "
import sys; sys.path.insert(0,'.')
import keyhunt_b200 as K, json
kh=K.KeyHunt(0); print(json.dumps({'int':kh.int_peak(),'pipe':kh.pipe_peak()}))" > gpurun_out/j3_peaks.json 2>&1; cat gpurun_out/j3_peaks.json
for v in inloop sigw1 sigw2; do
  echo "=== parity subset with $v"
  KH_B200_LIB=$PWD/gpurun_variants/libkh_$v.so python -m pytest tests/test_gpu_scan.py tests/test_gpu_golden.py tests/test_gpu_configs.py tests/test_gpu_field.py tests/test_gpu_bsgs.py -q -x 2>&1 | tail -2
done > gpurun_out/j3_parity.log 2>&1; cat gpurun_out/j3_parity.log
mkdir -p gpurun_variants/keep && mv gpurun_variants/libkh_noe0.so gpurun_variants/libkh_sigw1.so gpurun_variants/libkh_sigw2.so gpurun_variants/keep/
TPS=4096 bash tools/ab_c4.sh > gpurun_out/j3_ab_c4.log 2>&1; cat gpurun_out/j3_ab_c4.log
