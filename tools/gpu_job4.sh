#!/bin/bash
# round-2 GPU job 4: A/B of the in-loop walk shape with the parked cold move, and of the scratch prefetch
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TPS=4096 bash tools/ab.sh > gpurun_out/j4_ab_scan.log 2>&1; cat gpurun_out/j4_ab_scan.log
echo "=== parity subset with inloop"
KH_B200_LIB=$PWD/gpurun_variants/libkh_inloop.so python -m pytest tests/test_gpu_scan.py tests/test_gpu_golden.py tests/test_gpu_configs.py tests/test_gpu_bsgs.py -q -x 2>&1 | tail -2
TPS=4096 bash tools/ab_c4.sh > gpurun_out/j4_ab_c4.log 2>&1; cat gpurun_out/j4_ab_c4.log
