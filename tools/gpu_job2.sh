#!/bin/bash
# round-2 GPU job 2: box-variance check of the C2 kernel, CLI -n test evidence, ncu --set full captures (one launch each)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,uuid,serial,vbios_version,clocks.max.sm,power.limit --format=csv > gpurun_out/j2_gpu.txt 2>&1
TPS=4096 bash tools/ab.sh > gpurun_out/j2_ab_scan.log 2>&1; cat gpurun_out/j2_ab_scan.log
python -m pytest tests/test_gpu_cli.py -q -k small_n > gpurun_out/j2_cli.log 2>&1; tail -2 gpurun_out/j2_cli.log; cat gpurun_out/cli_small_n.json; echo
python -c "
import sys; sys.path.insert(0,'.')
import keyhunt_b200 as K, json
kh=K.KeyHunt(0); print(json.dumps({'int':kh.int_peak(),'pipe':kh.pipe_peak(),'hash2':kh.hash_peak(2),'hash8':kh.hash_peak(8)}))" > gpurun_out/j2_peaks.json 2>&1; cat gpurun_out/j2_peaks.json
cap() {  # cap <tag> <kernel regex> <skip> <command...>
  tag=$1; rx=$2; skip=$3; shift 3
  "$@" > gpurun_out/j2_${tag}_plain.log 2>&1 || { echo "$tag: plain run failed"; tail -5 gpurun_out/j2_${tag}_plain.log; return; }
  tail -1 gpurun_out/j2_${tag}_plain.log
  ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o gpurun_out/j2_$tag "$@" > gpurun_out/j2_${tag}_ncu.log 2>&1
  echo "$tag ncu rc=$?"; tail -1 gpurun_out/j2_${tag}_ncu.log
}
cap both   kh_scan_kernel 2 python tools/prof_kernel.py both 27
cap eth    kh_scan_kernel 2 python tools/prof_kernel.py eth 27
cap uncomp kh_scan_kernel 2 python tools/prof_kernel.py uncomp 27
cap comp   kh_scan_kernel 2 python tools/prof_kernel.py comp 27
cap xpoint kh_scan_kernel 2 python tools/prof_kernel.py xpoint 28 1000000
cap giant  kh_giant_kernel 0 python tools/prof_giant.py
cap refine kh_refine_kernel 0 python tools/prof_refine.py
cap baby   kh_baby_kernel 1 python tools/prof_baby.py 64
ls -la gpurun_out/j2_*.ncu-rep
