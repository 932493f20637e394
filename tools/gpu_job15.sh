#!/bin/bash
# round-2 job 15: rare-branch final reduction of the multiplier (KH_RARE_REDUCE): field KATs + scans on the device, A/B
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_field.py tests/test_gpu_scan.py tests/test_gpu_golden.py tests/test_gpu_bsgs.py -x -q -m gpu ) 2>&1 | tail -6 | tee gpurun_out/j15_pytest.log
TPS=4096 bash tools/ab.sh 2>&1 | tee gpurun_out/j15_ab_scan.log
TPS=4096 bash tools/ab_c4.sh 2>&1 | tee gpurun_out/j15_ab_c4.log
TPS=4096 bash tools/ab.sh 2>&1 | tee -a gpurun_out/j15_ab_scan.log
