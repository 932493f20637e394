#!/bin/bash
# round-2 job 17: "hi << 32" folded into the odd accumulator of the reduction (KH_FOLD_IN_O): field KATs on the device, A/B
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_field.py tests/test_gpu_scan.py -x -q -m gpu ) 2>&1 | tail -6 | tee gpurun_out/j17_pytest.log
TPS=4096 bash tools/ab.sh 2>&1 | tee gpurun_out/j17_ab_scan.log
TPS=4096 bash tools/ab_c4.sh 2>&1 | tee gpurun_out/j17_ab_c4.log
TPS=4096 bash tools/ab.sh 2>&1 | tee -a gpurun_out/j17_ab_scan.log
