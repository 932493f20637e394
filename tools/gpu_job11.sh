#!/bin/bash
# round-2 job 11: SHA-256 schedule table of the uncompressed key's second block: parity on the device, A/B
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_scan.py tests/test_gpu_golden.py tests/test_gpu_endo.py tests/test_gpu_vanity.py tests/test_gpu_configs.py -x -q -m gpu ) 2>&1 | tail -8 | tee gpurun_out/j11_pytest.log
TPS=4096 bash tools/ab.sh 2>&1 | tee gpurun_out/j11_ab_scan.log
TPS=4096 bash tools/ab.sh 2>&1 | tee -a gpurun_out/j11_ab_scan.log
