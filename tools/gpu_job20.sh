#!/bin/bash
# round-2 job 20: A/B of the out-of-line squaring in the hash kernels (KH_INV_SQR_HASH) and of Keccak's peeled first/last round (KH_KECCAK_PEEL); parity of the ETH path
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_scan.py tests/test_gpu_golden.py tests/test_gpu_endo.py tests/test_gpu_configs.py -x -q -m gpu ) 2>&1 | tail -6 | tee gpurun_out/j20_pytest.log
TPS=4096 bash tools/ab.sh 2>&1 | tee gpurun_out/j20_ab_scan.log
TPS=4096 bash tools/ab.sh 2>&1 | tee -a gpurun_out/j20_ab_scan.log
