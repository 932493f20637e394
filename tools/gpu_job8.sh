#!/bin/bash
# round-2 job 8: binned BSGS build (tests + timing), A/B of the carry-free first rows (KH_PLAIN_HEAD)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bsgs.py tests/test_gpu_field.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/j8_pytest.log
cat gpurun_out/bsgs_build_binned.json
bash tools/ab.sh 2>&1 | tee gpurun_out/j8_ab_scan.log
bash tools/ab_c4.sh 2>&1 | tee gpurun_out/j8_ab_c4.log
