#!/bin/bash
# round-2 job 16: per-kernel choice of the final-reduction form: field KATs of both forms + whole suite, A/B, sanitizer pass
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6 | tee gpurun_out/j16_pytest.log
TPS=4096 bash tools/ab.sh 2>&1 | tee gpurun_out/j16_ab_scan.log
TPS=4096 bash tools/ab.sh 2>&1 | tee -a gpurun_out/j16_ab_scan.log
