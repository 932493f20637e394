#!/bin/bash
# round-2 GPU job 6: the planned-scan build (round-1 kernels + host-side plan): full GPU suite, A/B against round 1, full bench
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TPS=4096 bash tools/ab.sh > gpurun_out/j6_ab_scan.log 2>&1; cat gpurun_out/j6_ab_scan.log
TPS=4096 bash tools/ab_c4.sh > gpurun_out/j6_ab_c4.log 2>&1; cat gpurun_out/j6_ab_c4.log
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/j6_pytest.log 2>&1; tail -5 gpurun_out/j6_pytest.log
( time python bench.py ) > gpurun_out/j6_bench.json 2> gpurun_out/j6_bench.err
echo "bench rc=$?"; tail -8 gpurun_out/j6_bench.err; head -c 400 gpurun_out/j6_bench.json
