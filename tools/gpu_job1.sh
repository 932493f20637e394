#!/bin/bash
# round-2 GPU job 1: full GPU test suite, A/B of the library variants, the full bench line, the reference arm, ncu metric list
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/j1_gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/j1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
tail -5 gpurun_out/j1_pytest.log
TPS=4096 bash tools/ab.sh > gpurun_out/j1_ab_scan.log 2>&1
cat gpurun_out/j1_ab_scan.log
TPS=4096 bash tools/ab_c4.sh > gpurun_out/j1_ab_c4.log 2>&1
cat gpurun_out/j1_ab_c4.log
( time python bench.py ) > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/j1_bench.err; head -c 600 gpurun_out/j1_bench.json
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/j1_ref.json 2> gpurun_out/j1_ref.err
tail -4 gpurun_out/j1_ref.err
timeout 120 ncu --query-metrics > gpurun_out/j1_ncu_metrics.txt 2>&1
wc -l gpurun_out/j1_ncu_metrics.txt
