#!/bin/bash
# Run under gpurun: plain bench, then the ncu launch list of the same command (one pass, no replay),
# then section captures of the top kernels on short launches (tools/prof2.sh).
set -u
cd "${GRAFT_REPO_ROOT:-.}"
CMD="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline ${BENCH_EXTRA:-}"
$CMD > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -20 gpurun_out/plain.err; exit 1; }
cat gpurun_out/plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in both comp; do bash tools/prof2.sh $k ${k}_final; done
