#!/bin/bash
# Run under gpurun: plain bench, then the ncu launch list, then one --set full capture of the top kernel.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
CMD="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline ${BENCH_EXTRA:-}"
$CMD > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -20 gpurun_out/plain.err; exit 1; }
cat gpurun_out/plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:kh_scan_kernel -s 13 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
