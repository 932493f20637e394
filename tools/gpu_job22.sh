#!/bin/bash
# round-2 job 22: the bench line of the final build with the final roofline constants
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/j26_bench.json 2> gpurun_out/j26_bench.err
echo "bench rc=$?"; tail -8 gpurun_out/j26_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/j26_bench_ref.json 2> gpurun_out/j26_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/j26_bench_ref.json | cut -c1-600
