#!/bin/bash
# round-2 job 10: whole GPU suite on the two-level centre set-up and the sliced binned build; per-call cost of kh_bsgs_search again
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -12 | tee gpurun_out/j10_pytest.log
python tools/c4_calls.py 512 > gpurun_out/j10_c4_calls.json 2>&1; cat gpurun_out/j10_c4_calls.json | tr -d '\n' | cut -c1-3000; echo
python tools/c4_calls.py 4096 > gpurun_out/j10_c4_calls_k4096.json 2>&1; head -6 gpurun_out/j10_c4_calls_k4096.json
for kb in 16384 32768 98304; do echo "slice KB $kb"; KH_BABY_SLICE_KB=$kb python tools/prof_baby.py 4096 | tail -1; done
python tools/perf_probe.py 4096 2>&1 | grep "tp=" | awk '{print $2, $7, $5}'
