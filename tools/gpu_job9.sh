#!/bin/bash
# round-2 job 9: carry-form microbenchmarks, per-call cost of kh_bsgs_search, k=4096 build time, ncu of the binned table build, launch list of bench.py
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
import keyhunt_b200 as K, json
kh=K.KeyHunt(0); print(json.dumps({'int':kh.int_peak(),'pipe':kh.pipe_peak()}))" > gpurun_out/j9_peaks.json 2>&1; cat gpurun_out/j9_peaks.json
python tools/c4_calls.py 512 > gpurun_out/j9_c4_calls.json 2>&1; cat gpurun_out/j9_c4_calls.json
python tools/c4_calls.py 4096 > gpurun_out/j9_c4_calls_k4096.json 2>&1; head -8 gpurun_out/j9_c4_calls_k4096.json
cap() {  # cap <tag> <kernel regex> <skip> <command...>
  tag=$1; rx=$2; skip=$3; shift 3
  ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o gpurun_out/j9_$tag "$@" > gpurun_out/j9_${tag}_ncu.log 2>&1
  echo "$tag ncu rc=$?"; tail -1 gpurun_out/j9_${tag}_ncu.log
}
python tools/prof_baby.py 512 > gpurun_out/j9_baby_plain.log 2>&1; tail -1 gpurun_out/j9_baby_plain.log
cap baby  kh_baby_kernel 1 python tools/prof_baby.py 512
cap apply kh_baby_apply 1 python tools/prof_baby.py 512
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/j9_baby_launches.csv python tools/prof_baby.py 512 > gpurun_out/j9_baby_launches.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/j9_bench_short.json 2> gpurun_out/j9_bench_short.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j9_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/j9_bench_ncu.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/j9_*
