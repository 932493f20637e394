#!/bin/bash
# round-2 job 18 (= job 13 on the last build): the final build: whole GPU suite, full bench line, ncu --set full of every walk kernel (summarised on the box: the
# reports themselves exceed the 64 MiB that come back), launch list of bench.py
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/j21_pytest.log 2>&1; tail -5 gpurun_out/j21_pytest.log
( time python bench.py ) > gpurun_out/j21_bench.json 2> gpurun_out/j21_bench.err
echo "bench rc=$?"; tail -12 gpurun_out/j21_bench.err
cap() {  # cap <tag> <kernel regex> <skip> <command...>
  tag=$1; rx=$2; skip=$3; shift 3
  "$@" > gpurun_out/j21_${tag}_plain.log 2>&1 || { echo "$tag: plain run failed"; tail -5 gpurun_out/j21_${tag}_plain.log; return; }
  ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/j21_$tag "$@" > gpurun_out/j21_${tag}_ncu.log 2>&1
  echo "$tag ncu rc=$?"
  python tools/ncu_summary.py /tmp/j21_$tag.ncu-rep gpurun_out/j21_${tag}_ncu_sections.txt > /dev/null
  python tools/ncu_opmix.py /tmp/j21_$tag.ncu-rep gpurun_out/j21_${tag}_opmix.txt > /dev/null
  ncu -i /tmp/j21_$tag.ncu-rep --page raw --csv > gpurun_out/j21_${tag}_raw.csv 2>/dev/null
  rm -f /tmp/j21_$tag.ncu-rep
}
cap both   kh_scan_kernel 2 python tools/prof_kernel.py both 27
cap uncomp kh_scan_kernel 2 python tools/prof_kernel.py uncomp 27
cap comp   kh_scan_kernel 2 python tools/prof_kernel.py comp 27
cap eth    kh_scan_kernel 2 python tools/prof_kernel.py eth 27
cap xpoint kh_scan_kernel 2 python tools/prof_kernel.py xpoint 28 1000000
cap giant  kh_giant_kernel 0 python tools/prof_giant.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/j21_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/j21_bench_ncu.log 2>&1; echo "ncu list rc=$?"
du -sh gpurun_out; ls -la gpurun_out/j21_* | head -50
