#!/bin/bash
# round-2 job 25: ncu sections of the table-build and set-up kernels on the final build (summarised on the box)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
cap() {  # cap <tag> <kernel regex> <skip> <command...>
  tag=$1; rx=$2; skip=$3; shift 3
  ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/j25_$tag "$@" > gpurun_out/j25_${tag}_ncu.log 2>&1
  echo "$tag ncu rc=$?"
  python tools/ncu_summary.py /tmp/j25_$tag.ncu-rep gpurun_out/j25_${tag}_ncu_sections.txt > /dev/null
  python tools/ncu_opmix.py /tmp/j25_$tag.ncu-rep gpurun_out/j25_${tag}_opmix.txt > /dev/null
  rm -f /tmp/j25_$tag.ncu-rep
}
python tools/prof_baby.py 512 | tail -1
cap baby  kh_baby_kernel 1 python tools/prof_baby.py 512
cap apply kh_baby_apply 1 python tools/prof_baby.py 512
cap setupfill kh_setup_fill_kernel 0 python tools/prof_kernel.py comp 27
cap setup 'kh_setup_kernel' 0 python tools/prof_kernel.py comp 27
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/j25_baby_launches.csv python tools/prof_baby.py 512 > /dev/null 2>&1
grep -E "gpu__time_duration.sum|Kernel Name" gpurun_out/j25_*_ncu_sections.txt
