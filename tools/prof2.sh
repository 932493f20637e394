#!/bin/bash
# targeted ncu sections on one short launch of a given scan kind: tools_prof2.sh <kind> <tag>
cd "${GRAFT_REPO_ROOT:-.}"
kind=$1; tag=$2; ntg=${3:-1024}
python tools/prof_kernel.py $kind 27 $ntg > gpurun_out/prof2_${tag}_plain.log 2>&1 || { echo plain failed; cat gpurun_out/prof2_${tag}_plain.log; exit 1; }
cat gpurun_out/prof2_${tag}_plain.log
ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section InstructionStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis \
    --metrics gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:kh_scan_kernel -s 2 -c 1 -f -o gpurun_out/prof2_$tag python tools/prof_kernel.py $kind 27 $ntg > gpurun_out/prof2_${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/prof2_${tag}_ncu.log
