#!/bin/bash
# round-2 GPU job 5: the single-shape walk (cold work before/after the batch) vs round 1, and the load flavours of the bitmap probe
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out gpurun_variants/later
mv gpurun_variants/libkh_ld*.so gpurun_variants/later/
TPS=4096 bash tools/ab.sh > gpurun_out/j5_ab_scan.log 2>&1; cat gpurun_out/j5_ab_scan.log
mv gpurun_variants/later/libkh_ld*.so gpurun_variants/
TPS=4096 bash tools/ab_c4.sh > gpurun_out/j5_ab_c4.log 2>&1; cat gpurun_out/j5_ab_c4.log
for v in new ld1 ld2 ld3 ld4 ld5 ld6; do
  export KH_B200_LIB=$PWD/gpurun_variants/libkh_$v.so
  python tools/prof_giant.py > gpurun_out/j5_giant_${v}_plain.log 2>&1 && \
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum \
      --clock-control none -k regex:kh_giant_kernel -c 1 --csv --log-file gpurun_out/j5_giant_${v}_metrics.csv python tools/prof_giant.py > /dev/null 2>&1
  echo "== $v: $(tail -1 gpurun_out/j5_giant_${v}_plain.log)"; grep -E "dram__bytes_read|lts__t_sectors_srcunit|lts__t_requests|duration" gpurun_out/j5_giant_${v}_metrics.csv | awk -F'","' '{print "   " $(NF-2), $(NF-1), $NF}'
done
unset KH_B200_LIB
python -m pytest tests/test_gpu_scan.py tests/test_gpu_configs.py tests/test_gpu_bsgs.py -q -x 2>&1 | tail -2
