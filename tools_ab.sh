#!/bin/bash
# A/B the library variants in gpurun_variants/ with the throughput probe
cd "${GRAFT_REPO_ROOT:-.}"
for v in gpurun_variants/libkh_*.so; do
  echo "=== $v"
  KH_B200_LIB=$PWD/$v python tools_perf_probe.py 512 2>&1 | grep "tp=" | awk '{print $2, $7}'
done
