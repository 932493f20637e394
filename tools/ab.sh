#!/bin/bash
# A/B the library variants in gpurun_variants/ with the throughput probe; TPS = threads per SM list
cd "${GRAFT_REPO_ROOT:-.}"
for v in gpurun_variants/libkh_*.so; do
  for tp in ${TPS:-512}; do
    echo "=== $v tp=$tp"
    KH_B200_LIB=$PWD/$v python tools/perf_probe.py $tp 2>&1 | grep "tp=" | awk '{print $2, $7}'
  done
done
