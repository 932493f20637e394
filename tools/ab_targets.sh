#!/bin/bash
# A/B the library variants in gpurun_variants/ against target-set size (NTGS list): bloom in L1 / L2 / HBM
cd "${GRAFT_REPO_ROOT:-.}"
for n in ${NTGS:-1024 1000000 50000000}; do
  for v in gpurun_variants/libkh_*.so; do
    echo "=== $v targets=$n"
    KH_NTG=$n KH_B200_LIB=$PWD/$v python tools/perf_probe.py ${TPS:-4096} 2>&1 | grep "tp=" | awk '{print $2, $7}' | tr '\n' ' '; echo
  done
done
