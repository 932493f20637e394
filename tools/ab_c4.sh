#!/bin/bash
# A/B the library variants in gpurun_variants/ on the C4 (bsgs -k 512) sweep; TPS = threads per SM list
cd "${GRAFT_REPO_ROOT:-.}"
make -C oracle oracle >/dev/null
for v in gpurun_variants/libkh_*.so; do
  for tp in ${TPS:-512}; do
    echo -n "=== $v tp=$tp "
    KH_TPS=$tp KH_B200_LIB=$PWD/$v python tools/c4.py 512 | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ['giant_steps_per_s','sweep_walk_ms','build_walk_ms','last_window_found','sweep_tier1_pos']})"
  done
done
