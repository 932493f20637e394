#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
make -C oracle oracle >/dev/null
for v in gpurun_variants/libkh_*.so; do
  echo -n "=== $v "
  KH_B200_LIB=$PWD/$v python tools/c4.py 512 | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ['giant_steps_per_s','sweep_walk_ms','last_window_found','sweep_tier1_pos']})"
done
