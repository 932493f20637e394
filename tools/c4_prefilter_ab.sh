cd "${GRAFT_REPO_ROOT:-.}"
make -C oracle oracle >/dev/null
timeout 500 python -m pytest tests/test_gpu_bsgs.py tests/test_gpu_bsgsd.py -x -q -m gpu 2>&1 | tail -3
for pf in 1 0; do
  echo "== bsgs_prefilter=$pf"
  KH_BSGS_PREFILTER=$pf KH_TPS=4096 timeout 300 python tools/c4.py 512 | python -c "import json,sys; d=json.load(sys.stdin); print({k:d.get(k) for k in ['giant_steps_per_s','sweep_walk_ms','build_walk_ms','last_window_found','sweep_tier1_pos','planted_found','members_tier1','sorted_ok']})"
done
