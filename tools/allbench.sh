cd "${GRAFT_REPO_ROOT:-.}"
python bench.py 2>/dev/null | tail -1 > gpurun_out/fb_c2.json
for w in c1 c3 c5btc c5eth; do python bench.py --workload $w --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/fb_$w.json; done
python - <<'PY'
import json
for w in ("c2","c1","c3","c5btc","c5eth"):
    d=json.load(open("gpurun_out/fb_%s.json"%w))
    print(w, round(d["value"],1), round(d["e2e"]["value"],1), round(d["roofline"]["frac"],3), d["roofline"].get("binding_pipe") and round(d["roofline"]["binding_pipe"]["frac"],3), round(d["cpu_baseline"]["value"],1), d["cpu_baseline"].get("hits_equal_gpu"), d["hits"])
PY
