#!/bin/bash
# round-2 job 24: the variants outside the headline configs on the final build (-m vanity, -e) and the BSGSD.md server example at -k 4096
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
make -C oracle oracle >/dev/null 2>&1
python tools/perf_extra.py 2>&1 | tee gpurun_out/j24_perf_extra.log
bash tools/bsgsd_demo.sh 4096 2>&1 | tail -8 | tee gpurun_out/j24_bsgsd_demo.log
python tools/c4_calls.py 512 > gpurun_out/j24_c4_calls.json 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/j24_c4_calls.json'))
for c in d['calls']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in c.items()})
PY
