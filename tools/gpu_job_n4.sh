#!/bin/bash
# round-2 job (run with gpurun --gpus 4): the N = 4 paths on hardware: weak-scaled headline + strong C5 / C4 blocks
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n4_gpus.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 6 --warmup 3 ) > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/n4_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/n4_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['hits'], d['n_gpus'])
s=d['strong']
for c in ('c5btc','c5eth'): print(c, s['c5'][c]['time_s'], s['c5'][c]['efficiency'], s['c5'][c]['planted_found_and_nothing_else'])
print(s['c4']['sweep']['time_s'], s['c4']['sweep']['efficiency'], s['c4']['build_s_every_gpu'], s['c4']['planted']['planted_found'])
PY
