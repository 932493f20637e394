#!/bin/bash
# round-2 GPU job 14 (as job 7, final build) (run with gpurun --gpus 2): the N > 1 paths on hardware
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/j23_gpus.txt
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_scan.py -q -m gpu ) > gpurun_out/j23_multi.log 2>&1; tail -4 gpurun_out/j23_multi.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 ) > gpurun_out/j23_bench_n2.json 2> gpurun_out/j23_bench_n2.err
echo "bench rc=$?"; tail -5 gpurun_out/j23_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/j23_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['hits'])
print(json.dumps(d.get('strong'), indent=1)[:3000])
PY
