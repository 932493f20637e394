#!/bin/bash
# round-2 job 19: dedicated squaring in the inversion (KH_INV_SQR), target table packed on the device: field KATs + scans, A/B
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_field.py tests/test_gpu_scan.py tests/test_gpu_golden.py tests/test_gpu_bsgs.py tests/test_gpu_configs.py -x -q -m gpu ) 2>&1 | tail -6 | tee gpurun_out/j19_pytest.log
TPS=4096 bash tools/ab.sh 2>&1 | tee gpurun_out/j19_ab_scan.log
TPS=4096 bash tools/ab_c4.sh 2>&1 | tee gpurun_out/j19_ab_c4.log
TPS=4096 bash tools/ab.sh 2>&1 | tee -a gpurun_out/j19_ab_scan.log
python bench.py --workload c3 --steps 4 --warmup 3 --no-side-workloads --no-strong --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3 value', d['value'], 'e2e', d['e2e']['value'])"
