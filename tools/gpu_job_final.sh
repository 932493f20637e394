#!/bin/bash
# what the driver runs at round end, on the final tree: the whole GPU suite and smoke()
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6 | tee gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
python tools/sanitize_small.py 2>&1 | tail -4
