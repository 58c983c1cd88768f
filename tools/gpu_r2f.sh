#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
python tools/bench_legs.py 2 > gpurun_out/legs2.log 2>&1; python - <<PY
import json
for l in open('gpurun_out/legs2.log'):
    if l.startswith('{'):
        v=json.loads(l); print({kk: v[kk] for kk in ('ms','e2e_ms','spot_check_ok','stage_ms','scan_kernel_ms','launches') if kk in v})
    else: print(l[:300])
PY
