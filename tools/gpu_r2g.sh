#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
python tools/bench_legs.py 2 > gpurun_out/legs2.log 2>&1; python - <<PY
import json
for l in open('gpurun_out/legs2.log'):
    if l.startswith('{'):
        v=json.loads(l); print({kk: v[kk] for kk in ('ms','e2e_ms','spot_check_ok','stage_ms','scan_kernel_ms','launches') if kk in v})
    else: print(l[:300])
PY
export QADC_PROFILE_RANGE=1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv
unset QADC_PROFILE_RANGE
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv \
    python tools/bench_legs.py 2 > gpurun_out/ncu_list2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_c2.csv | tail -14
