#!/bin/bash
# round 2, third session: per-step fixed cost of the flat search on rank 0's shard of an 8-way sharding (base build vs the
# current one), bit-for-bit comparison of their results, the GPU test suite, and a warm-cache launch list of the new build.
mkdir -p gpurun_out
echo "== tests"
timeout ${TEST_TIMEOUT:-540} python -X faulthandler -m pytest tests -m gpu -q -x --timeout 400 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -${TAIL:-6} gpurun_out/pytest_gpu.log
echo "== base"; QADC_LIB=$PWD/build_ab/base.so timeout 300 python tools/bench_fixed.py gpurun_out/fixed_base.npz 2>&1 | tail -4
echo "== new";  timeout 300 python tools/bench_fixed.py gpurun_out/fixed_new.npz 2>&1 | tail -6
python tools/bench_fixed.py gpurun_out/fixed_base.npz gpurun_out/fixed_new.npz
echo "== launch list (warm caches) of the new build"
STEPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"prefix|tables_kernel|quantize|scan_flat|merge_lists|flat_" -c 400 --csv --log-file gpurun_out/r02c_launches_fixed_g8.csv \
    python tools/bench_fixed.py > gpurun_out/ncu_fixed.log 2>&1
python tools/launch_summary.py gpurun_out/r02c_launches_fixed_g8.csv > gpurun_out/r02c_launches_fixed_g8.txt; cat gpurun_out/r02c_launches_fixed_g8.txt
