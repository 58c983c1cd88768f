#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "warp_ring or flat_scan" > gpurun_out/pytest_wr.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_wr.log
BENCH_ARGS="--no-configs --flat-ring 0" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 1" STEPS=10 bash tools/gpu_ab.sh cur wr_n16s6 wr_n16s3
BENCH_ARGS="--no-configs --flat-ring 0" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 1" STEPS=10 NCU=cur bash tools/gpu_ab.sh cur
