#!/bin/bash
# ncu of the fixed-cost kernels of a bench step (prefix scan, merges, prefix histogram) + ADD variants A/B
mkdir -p gpurun_out
export QADC_PROFILE_RANGE=1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"prefix_scan|merge_lists|prefix_hist" -c 4 -f -o gpurun_out/prof_fixed \
    python bench.py --steps 1 --warmup 3 --no-cpu --verify 0 > gpurun_out/ncu_fixed.log 2>&1
tail -2 gpurun_out/ncu_fixed.log | cut -c1-200
unset QADC_PROFILE_RANGE
STEPS=10 bash tools/gpu_ab.sh trim add1 add2 trim add1 add2
