#!/bin/bash
# ncu --set full of the fixed-cost kernels of a flat step (rank 0 of an 8-way sharding): one launch each, after warm-up
mkdir -p gpurun_out
STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${REGEX:-prefix_scan|prefix_hist|quantize_kernel|merge_lists}" -s ${SKIP:-16} -c ${COUNT:-4} -f \
    -o gpurun_out/prof_fixed python tools/bench_fixed.py > gpurun_out/ncu_fixed_full.log 2>&1
tail -3 gpurun_out/ncu_fixed_full.log | cut -c1-200
