#!/bin/bash
# one ncu --set full capture of the scan kernel + a launch list of one bench step
mkdir -p gpurun_out
N=${1:-134217728}
ncu --set full --clock-control none --import-source on -k regex:scan_flat -s 2 -c 1 -f -o gpurun_out/prof_scan \
    python bench.py --n-vectors $N --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --n-vectors $N --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
