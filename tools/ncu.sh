#!/bin/bash
# one ncu --set full capture of the scan kernel + a launch list of the timed bench steps
# (never a bench value: numbers printed under ncu are discarded)
mkdir -p gpurun_out
N=${1:-1000000000}
export QADC_PROFILE_RANGE=1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_flat -c 1 -f -o gpurun_out/prof_scan \
    python bench.py --n-vectors $N --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --n-vectors $N --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
