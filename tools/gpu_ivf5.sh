#!/bin/bash
# per-GPU work of BASELINE config 5 (1e9 vectors, IVF-65536, nprobe 128, 8-way sharded lists) on ONE GPU:
# timing + per-stage metrics, then an ncu launch list of the timed batches (shares only).
mkdir -p gpurun_out
N=${1:-1000000000}; W=${2:-8}; Q=${3:-10000}
python tools/bench_ivf_sharded.py --n-vectors $N --as-rank-of $W --queries $Q --steps 3 > gpurun_out/ivf5.log 2>&1
tail -1 gpurun_out/ivf5.log | cut -c1-900
[ -n "$NO_NCU" ] && exit 0
QADC_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv \
    --log-file gpurun_out/launches_ivf5.csv python tools/bench_ivf_sharded.py --n-vectors $N --as-rank-of $W --queries $Q --steps 1 \
    > gpurun_out/ivf5_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_ivf5.csv | tee gpurun_out/launches_ivf5.txt
