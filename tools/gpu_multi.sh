#!/bin/bash
# bench.py on N GPUs exactly as the driver launches it (run under gpurun --gpus N). usage: tools/gpu_multi.sh N [bench args]
mkdir -p gpurun_out
N=${1:-2}; shift
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" ) > gpurun_out/bench_n$N.log 2>&1
grep '^{' gpurun_out/bench_n$N.log | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n$N.json'))
    print('N=$N value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'],'share %.3f'%d['roofline']['kernel_share_of_step'], 'e2e %.1f'%(d['e2e']['value']/1e9), d['verify'], d['clocks'])
    for k,v in (d.get('configs') or {}).items():
        print(k, json.dumps(v)[:1500])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_n$N.log').read()[-4000:])
PY
grep real gpurun_out/bench_n$N.log
