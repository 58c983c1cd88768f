#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py "$@" ) > gpurun_out/bench_default.log 2>&1
grep '^{' gpurun_out/bench_default.log | tail -1 > gpurun_out/bench_default.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    print('value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'],'share %.3f'%d['roofline']['kernel_share_of_step'], 'e2e %.1f'%(d['e2e']['value']/1e9), d['verify']['ok'], d['clocks'], 'launches', d['gpu_launches'])
    for k,v in (d.get('configs') or {}).items():
        if isinstance(v, dict):
            print(k, {kk: v[kk] for kk in ('value','queries_per_s','ms','e2e_ms','spot_check_ok','queries_per_pass','stage_ms','scan_kernel_ms') if kk in v}, v.get('roofline',{}).get('frac'))
            for vv in v.get('variants', []): print('    qb', vv['queries_per_pass'], 'ms %.3f'%vv['ms'], 'scan %.3f'%vv['scan_kernel_ms'], 'pairs/s %.1f G'%(vv['pairs_per_s']/1e9), vv['spot_check_ok'])
        else: print(k, v)
    print('cpu', d.get('cpu_baseline')); print('recall', d.get('recall_check'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_default.log').read()[-3000:])
PY
grep real gpurun_out/bench_default.log
