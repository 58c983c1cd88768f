#!/usr/bin/env python
"""Experiment: scan time of the flat bench step as a function of the shared-bound seed (seed - k through the
experiment-only option seed_minus; results are compared with k = 0, a seed below the true r-th distance loses results).
Needs a library built with the experiment hook:  QADC_NVCC_EXTRA=-DQADC_EXPERIMENT python quick-adc_b200/build.py --force
(the product build does not know the option).  Measured on rank 0 of an 8-way sharding of the 1e9 bench (profiles/README.md):
seed - 0 / 5 / 10 / 15: 2.78 / 2.68 / 2.60 / 2.57 ms without sharing the bound across a query's CTAs, 2.64 / 2.63 / 2.61 / 2.61
with it (option flat_share).
usage: [G=8] python tools/exp_seed.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, qadc_b200
from qadc_b200 import sharding
G = int(os.environ.get("G", "8")); STEPS = 20
N, nq, R = 10 ** 9, 16, bench.R
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
cb, queries = bench.make_quantizer_and_queries(nq)
ix = qadc_b200.Index(0, stream.cuda_stream); ix.set_pq(bench.DIM, bench.M, cb)
lo, hi = sharding.flat_shard_range(N, 0, G); n_local = hi - lo
ix.begin_database([n_local], False)
for c0 in range(lo, hi, 1 << 24):
    c1 = min(c0 + (1 << 24), hi); t = bench.codes_torch(c0, c1, dev); torch.cuda.synchronize(dev)
    ix.upload_codes_device(0, c0 - lo, c1 - c0, t.data_ptr()); del t
n_prefix = sharding.start_size(N, bench.KEEP)
if G > 1:
    ix.set_position_base(0, lo); pre = bench.codes_torch(0, n_prefix, dev); torch.cuda.synchronize(dev)
    ix.set_prefix_device(0, pre.data_ptr(), n_prefix); del pre
ix.finalize(bench.KEEP); ix.set_option("flat_qb", 1); ix.set_option("time_scan", 1)
d_q = torch.from_numpy(queries).to(dev)
d_ids = torch.empty((nq, R), dtype=torch.int32, device=dev); d_d = torch.empty((nq, R), dtype=torch.int8, device=dev)
d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
base = None
for share in (0, 1, 0, 1):
    ix.set_option("flat_share", share)
    for k in (0, 5, 10, 15):
        ix.set_option("seed_minus", k)
        for _ in range(3 + STEPS):
            ix.search_device(d_q.data_ptr(), nq, 1, R, d_ids.data_ptr(), d_d.data_ptr(), d_cnt.data_ptr())
        torch.cuda.synchronize(dev)
        scan = float(np.mean(ix.scan_ms_history(STEPS)))
        res = (d_ids.cpu().numpy(), d_d.cpu().numpy())
        if base is None: base = res
        same = np.array_equal(res[0], base[0]) and np.array_equal(res[1], base[1])
        print(f"flat_share {share} seed - {k:2d}: scan {scan:.4f} ms  results {'same' if same else 'DIFFER (seed below the true bound)'}", flush=True)
ix.close()
