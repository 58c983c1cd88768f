#!/usr/bin/env python
"""Experiments on BASELINE config 1 (1M x 16x4 flat, 10 000 queries): keep, queries per pass, filter, chunking.
usage: python tools/exp_c1.py [ncu]   (ncu: one search only, for a profiler capture)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, qadc_b200, bench_legs as bl

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
n, dim, m, nq = int(float(os.environ.get("N", "1e6"))), int(os.environ.get("DIM", "128")), int(os.environ.get("M", "16")), 10000
rng = np.random.default_rng(1235)
cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
q = rng.standard_normal((nq, dim)).astype(np.float32)
ncu = len(sys.argv) > 1 and sys.argv[1] == "ncu"
for keep in ((0.01,) if ncu else tuple(float(x) for x in os.environ.get("KEEPS", "0.01").split(","))):
    ix = qadc_b200.Index(0, stream.cuda_stream)
    ix.set_pq(dim, m, cb); ix.load_flat(codes, keep)
    for qb, filt, chunks, ring in ((int(os.environ.get("QB", "2")), 1, 0, int(os.environ.get("RING", "1"))),) if ncu else ((1, 1, 0, 1), (2, 1, 0, 0), (2, 1, 0, 1), (4, 1, 0, 0), (4, 1, 0, 1)):
        ix.set_option("flat_qb", qb); ix.set_option("flat_filter", filt); ix.set_option("flat_chunks", chunks); ix.set_option("flat_ring", ring)
        if ncu:
            d_q = torch.from_numpy(q).to(dev)
            d_ids = torch.empty((nq, 100), dtype=torch.int32, device=dev); d_d = torch.empty((nq, 100), dtype=torch.int8, device=dev)
            d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
            for _ in range(2):
                ix.search_device(d_q.data_ptr(), nq, 1, 100, d_ids.data_ptr(), d_d.data_ptr(), d_cnt.data_ptr())
            ix.synchronize()
            continue
        t, _ = bl.time_search(torch, ix, stream, dev, q, 1)
        print(f"keep {keep} qb {qb} filter {filt} chunks {chunks} warp-rings {ring}: ms {t['ms']:.3f} scan {t['scan_kernel_ms']:.3f} "
              f"pairs/s {n * nq / t['ms'] / 1e6:.1f} G", flush=True)
    ix.close()
