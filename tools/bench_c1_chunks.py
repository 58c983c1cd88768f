import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, qadc_b200
rng = np.random.default_rng(1235)
n, dim, m, nq, R = 10 ** 6, 128, 16, 10000, 100
cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
q = rng.standard_normal((nq, dim)).astype(np.float32)
ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.load_flat(codes, 0.01)
ix.set_option("time_scan", 1)
for qb in (1, 2, 4):
    for chunks in (1, 2, 4, 8, 16):
        ix.set_option("flat_qb", qb); ix.set_option("flat_chunks", chunks)
        ix.search(q, 1, R); ix.search(q, 1, R)
        print(f"qb={qb} chunks={chunks} scan kernel {ix.last_scan_ms():.2f} ms -> {n*nq/ix.last_scan_ms()/1e6:.0f} G pairs/s", flush=True)
