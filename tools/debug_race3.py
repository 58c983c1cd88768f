import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qadc_b200
rng = np.random.default_rng(1235)
n, dim, m, R, nq = 10 ** 6, 128, 16, 100, 3000
cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.load_flat(codes, 0.01)
q = rng.standard_normal((nq, dim)).astype(np.float32)
out = ix.build_tables(q, 1, R)
assign = out["assign"]; qt = out["qtables"]
ix.set_option("flat_qb", 4)
base_ids, base_d, _ = ix.scan_with_tables(assign, qt, R)
ix.set_option("flat_qb", 1)
for rep in range(3):
    ids, d, cnt = ix.scan_with_tables(assign, qt, R)
    badq = np.nonzero((ids != base_ids).any(1) | (d != base_d).any(1))[0]
    print(f"{os.environ.get('QADC_LIB','default').split('_')[-1]} rep {rep}: {len(badq)} of {nq} differ; first {badq[:6].tolist()} last {badq[-3:].tolist()}", flush=True)
