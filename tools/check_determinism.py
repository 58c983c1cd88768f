import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qadc_b200
rng = np.random.default_rng(1235)
n, dim, m, R, nq = int(os.environ.get("N", 10 ** 6)), 128, 16, 100, int(os.environ.get("NQ", 10000))
cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.load_flat(codes, 0.01)
q = rng.standard_normal((nq, dim)).astype(np.float32)
out = ix.build_tables(q, 1, R)
assign = out["assign"]; qt = out["qtables"]
base = None
for qb in (4, 1, 2, 4, 1, 2):
    ix.set_option("flat_qb", qb)
    for rep in range(3):
        ids, d, cnt = ix.scan_with_tables(assign, qt, R)
        if base is None:
            base = (ids.copy(), d.copy())
        badq = np.nonzero((ids != base[0]).any(1) | (d != base[1]).any(1))[0]
        extra = ""
        if len(badq):
            s = badq[0]
            miss = set(base[0][s].tolist()) - set(ids[s].tolist()); add = set(ids[s].tolist()) - set(base[0][s].tolist())
            extra = f" e.g. q={s} missing={sorted(miss)[:4]} extra={sorted(add)[:4]}"
        print(f"qb={qb} rep={rep} differing queries: {len(badq)}{extra}", flush=True)
