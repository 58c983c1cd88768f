import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qadc_b200
from oracle.pyoracle import Oracle
oracle = Oracle()
rng = np.random.default_rng(1235)
n, dim, m, R = 10 ** 6, 128, 16, 100
cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.load_flat(codes, 0.01)
db = dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=0.01, offsets=np.array([0, n], np.int64))
for nq in (8, 200, 10000):
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    sel = np.unique(np.linspace(0, nq - 1, 12).astype(int))
    exp = oracle.search(db, q[sel], 1, R, want_tables=False)
    for qb in (1, 2, 4):
        ix.set_option("flat_qb", qb)
        for rep in range(2):
            ids, d, cnt = ix.search(q, 1, R)
            bad = [int(s) for k, s in enumerate(sel) if not (np.array_equal(ids[s], exp["ids"][k]) and np.array_equal(d[s], exp["d"][k]))]
            msg = ""
            if bad:
                s = bad[0]; k = list(sel).index(s)
                nd = int((d[s] != exp["d"][k]).sum()); ni = int((ids[s] != exp["ids"][k]).sum())
                msg = f" first bad q={s}: {nd} dist diffs, {ni} id diffs; gpu d[:8]={d[s][:8]} exp={exp['d'][k][:8]} cnt={cnt[s]} vs {exp['count'][k]}"
            print(f"nq={nq} qb={qb} rep={rep} bad={len(bad)}/{len(sel)}{msg}", flush=True)
