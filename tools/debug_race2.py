import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qadc_b200
rng = np.random.default_rng(1235)
n, dim, m, R, nq = 10 ** 6, 128, 16, 100, 10000
cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.load_flat(codes, 0.01)
q = rng.standard_normal((nq, dim)).astype(np.float32)
out = ix.build_tables(q, 1, R)
assign = out["assign"]; qt = out["qtables"]
ix.set_option("flat_qb", 4)
base_ids, base_d, _ = ix.scan_with_tables(assign, qt, R)
ix.set_option("flat_qb", 1)
for sub in (10000, 2000, 500, 148, 1):
    ids, d, cnt = ix.scan_with_tables(assign[:sub], qt[:sub], R)
    badq = np.nonzero((ids != base_ids[:sub]).any(1) | (d != base_d[:sub]).any(1))[0]
    print(f"first {sub} queries: {len(badq)} differ; first few {badq[:10].tolist()}")
# the queries that failed in the 10000 run, alone and in small groups
ids, d, cnt = ix.scan_with_tables(assign, qt, R)
badq = np.nonzero((ids != base_ids).any(1))[0]
print("bad in full run:", len(badq), badq[:20].tolist())
for s in badq[:5]:
    i1, d1, _ = ix.scan_with_tables(assign[s:s + 1], qt[s:s + 1], R)
    print(f"q={s} alone: equal={np.array_equal(i1[0], base_ids[s])}")
    dd = np.nonzero(d[s] != base_d[s])[0]
    print("   dist diffs at ranks", dd[:10].tolist(), "gpu", d[s][dd[:6]].tolist(), "base", base_d[s][dd[:6]].tolist())
    miss = sorted(set(base_ids[s].tolist()) - set(ids[s].tolist()))
    print("   missing ids", miss, " (id//256)%15 =", [(x // 256) % 15 for x in miss], " id//256//15 =", [(x // 256) // 15 for x in miss])
