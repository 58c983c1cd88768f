"""compute-sanitizer driver: one flat search per scan variant on a database large enough for the TMA ring to wrap.
usage: compute-sanitizer --tool racecheck python tools/racecheck_flat.py [qb] [m] [ring]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qadc_b200
qb, m, ring = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(1)
dim, n, nq, r = 8 * m, 200000, 4, 100
cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
q = rng.standard_normal((nq, dim)).astype(np.float32)
ix = qadc_b200.Index(0)
ix.set_pq(dim, m, cb)
ix.load_flat(codes, 0.01)
ix.set_option("flat_qb", qb); ix.set_option("flat_ring", ring); ix.set_option("flat_chunks", 4)
ids, d, cnt = ix.search(q, 1, r)
print("ok", qb, m, ring, int(ids[0, 0]), int(cnt[0]))
