"""Every flat-scan variant (queries per pass, chunking, ring type) against the one-query-per-pass kernel on one database (also a racecheck target)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qadc_b200 as qadc, synth
rng = np.random.default_rng(77)
n, dim, m, nq, r = int(os.environ.get("N", 300000)), 128, 16, int(os.environ.get("NQ", 3000)), 100
cb = synth.make_pq(rng, dim, m); codes = synth.make_codes(rng, n, m); q = synth.make_queries(rng, nq, dim)
ix = qadc.Index(0); ix.set_pq(dim, m, cb); ix.load_flat(codes, 0.01)
out = ix.build_tables(q, 1, r)
base = None
variants = [(1, 0, 1)] + [tuple(int(x) for x in v.split(":")) for v in os.environ.get("VARIANTS", "4:0:1,2:0:1,4:0:0,2:0:0,4:0:1").split(",")]
for qb, chunks, ring in variants:
    ix.set_option("flat_qb", qb); ix.set_option("flat_chunks", chunks); ix.set_option("flat_ring", ring)
    ids, d, cnt = ix.scan_with_tables(out["assign"], out["qtables"], r)
    if base is None: base = (ids, d, cnt); continue
    bad = np.nonzero((ids != base[0]).any(1) | (d != base[1]).any(1) | (cnt != base[2]))[0]
    print("qb", qb, "chunks", chunks, "ring", ring, "bad queries", len(bad), bad[:10])
    for s in bad[:6]:
        dist = ix.dump_distances(0, out["qtables"][s, 0])
        extra = [int(x) for x in ids[s] if x not in set(base[0][s])]
        miss = [int(x) for x in base[0][s] if x not in set(ids[s])]
        print("   q", s, "extra", [(e, int(dist[e]), int(d[s][list(ids[s]).index(e)])) for e in extra], "missing", [(e, int(dist[e])) for e in miss],
              "r-th", int(base[1][s][-1]), "group", s // qb * qb)
        for o in range(s // qb * qb, s // qb * qb + qb):
            if o != s:
                od = ix.dump_distances(0, out["qtables"][o, 0])
                print("        under query", o, [(e, int(od[e])) for e in extra + miss], "bad" if o in bad else "")
