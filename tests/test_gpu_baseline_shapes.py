"""GPU parity at BASELINE.json's own shapes (-m gpu): the CUDA path through the C ABI against the oracle
(bit-exact) AND against the live unmodified reference (oracle/_ref) for a handful of queries of

  config 1   1M x 128-d, flat PQ 16x4, keep 1 %
  config 2   1M x 128-d, IVF-4096, nprobe 64, full qadc_search
  config 3   10M x 96-d, flat PQ 32x4 (sq_dim 3), keep 0.1 %
  config 5   scaled: 20M x 96-d, IVF-65536, nprobe 128, PQ 16x4 (sq_dim 6)

Protocol per query: Stage S the whole search == oracle.search bit for bit; Stage T float
tables / qmin / qmax within 1e-5 of the reference (norm-scaled for its sgemm-form builder), int8
tables within one LSB; Stage R the reference's own int8 tables injected on both sides: canonical result
tie-class equivalent to the raw reference heap.  The reference's coarse assignment is wrong for K > 256
(neighbors.cpp:64), so the inverted-list configs hand the fixed assignment (device == oracle,
asserted) to the reference's scanner, exactly as SURVEY §8c prescribes.
"""
import numpy as np
import pytest

import synth
from test_oracle import FLOAT_RTOL, rel_err, blas_scale, tie_class_check

pytestmark = pytest.mark.gpu
R = 100


def _ref_stage(ref_h, ref, oracle, db, q, assign, r):
    """The reference on one query with a given assignment: its float tables (single-vector builder for
    flat/ma = 1, sgemm form otherwise, query_common.hpp:292-297), bounds, int8 tables and raw heap."""
    m, cb = db["m"], db["codebooks"]
    ivf = "centroids" in db
    resid = (q[None, :] - db["centroids"][assign]).astype(np.float32) if ivf else q[None, :].copy()
    tables = ref.tables(resid, m, cb, blas_form=len(assign) > 1)
    qmin, qmax = ref_h.query_bounds(assign, tables, r)
    keys, vals, n = ref_h.query_scan(assign, tables, r)
    qt = ref.quantize(np.maximum(tables, 0), np.float32(qmin), np.float32(qmax))
    return resid, tables, qmin, qmax, qt, keys, vals, n


def _check(qadc, oracle, ref, ix, ref_h, db, queries, ma, r, min_checked=1):
    ivf = "centroids" in db
    m = db["m"]
    codes, offsets = db["codes"], db["offsets"]
    labels = db.get("labels")
    # Stage S: whole search, bit-exact against the oracle
    ids, d, cnt = ix.search(queries, ma, r)
    exp = oracle.search(db, queries, ma, r)
    assert exp["rc"] == 0
    assert np.array_equal(cnt, exp["count"]) and np.array_equal(d, exp["d"]) and np.array_equal(ids, exp["ids"])
    out = ix.build_tables(queries, ma, r)
    assert out["rc"] == 0
    assert np.array_equal(out["assign"], exp["assign"])
    assert np.array_equal(out["tables"], exp["tables"]) and np.array_equal(out["qtables"], exp["qtables"])
    assert np.array_equal(out["qmin"], exp["qmin"]) and np.array_equal(out["qmax"], exp["qmax"])
    if ivf:   # the fixed coarse assignment, also straight from the oracle's find_k_neighbors restatement
        assert np.array_equal(out["assign"], oracle.coarse_assign(queries, db["centroids"], ma)[0])
    checked = 0
    for qi in range(queries.shape[0]):
        a = out["assign"][qi]
        resid, rt, rqmin, rqmax, rqt, rkeys, rvals, rn = _ref_stage(ref_h, ref, oracle, db, queries[qi], a, r)
        # Stage T against the reference
        scale = blas_scale(resid, db["codebooks"], m) if len(a) > 1 else 1e-30
        assert rel_err(out["tables"][qi], rt, scale) <= FLOAT_RTOL
        assert abs(out["qmax"][qi] - rqmax) <= FLOAT_RTOL * rqmax
        assert abs(out["qmin"][qi] - rqmin) <= FLOAT_RTOL * max(float(rt.max()), 1e-30)
        lsb = np.abs(out["qtables"][qi].astype(int) - rqt.astype(int))
        assert lsb.max() <= 1 and (lsb != 0).mean() <= 0.02
        # Stage R: the reference's int8 tables on both sides
        g_ids, g_d, g_cnt = ix.scan_with_tables(a[None, :], rqt[None], r)
        d_parts = [oracle.distances(codes[offsets[p]:offsets[p + 1]], rqt[k]) if offsets[p + 1] > offsets[p]
                   else np.zeros(0, np.int8) for k, p in enumerate(a)]
        l_parts = [None if labels is None else labels[offsets[p]:offsets[p + 1]] for p in a]
        e_ids, e_d, e_cnt = synth.canonical_from_distances(d_parts, l_parts, r)
        assert g_cnt[0] == e_cnt and np.array_equal(g_d[0], e_d) and np.array_equal(g_ids[0], e_ids)
        rk, rv = np.zeros(r, np.uint32), np.full(r, 127, np.int8)
        rk[:rn], rv[:rn] = rkeys[:rn], rvals[:rn]
        checked += tie_class_check(g_ids[0], g_d[0], int(g_cnt[0]), rk, rv, d_parts, l_parts)
    assert checked >= min_checked
    return ids, d, cnt


def _flat(qadc, oracle, ref, n, dim, m, keep, nq, seed):
    rng = np.random.default_rng(seed)
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    db = dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=keep, offsets=np.array([0, n], np.int64))
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb)
    ix.load_flat(codes, keep)
    h = ref.flat(dim, m, cb, codes)
    h.prepare(keep)
    try:
        _check(qadc, oracle, ref, ix, h, db, q, 1, R)
    finally:
        h.close()
        ix.close()


def _ivf(qadc, oracle, ref, n, dim, m, K, ma, keep, nq, seed):
    rng = np.random.default_rng(seed)
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m)
    q = synth.make_queries(rng, nq, dim)
    db = dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=keep, offsets=offsets)
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb)
    ix.set_coarse(cents)
    ix.load_ivf(codes, labels, offsets, keep)
    h = ref.ivf(dim, m, cb, cents, codes, labels, offsets)
    h.prepare(keep)
    try:
        _check(qadc, oracle, ref, ix, h, db, q, ma, R)
    finally:
        h.close()
        ix.close()


def test_config1_flat_1m_16x4(qadc, oracle, ref):
    _flat(qadc, oracle, ref, 10 ** 6, 128, 16, 0.01, 8, 101)


def test_config2_ivf4096_nprobe64(qadc, oracle, ref):
    _ivf(qadc, oracle, ref, 10 ** 6, 128, 16, 4096, 64, 0.01, 8, 102)


def test_config3_flat_10m_32x4_dim96(qadc, oracle, ref):
    _flat(qadc, oracle, ref, 10 ** 7, 96, 32, 0.001, 4, 103)


def test_config5_scaled_ivf65536_nprobe128_dim96(qadc, oracle, ref):
    _ivf(qadc, oracle, ref, 20 * 10 ** 6, 96, 16, 65536, 128, 0.01, 6, 105)


def test_dim960_tables_need_more_than_48k_shared_memory(qadc, oracle):
    """GIST-shaped vectors (960-d): the table kernel's dynamic shared memory (64 bytes per dimension)
    passes the 48 KB default and must be opted in."""
    rng = np.random.default_rng(960)
    dim, m, n, nq = 960, 16, 20000, 5
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb)
    ix.load_flat(codes, 0.05)
    ids, d, cnt = ix.search(q, 1, R)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=0.05, offsets=np.array([0, n], np.int64)),
                        q, 1, R, want_tables=False)
    assert np.array_equal(cnt, exp["count"]) and np.array_equal(d, exp["d"]) and np.array_equal(ids, exp["ids"])
    ix.close()
    with pytest.raises(qadc.QadcError):
        qadc.Index(0).set_pq(16 * 240, 16, np.zeros((16, 16, 240), np.float32))
