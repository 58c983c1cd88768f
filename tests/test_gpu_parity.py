"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs and against the golden vectors of the unmodified reference.

Parity protocol: Stage T float tables/bounds (tolerance 1e-5 relative; bit-equal
to the oracle, which uses the same explicit float operations), Stage D per-vector int8
distances bit-exact, Stage S canonical top-r bit-exact, Stage R tie-class equivalence against
the raw reference heap."""
import os

import numpy as np
import pytest

import synth
from test_oracle import (FLAT, IVF, OPQ, ADC, FLOAT_RTOL, load, rel_err, blas_scale, tie_class_check, golden_db, rotated,
                         adc_db, check_adc_against_reference)

pytestmark = pytest.mark.gpu


def flat_index(qadc, dim, m, cb, codes, keep, rotation=None):
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb, rotation=rotation)
    ix.load_flat(codes, keep)
    return ix


def ivf_index(qadc, dim, m, cb, cents, codes, labels, offsets, keep, rotation=None):
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb, rotation=rotation)
    ix.set_coarse(cents)
    ix.load_ivf(codes, labels, offsets, keep)
    return ix


def golden_index(qadc, g):
    """Device index of a golden fixture (flat / inverted lists, with its OPQ rotation if any)."""
    rot = g["rotation"] if "rotation" in g else None
    if "centroids" in g:
        return ivf_index(qadc, int(g["dim"]), int(g["m"]), g["codebooks"], g["centroids"], g["codes"], g["labels"],
                         g["offsets"], float(g["keep"]), rot)
    return flat_index(qadc, int(g["dim"]), int(g["m"]), g["codebooks"], g["codes"], float(g["keep"]), rot)


# ---- layout ------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m", [(1, 16), (255, 16), (256, 16), (257, 32), (3003, 16), (70000, 32)])
def test_layout_round_trip(qadc, n, m):
    rng = np.random.default_rng(n + m)
    codes = synth.make_codes(rng, n, m)
    ix = flat_index(qadc, 8 * m, m, synth.make_pq(rng, 8 * m, m), codes, 1.0)
    assert np.array_equal(ix.download_codes(0), codes)
    ix.close()


def test_chunked_upload_equals_single_upload(qadc):
    rng = np.random.default_rng(3)
    m, n = 16, 256 * 9 + 77
    codes = synth.make_codes(rng, n, m)
    ix = qadc.Index(0)
    ix.set_pq(128, m, synth.make_pq(rng, 128, m))
    ix.begin_database([n], False)
    for first in range(0, n, 768):
        ix.upload_codes(0, first, codes[first:first + 768])
    ix.finalize(0.1)
    assert np.array_equal(ix.download_codes(0), codes)
    ix.close()


# ---- Stage D: per-vector distances ---------------------------------------------------------
@pytest.mark.parametrize("n,m", [(1, 16), (300, 16), (4099, 16), (2048, 32), (50001, 32), (200000, 16)])
def test_distances_bit_exact_vs_oracle(qadc, oracle, n, m):
    rng = np.random.default_rng(n * 7 + m)
    codes = synth.make_codes(rng, n, m)
    ix = flat_index(qadc, 8 * m, m, synth.make_pq(rng, 8 * m, m), codes, 1.0)
    for p_sat in (0.0, 0.15, 0.6):
        qt = synth.make_qtables(rng, (), m, hi=128 // m * 3, p_sat=p_sat)
        assert np.array_equal(ix.dump_distances(0, qt), oracle.distances(codes, qt))
    # extreme tables: all zero, all 127, single large entries
    for qt in (np.zeros((m, 16), np.int8), np.full((m, 16), 127, np.int8)):
        assert np.array_equal(ix.dump_distances(0, qt), oracle.distances(codes, qt))
    ix.close()


@pytest.mark.parametrize("name", FLAT)
def test_distances_match_reference_golden(qadc, name):
    g = load(name)
    ix = flat_index(qadc, int(g["dim"]), int(g["m"]), g["codebooks"], g["codes"], float(g["keep"]))
    for q in range(g["queries"].shape[0]):
        assert np.array_equal(ix.dump_distances(0, g["ref_qtables"][q, 0]), g["ref_distances"][q])   # scan_avx_4 values
    ix.close()


def test_negative_table_entries_rejected(qadc):
    rng = np.random.default_rng(1)
    ix = flat_index(qadc, 128, 16, synth.make_pq(rng, 128, 16), synth.make_codes(rng, 100, 16), 1.0)
    qt = np.zeros((16, 16), np.int8)
    qt[3, 5] = -1
    with pytest.raises(qadc.QadcError):
        ix.dump_distances(0, qt)
    ix.close()


# ---- Stage S: canonical top-r with injected tables -----------------------------------------
@pytest.mark.parametrize("n,m,nq,r", [(300, 16, 3, 10), (5000, 16, 9, 100), (70001, 16, 5, 100), (4096, 32, 4, 50),
                                      (60000, 32, 7, 100), (1000, 16, 2, 700), (50, 16, 2, 100)])
@pytest.mark.parametrize("qb", [0, 1, 2, 4])
def test_flat_scan_with_tables_bit_exact(qadc, oracle, n, m, nq, r, qb):
    rng = np.random.default_rng(n + 13 * m + nq)
    codes = synth.make_codes(rng, n, m)
    ix = flat_index(qadc, 8 * m, m, synth.make_pq(rng, 8 * m, m), codes, 1.0)
    ix.set_option("flat_qb", qb)
    qt = synth.make_qtables(rng, (nq, 1), m, hi=max(3, 200 // m), p_sat=0.1)
    ids, d, cnt = ix.scan_with_tables(np.zeros((nq, 1), np.int32), qt, r)
    offsets = np.array([0, n], np.int64)
    for q in range(nq):
        e_ids, e_d, e_cnt, _ = oracle.scan_with_tables(codes, None, offsets, np.zeros(1, np.int32), qt[q], r)
        assert cnt[q] == e_cnt
        assert np.array_equal(d[q], e_d)
        assert np.array_equal(ids[q], e_ids)
    ix.close()


@pytest.mark.parametrize("n,nq,r,chunks,filt", [(300, 3, 10, 0, 1), (70001, 5, 100, 0, 1), (123457, 4, 100, 3, 1), (123457, 4, 100, 64, 0),
                                                (1000, 2, 700, 0, 1), (50, 2, 100, 0, 1), (400000, 6, 100, 0, 1), (400000, 3, 1, 5, 1)])
def test_flat_warp_ring_kernel_bit_exact(qadc, oracle, n, nq, r, chunks, filt):
    """scan_flat_wr_kernel (per-warp TMA rings, option flat_ring = 1): canonical top-r bit-exact against the oracle, with
    and without the pre-filter, across chunkings, tiny and ragged databases, r = 1 and the two-pass emit (r = 700)."""
    rng = np.random.default_rng(n + nq + r)
    m = 16
    codes = synth.make_codes(rng, n, m)
    ix = flat_index(qadc, 128, m, synth.make_pq(rng, 128, m), codes, 1.0)
    ix.set_option("flat_ring", 1)
    ix.set_option("flat_qb", 1)
    ix.set_option("flat_chunks", chunks)
    ix.set_option("flat_filter", filt)
    qt = synth.make_qtables(rng, (nq, 1), m, hi=14, p_sat=0.05)
    ids, d, cnt = ix.scan_with_tables(np.zeros((nq, 1), np.int32), qt, r)
    offsets = np.array([0, n], np.int64)
    for q in range(nq):
        e_ids, e_d, e_cnt, _ = oracle.scan_with_tables(codes, None, offsets, np.zeros(1, np.int32), qt[q], r)
        assert cnt[q] == e_cnt and np.array_equal(d[q], e_d) and np.array_equal(ids[q], e_ids)
    ix.close()


@pytest.mark.parametrize("chunks", [1, 3, 64])
def test_flat_scan_independent_of_chunking(qadc, oracle, chunks):
    rng = np.random.default_rng(99)
    n, m, nq, r = 123457, 16, 3, 100
    codes = synth.make_codes(rng, n, m)
    ix = flat_index(qadc, 128, m, synth.make_pq(rng, 128, m), codes, 1.0)
    ix.set_option("flat_chunks", chunks)
    qt = synth.make_qtables(rng, (nq, 1), m, hi=14, p_sat=0.05)
    ids, d, cnt = ix.scan_with_tables(np.zeros((nq, 1), np.int32), qt, r)
    for q in range(nq):
        e_ids, e_d, e_cnt, _ = oracle.scan_with_tables(codes, None, np.array([0, n], np.int64), np.zeros(1, np.int32), qt[q], r)
        assert cnt[q] == e_cnt and np.array_equal(d[q], e_d) and np.array_equal(ids[q], e_ids)
    ix.close()


def test_flat_scan_degenerate_ties(qadc, oracle):
    """All vectors identical / all-zero tables: the answer is the first r positions."""
    n, m, r = 10000, 16, 100
    codes = np.zeros((n, m // 2), np.uint8)
    ix = flat_index(qadc, 128, m, synth.make_pq(np.random.default_rng(0), 128, m), codes, 1.0)
    for qb in (1, 2, 4):   # 4 queries per pass uses half-superblock candidate lists (two-pass emit)
        ix.set_option("flat_qb", qb)
        for val in (0, 5):
            qt = np.full((5, 1, m, 16), val, np.int8)
            ids, d, cnt = ix.scan_with_tables(np.zeros((5, 1), np.int32), qt, r)
            assert np.all(cnt == r) and np.all(d == min(127, val * m))
            for q in range(5):
                assert np.array_equal(ids[q], np.arange(r, dtype=np.uint32))
    # nothing below 127 -> sentinels only (db_query_4.cpp:276)
    qt = np.full((1, 1, m, 16), 127, np.int8)
    ids, d, cnt = ix.scan_with_tables(np.zeros((1, 1), np.int32), qt, r)
    assert cnt[0] == 0 and np.all(ids == 0) and np.all(d == 127)
    ix.close()


@pytest.mark.parametrize("n_distinct", [1, 2, 40])
def test_flat_merge_of_many_full_lists(qadc, oracle, n_distinct):
    """More per-CTA lists than the merge buffer holds (31 chunks x r = 100 keys > 2048 slots), every list full: the merge
    first narrows its bound to the r-th distance of the union (128-bin histogram of the keys' distances).  One distinct
    distance = every key ties with the r-th (the narrowed bound keeps them all and the streaming rounds take over), two =
    the answer straddles the two classes, 40 = a few dozen classes."""
    rng = np.random.default_rng(500 + n_distinct)
    n, m, r, nq = 123457, 16, 100, 4
    codes = np.zeros((n, m // 2), np.uint8)
    codes[:, 0] = rng.integers(0, min(n_distinct, 16), n)                    # low nibble of byte 0 = sub-quantiser 0
    codes[:, 1] = rng.integers(0, max(1, n_distinct // 16), n)               # low nibble of byte 1 = sub-quantiser 2
    ix = flat_index(qadc, 128, m, synth.make_pq(rng, 128, m), codes, 1.0)
    ix.set_option("flat_qb", 1)
    ix.set_option("flat_chunks", 64)
    qt = np.zeros((nq, 1, m, 16), np.int8)
    qt[:, 0, 0, :] = rng.integers(0, 60, (nq, 16))
    qt[:, 0, 2, :] = rng.integers(0, 60, (nq, 16))
    ids, d, cnt = ix.scan_with_tables(np.zeros((nq, 1), np.int32), qt, r)
    for q in range(nq):
        e_ids, e_d, e_cnt, _ = oracle.scan_with_tables(codes, None, np.array([0, n], np.int64), np.zeros(1, np.int32), qt[q], r)
        assert cnt[q] == e_cnt and np.array_equal(d[q], e_d) and np.array_equal(ids[q], e_ids)
    ix.close()


@pytest.mark.parametrize("name", IVF + OPQ)
def test_ivf_scan_with_reference_tables(qadc, oracle, name):
    """Injected assign + the reference's own int8 tables: canonical result bit-exact vs the
    oracle, tie-class equivalent to the raw reference heap (Stage R)."""
    g = load(name)
    r, m, ma = int(g["r"]), int(g["m"]), int(g["ma"])
    codes, labels, offsets = g["codes"], g["labels"], g["offsets"]
    ix = golden_index(qadc, g)
    ids, d, cnt = ix.scan_with_tables(g["ref_assign"], g["ref_qtables"], r)
    checked = 0
    for q in range(g["queries"].shape[0]):
        a = g["ref_assign"][q]
        e_ids, e_d, e_cnt, _ = oracle.scan_with_tables(codes, labels, offsets, a, g["ref_qtables"][q], r)
        assert cnt[q] == e_cnt and np.array_equal(d[q], e_d) and np.array_equal(ids[q], e_ids)
        d_parts = [oracle.distances(codes[offsets[p]:offsets[p + 1]], g["ref_qtables"][q, k])
                   if offsets[p + 1] > offsets[p] else np.zeros(0, np.int8) for k, p in enumerate(a)]
        l_parts = [labels[offsets[p]:offsets[p + 1]] for p in a]
        checked += tie_class_check(ids[q], d[q], cnt[q], g["ref_heap_keys"][q], g["ref_heap_vals"][q], d_parts, l_parts)
    assert checked >= 1
    ix.close()


@pytest.mark.parametrize("name", FLAT)
def test_flat_scan_vs_reference_heap(qadc, oracle, name):
    g = load(name)
    r = int(g["r"])
    ix = flat_index(qadc, int(g["dim"]), int(g["m"]), g["codebooks"], g["codes"], float(g["keep"]))
    ids, d, cnt = ix.scan_with_tables(np.zeros((g["queries"].shape[0], 1), np.int32), g["ref_qtables"], r)
    checked = 0
    for q in range(g["queries"].shape[0]):
        checked += tie_class_check(ids[q], d[q], cnt[q], g["ref_heap_keys"][q], g["ref_heap_vals"][q],
                                   [g["ref_distances"][q]], [None])
    assert checked >= 1
    ix.close()


def test_ivf_random_large(qadc, oracle):
    rng = np.random.default_rng(21)
    n, K, m, ma, nq, r = 40000, 64, 16, 8, 12, 100
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(5, 17))
    cb = synth.make_pq(rng, 128, m)
    cents = rng.standard_normal((K, 128)).astype(np.float32)
    ix = ivf_index(qadc, 128, m, cb, cents, codes, labels, offsets, 0.05)
    assign = np.stack([rng.permutation(K)[:ma] for _ in range(nq)]).astype(np.int32)
    assign[0, 2] = 5   # an empty partition among the probes
    qt = synth.make_qtables(rng, (nq, ma), m, hi=12, p_sat=0.1)
    ids, d, cnt = ix.scan_with_tables(assign, qt, r)
    for q in range(nq):
        e_ids, e_d, e_cnt, _ = oracle.scan_with_tables(codes, labels, offsets, assign[q], qt[q], r)
        assert cnt[q] == e_cnt and np.array_equal(d[q], e_d) and np.array_equal(ids[q], e_ids)
    ix.close()


# ---- long flat keep-prefixes: int8 lower bounds in front of the float prefix scan (qadc_flatprep.cuh) ------------
def _flat_prefix_case(rng, kind, n, m):
    codes = synth.make_codes(rng, n, m)
    if kind == "identical":            # every prefix vector ties: more candidates than slots, every vector evaluated
        codes[:] = codes[0]
    elif kind == "two_values":         # two distinct vectors: the r-th smallest sits inside one huge tie class
        codes[:] = codes[rng.integers(0, 2, n)]
    elif kind == "sample_unlike_rest": # the first 65 536 vectors (the provisional bound's sample) are all the same vector
        codes[:65536] = codes[0]
    return codes


@pytest.mark.parametrize("kind,m,dim,n,keep,r", [
    ("random", 16, 128, 300000, 0.5, 100), ("random", 32, 96, 280000, 0.5, 16), ("random", 16, 64, 270000, 0.5, 512),
    ("random", 16, 128, 270000, 0.5, 513), ("random", 16, 128, 300000, 0.45, 1), ("identical", 16, 128, 270000, 0.5, 100),
    ("two_values", 32, 128, 270000, 0.5, 50), ("sample_unlike_rest", 16, 128, 300000, 0.5, 100),
    ("random", 16, 128, 60000, 0.5, 100)])
def test_flat_long_prefix_bounds_bit_exact(qadc, oracle, kind, m, dim, n, keep, r):
    """qmax / qmin / int8 tables of a flat database whose keep-prefix is long enough for the int8 pre-selection
    (>= 131 072 vectors): bit-identical to the plain float prefix scan (option flat_prep = 0) and to the oracle, and so
    are the search results; r = 513 is past the pre-selection's limit and the 30 000-vector prefix below its minimum
    (plain path either way)."""
    rng = np.random.default_rng(len(kind) * 1000 + m + r)
    cb = synth.make_pq(rng, dim, m)
    codes = _flat_prefix_case(rng, kind, n, m)
    q = synth.make_queries(rng, 6, dim)
    ix = flat_index(qadc, dim, m, cb, codes, keep)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=keep, offsets=np.array([0, n], np.int64)), q, 1, r)
    outs = {}
    for prep in (1, 0):
        ix.set_option("flat_prep", prep)
        out = ix.build_tables(q, 1, r)
        assert out["rc"] == 0
        for k in ("qmin", "qmax", "qtables"):
            assert np.array_equal(out[k], exp[k]), (prep, k)
        ids, d, cnt = ix.search(q, 1, r)
        assert np.array_equal(ids, exp["ids"]) and np.array_equal(d, exp["d"]) and np.array_equal(cnt, exp["count"]), prep
        # injected int8 tables: the bound is then seeded by the histogram pass over the nibble-plane prefix (its r-th
        # smallest sum found by the last CTA to finish) instead of the table pipeline's candidates
        ids, d, cnt = ix.scan_with_tables(out["assign"], out["qtables"], r)
        assert np.array_equal(ids, exp["ids"]) and np.array_equal(d, exp["d"]) and np.array_equal(cnt, exp["count"]), prep
        outs[prep] = out
    ix.close()


def test_flat_long_prefix_degenerate_scale(qadc, oracle):
    """Sub-quantiser tables that are almost constant next to their sum (every centroid of a sub-quantiser at nearly the
    same large distance from the query): the provisional scale is rejected and the pre-selection keeps every vector."""
    rng = np.random.default_rng(4242)
    dim, m, n, r = 64, 16, 280000, 100
    cb = (synth.make_pq(rng, dim, m) * 1e-3).astype(np.float32)
    q = (synth.make_queries(rng, 4, dim) + 50.0).astype(np.float32)
    codes = synth.make_codes(rng, n, m)
    ix = flat_index(qadc, dim, m, cb, codes, 0.5)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=0.5, offsets=np.array([0, n], np.int64)), q, 1, r)
    out = ix.build_tables(q, 1, r)
    assert out["rc"] == 0
    for k in ("qmin", "qmax", "qtables"):
        assert np.array_equal(out[k], exp[k]), k
    ids, d, cnt = ix.search(q, 1, r)
    assert np.array_equal(ids, exp["ids"]) and np.array_equal(d, exp["d"]) and np.array_equal(cnt, exp["count"])
    ix.close()


# ---- Stage T: tables, bounds, int8 tables ---------------------------------------------------
@pytest.mark.parametrize("name", FLAT + IVF + OPQ)
def test_table_pipeline_vs_oracle_and_reference(qadc, oracle, name):
    g = load(name)
    r, m = int(g["r"]), int(g["m"])
    ivf = "centroids" in g
    ma = int(g["ma"]) if ivf else 1
    ix = golden_index(qadc, g)
    out = ix.build_tables(g["queries"], ma, r)
    assert out["rc"] == 0
    exp = oracle.search(golden_db(g), g["queries"], ma, r)
    # the device uses the oracle's explicit float operations: expect bit equality
    assert np.array_equal(out["assign"], exp["assign"])
    assert np.array_equal(out["tables"], exp["tables"])
    assert np.array_equal(out["qmin"], exp["qmin"]) and np.array_equal(out["qmax"], exp["qmax"])
    assert np.array_equal(out["qtables"], exp["qtables"])
    # and the stated tolerance against the reference itself
    if ivf:
        assert np.array_equal(out["assign"], g["ref_assign"])
        for q in range(g["queries"].shape[0]):
            resid = rotated(g, g["queries"][q][None, :] - g["centroids"][out["assign"][q]])
            assert rel_err(out["tables"][q], g["ref_tables_used"][q], blas_scale(resid, g["codebooks"], m)) <= FLOAT_RTOL
    else:
        assert rel_err(out["tables"][:, 0], g["ref_tables_direct"], 1e-30) <= FLOAT_RTOL
    assert np.all(np.abs(out["qmax"] - g["ref_qmax"]) <= FLOAT_RTOL * g["ref_qmax"])
    # qmin is a minimum over BLAS-form entries on the reference side: same norm-scaled tolerance
    assert np.all(np.abs(out["qmin"] - g["ref_qmin"]) <= FLOAT_RTOL * float(g["ref_tables_used"].max()))
    lsb = np.abs(out["qtables"].astype(int) - g["ref_qtables"].astype(int))
    assert lsb.max() <= 1 and (lsb != 0).mean() <= 0.02
    ix.close()


def test_opq_rotation(qadc, oracle):
    rng = np.random.default_rng(8)
    dim, m, n, nq, r = 64, 16, 4000, 5, 20
    cb = synth.make_pq(rng, dim, m)
    rot = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb, rotation=rot)
    ix.load_flat(codes, 0.05)
    out = ix.build_tables(q, 1, r)
    db = dict(dim=dim, m=m, codebooks=cb, rotation=rot, codes=codes, keep=0.05, offsets=np.array([0, n], np.int64))
    exp = oracle.search(db, q, 1, r)
    assert np.array_equal(out["tables"], exp["tables"]) and np.array_equal(out["qtables"], exp["qtables"])
    ids, d, cnt = ix.search(q, 1, r)
    assert np.array_equal(ids, exp["ids"]) and np.array_equal(d, exp["d"]) and np.array_equal(cnt, exp["count"])
    ix.close()


@pytest.mark.parametrize("K,ma,order", [(1500, 16, "random"), (5000, 100, "random"), (4096, 64, "ascending"),
                                        (4096, 64, "descending"), (8192, 64, "clustered"), (2048, 128, "random"),
                                        (3000, 200, "random")])
def test_coarse_assignment_large_k(qadc, oracle, K, ma, order):
    """More than 256 cells: the fixed assignment (the reference's own is wrong there).
    K >= 2048 with ma <= 128 takes the pre-bounded selection; the orders are its corner cases:
    distances increasing / decreasing with the cell index, and all near cells dealt to 32 of the
    256 selection threads (cell index mod 256 < 32), which floods the streaming rounds."""
    rng = np.random.default_rng(12)
    dim, m, n, nq, r = 32, 16, 30000, 20, 50
    cb = synth.make_pq(rng, dim, m)
    cents = rng.standard_normal((K, dim)).astype(np.float32)
    q = synth.make_queries(rng, nq, dim)
    if order in ("ascending", "descending"):
        direction = rng.standard_normal(dim).astype(np.float32)
        scale = np.arange(K, dtype=np.float32) if order == "ascending" else np.arange(K, 0, -1).astype(np.float32)
        cents = (10.0 + scale[:, None]) * direction[None, :] / np.linalg.norm(direction)
        q = (0.01 * q).astype(np.float32)
    elif order == "clustered":
        near = (np.arange(K) % 256) < 32
        cents[~near] += 50.0
    codes, labels, offsets = synth.make_ivf(rng, n, K, m)
    ix = ivf_index(qadc, dim, m, cb, cents.astype(np.float32), codes, labels, offsets, 0.3)
    out = ix.build_tables(q, ma, r)
    exp_assign, _ = oracle.coarse_assign(q, cents.astype(np.float32), ma)
    assert np.array_equal(out["assign"], exp_assign)
    ix.close()


# ---- end to end ----------------------------------------------------------------------------
@pytest.mark.parametrize("name", FLAT + IVF + OPQ)
def test_search_end_to_end_golden_inputs(qadc, oracle, name):
    g = load(name)
    r, m = int(g["r"]), int(g["m"])
    ivf = "centroids" in g
    ma = int(g["ma"]) if ivf else 1
    db = golden_db(g)
    ix = golden_index(qadc, g)
    ids, d, cnt = ix.search(g["queries"], ma, r)
    exp = oracle.search(db, g["queries"], ma, r)
    assert np.array_equal(cnt, exp["count"]) and np.array_equal(d, exp["d"]) and np.array_equal(ids, exp["ids"])
    # recall-style agreement with the raw reference heap: same distance multiset wherever the
    # int8 tables agree exactly with the reference's
    for q in range(g["queries"].shape[0]):
        if np.array_equal(exp["qtables"][q], g["ref_qtables"][q]):
            real = g["ref_heap_vals"][q] < 127
            if len(np.unique(g["ref_heap_keys"][q][real])) == real.sum():
                assert np.array_equal(np.sort(d[q]), np.sort(g["ref_heap_vals"][q]))
    ix.close()


@pytest.mark.parametrize("n,m,dim,nq,r,keep", [(200000, 16, 128, 33, 100, 0.01), (150000, 32, 96, 17, 100, 0.01)])
def test_search_flat_medium(qadc, oracle, n, m, dim, nq, r, keep):
    rng = np.random.default_rng(n + m)
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    ix = flat_index(qadc, dim, m, cb, codes, keep)
    ids, d, cnt = ix.search(q, 1, r)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=keep, offsets=np.array([0, n], np.int64)),
                        q, 1, r, want_tables=False)
    assert np.array_equal(cnt, exp["count"]) and np.array_equal(d, exp["d"]) and np.array_equal(ids, exp["ids"])
    ix.close()


def test_search_reports_bound_error(qadc):
    """Prefix smaller than r -> QADC_EBOUND, the reference's 'Max quantization bound too high' exit."""
    rng = np.random.default_rng(5)
    ix = flat_index(qadc, 128, 16, synth.make_pq(rng, 128, 16), synth.make_codes(rng, 500, 16), 0.01)
    with pytest.raises(qadc.QadcError) as e:
        ix.search(synth.make_queries(rng, 2, 128), 1, 10)
    assert e.value.code == qadc.QADC_EBOUND
    ix.close()


def test_argument_errors(qadc):
    rng = np.random.default_rng(5)
    ix = qadc.Index(0)
    with pytest.raises(qadc.QadcError):
        ix.set_pq(128, 8, synth.make_pq(rng, 128, 8))          # (8,4) unsupported, db_query_4.cpp:32-34
    # an 8-bit quantiser is accepted for the plain ADC path only: the Quick ADC database refuses it
    # ("Quantizer must have sq_bits=4", load_database_check, db_query_4.cpp:393-402)
    ix.set_pq(128, 16, rng.standard_normal((16, 256, 8)).astype(np.float32), bits=8)
    with pytest.raises(qadc.QadcError) as e:
        ix.load_flat(rng.integers(0, 256, (1000, 16), dtype=np.uint8), 0.5)
    assert "sq_bits=4" in str(e.value)
    ix.set_pq(128, 16, synth.make_pq(rng, 128, 16))
    with pytest.raises(qadc.QadcError):
        ix.search(synth.make_queries(rng, 1, 128), 1, 10)      # no database yet
    ix.load_flat(synth.make_codes(rng, 1000, 16), 0.5)
    with pytest.raises(qadc.QadcError):
        ix.search(synth.make_queries(rng, 1, 128), 2, 10)      # flat needs ma = 1
    ix.close()


# ---- sharded flat database on one GPU: N "virtual" shards + merge == unsharded ---------------
@pytest.mark.parametrize("G", [2, 3, 8])
def test_virtual_shards_merge_equals_single(qadc, oracle, G):
    import torch
    rng = np.random.default_rng(G)
    dim, m, n, nq, r, keep = 128, 16, 100000, 9, 100, 0.01
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    single = flat_index(qadc, dim, m, cb, codes, keep)
    s_ids, s_d, s_cnt = single.search(q, 1, r)
    single.close()
    from qadc_b200 import sharding
    prefix = codes[:sharding.start_size(n, keep)]
    dq = torch.from_numpy(q).cuda()
    keys = torch.empty((G, nq, r), dtype=torch.int64, device="cuda")
    ids = torch.empty((G, nq, r), dtype=torch.int32, device="cuda")
    for g in range(G):
        lo, hi = sharding.flat_shard_range(n, g, G)
        ix = qadc.Index(0)
        ix.set_pq(dim, m, cb)
        ix.begin_database([hi - lo], False)
        ix.upload_codes(0, 0, codes[lo:hi])
        ix.set_position_base(0, lo)
        ix.set_prefix(0, prefix)
        ix.finalize(keep)
        d_tmp = torch.empty((nq, r), dtype=torch.int8, device="cuda")
        c_tmp = torch.empty(nq, dtype=torch.int32, device="cuda")
        ix.search_device(dq.data_ptr(), nq, 1, r, ids[g].data_ptr(), d_tmp.data_ptr(), c_tmp.data_ptr(), keys[g].data_ptr())
        ix.synchronize()
        ix.close()
    mi = qadc.Index(0)
    o_ids = torch.empty((nq, r), dtype=torch.int32, device="cuda")
    o_d = torch.empty((nq, r), dtype=torch.int8, device="cuda")
    o_c = torch.empty(nq, dtype=torch.int32, device="cuda")
    mi.merge_shards_device(keys.data_ptr(), ids.data_ptr(), G, nq, r, o_ids.data_ptr(), o_d.data_ptr(), o_c.data_ptr())
    mi.synchronize()
    assert np.array_equal(o_ids.cpu().numpy().view(np.uint32), s_ids)
    assert np.array_equal(o_d.cpu().numpy(), s_d)
    assert np.array_equal(o_c.cpu().numpy(), s_cnt)
    mi.close()


# ---- many queries / large databases: invariance properties ---------------------------------
def test_many_queries_all_scan_variants_agree(qadc, oracle):
    """Thousands of queries whose per-warp lists are all full (n_lists * r > one merge round):
    every queries-per-pass variant and chunking returns the same result, repeatedly
    (regression test for a racy count read in the merge kernel), spot-checked against the oracle."""
    rng = np.random.default_rng(77)
    n, dim, m, nq, r = 300000, 128, 16, 3000, 100
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    ix = flat_index(qadc, dim, m, cb, codes, 0.01)
    out = ix.build_tables(q, 1, r)
    base = None
    for qb, chunks in ((4, 0), (1, 0), (2, 0), (1, 3), (1, 0)):
        ix.set_option("flat_qb", qb)
        ix.set_option("flat_chunks", chunks)
        ids, d, cnt = ix.scan_with_tables(out["assign"], out["qtables"], r)
        if base is None:
            base = (ids, d, cnt)
            for s in (0, 1499, 2999):
                e_ids, e_d, e_cnt, _ = oracle.scan_with_tables(codes, None, np.array([0, n], np.int64), out["assign"][s],
                                                               out["qtables"][s], r)
                assert np.array_equal(ids[s], e_ids) and np.array_equal(d[s], e_d) and cnt[s] == e_cnt
        else:
            assert np.array_equal(ids, base[0]) and np.array_equal(d, base[1]) and np.array_equal(cnt, base[2])
    ix.close()


def test_large_flat_planted_neighbours_and_dump_cross_check(qadc):
    """2^26 vectors (512 MB of codes, several L2 sizes): (1) planted exact copies of the
    query's own best code are returned first with the smallest distance; (2) the scan result
    equals the canonical rule evaluated on the dump_distances output (an independent kernel and
    an independent selection in numpy)."""
    import torch
    rng = np.random.default_rng(2026)
    n, dim, m, nq, r = 1 << 26, 128, 16, 4, 100
    cb = synth.make_pq(rng, dim, m)
    q = synth.make_queries(rng, nq, dim)
    g = torch.Generator(device="cuda").manual_seed(5)
    d_codes = torch.randint(0, 256, (n, m // 2), dtype=torch.uint8, device="cuda", generator=g)
    # plant, for query 0, the code of its nearest centroids at a few known positions
    best = np.array([np.argmin(((q[0, j * 8:(j + 1) * 8][None] - cb[j]) ** 2).sum(1)) for j in range(m)])
    code = (best[0::2] | (best[1::2] << 4)).astype(np.uint8)
    planted = np.array([5, 123456, 33554431, n - 1], np.int64)
    d_codes[torch.from_numpy(planted).cuda()] = torch.from_numpy(code).cuda()
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb)
    ix.begin_database([n], False)
    ix.upload_codes_device(0, 0, n, d_codes.data_ptr())
    ix.finalize(0.001)
    del d_codes
    ids, d, cnt = ix.search(q, 1, r)
    assert np.array_equal(np.sort(ids[0][:4]), planted.astype(np.uint32)) and np.all(d[0][:4] == d[0][0])
    assert np.all(d[0][4:] >= d[0][0])
    out = ix.build_tables(q, 1, r)
    for s in range(nq):
        dist = ix.dump_distances(0, out["qtables"][s, 0])
        thr = d[s][cnt[s] - 1] if cnt[s] == r else 126
        pos = np.nonzero(dist <= thr)[0]
        order = np.lexsort((pos, dist[pos]))[:r]
        assert np.array_equal(ids[s][:len(order)], pos[order].astype(np.uint32))
        assert np.array_equal(d[s][:len(order)], dist[pos[order]])
    ix.close()


@pytest.mark.parametrize("G", [2, 4])
def test_virtual_ivf_shards_merge_equals_single(qadc, oracle, G):
    """Sharded inverted lists: every shard owns whole lists (greedy by size), holds a
    replica of ALL keep-prefixes, and returns its local top-r with labels resolved locally; the
    merged result equals the unsharded one (and the oracle)."""
    import torch
    from qadc_b200 import sharding
    rng = np.random.default_rng(40 + G)
    dim, m, n, K, ma, nq, r, keep = 96, 16, 60000, 48, 10, 11, 50, 0.05
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(9,))
    q = synth.make_queries(rng, nq, dim)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=keep,
                             offsets=offsets), q, ma, r, want_tables=False)
    sizes = np.diff(offsets)
    owner = sharding.ivf_list_owner(sizes, G)
    dq = torch.from_numpy(q).cuda()
    keys = torch.empty((G, nq, r), dtype=torch.int64, device="cuda")
    ids = torch.empty((G, nq, r), dtype=torch.int32, device="cuda")
    for g in range(G):
        ix = qadc.Index(0)
        ix.set_pq(dim, m, cb)
        ix.set_coarse(cents)
        local = np.where(owner == g, sizes, 0).astype(np.uint32)
        ix.begin_database(local, True)
        for p in range(K):
            if sizes[p] == 0:
                continue
            if owner[p] == g:
                ix.upload_codes(p, 0, codes[offsets[p]:offsets[p + 1]], labels[offsets[p]:offsets[p + 1]])
            ix.set_prefix(p, codes[offsets[p]:offsets[p] + sharding.start_size(int(sizes[p]), keep)])
        ix.finalize(keep)
        d_tmp = torch.empty((nq, r), dtype=torch.int8, device="cuda")
        c_tmp = torch.empty(nq, dtype=torch.int32, device="cuda")
        ix.search_device(dq.data_ptr(), nq, ma, r, ids[g].data_ptr(), d_tmp.data_ptr(), c_tmp.data_ptr(), keys[g].data_ptr())
        ix.synchronize()
        ix.close()
    mi = qadc.Index(0)
    o_ids = torch.empty((nq, r), dtype=torch.int32, device="cuda")
    o_d = torch.empty((nq, r), dtype=torch.int8, device="cuda")
    o_c = torch.empty(nq, dtype=torch.int32, device="cuda")
    mi.merge_shards_device(keys.data_ptr(), ids.data_ptr(), G, nq, r, o_ids.data_ptr(), o_d.data_ptr(), o_c.data_ptr())
    mi.synchronize()
    assert np.array_equal(o_ids.cpu().numpy().view(np.uint32), exp["ids"])
    assert np.array_equal(o_d.cpu().numpy(), exp["d"])
    assert np.array_equal(o_c.cpu().numpy(), exp["count"])
    mi.close()


@pytest.mark.parametrize("G,keep,max_list", [(2, 0.05, None), (4, 0.05, None), (8, 0.3, None), (3, 0.9, 400)])
def test_virtual_ivf_shards_owner_computes_equals_single(qadc, oracle, G, keep, max_list):
    """"Owner computes" (qadc_tables_local_device -> all-gather -> qadc_search_bounded_device): every shard holds ONLY its
    own lists (no prefix replicas), builds tables for the probes it owns and contributes (min entry, r smallest prefix
    distances); bounds from the union, scan, shard merge == the unsharded search == the oracle, and qmin/qmax are
    bit-identical to the unsharded ones.  keep 0.3 / 0.9: prefixes longer than 128 vectors (the CTA-wide prefix kernel)."""
    import torch
    from qadc_b200 import sharding
    rng = np.random.default_rng(140 + G)
    dim, m, n, K, ma, nq, r = 96, 16, 60000, 48, 10, 11, 50
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(9,))
    if max_list:   # short lists so that keep 0.9 stays within the fused kernels' limits
        keepv = np.concatenate([np.arange(offsets[p], min(offsets[p + 1], offsets[p] + max_list)) for p in range(K)])
        sizes = np.minimum(np.diff(offsets), max_list)
        codes, labels = codes[keepv], labels[keepv]
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    q = synth.make_queries(rng, nq, dim)
    db = dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=keep, offsets=offsets)
    exp = oracle.search(db, q, ma, r, want_tables=True)
    sizes = np.diff(offsets)
    owner = sharding.ivf_list_owner(sizes, G)
    dq = torch.from_numpy(q).cuda()
    d_assign = torch.from_numpy(exp["assign"].astype(np.int32)).cuda()
    shards = []
    for g in range(G):
        ix = qadc.Index(0)
        ix.set_pq(dim, m, cb)
        ix.set_coarse(cents)
        ix.begin_database(np.where(owner == g, sizes, 0).astype(np.uint32), True)
        ix.set_owned_partitions(owner == g)   # the empty list has an owner too: its tables count for qmin
        for p in range(K):
            if sizes[p] and owner[p] == g:
                ix.upload_codes(p, 0, codes[offsets[p]:offsets[p + 1]], labels[offsets[p]:offsets[p + 1]])
            elif sizes[p] and G == 4:   # prefix replicas of the other shards' lists may be present: they must be ignored
                ix.set_prefix(p, codes[offsets[p]:offsets[p] + sharding.start_size(int(sizes[p]), keep)])
        ix.finalize(keep)
        shards.append(ix)
    local = torch.empty((G, nq, r + 1), dtype=torch.float32, device="cuda")
    for g, ix in enumerate(shards):
        ix.tables_local_device(dq.data_ptr(), d_assign.data_ptr(), nq, ma, r, local[g].data_ptr())
        ix.synchronize()
    loc = local.cpu().numpy()
    # the union's bounds are the unsharded bounds, bit for bit
    assert np.array_equal(np.minimum(loc[:, :, 0].min(0), np.float32(3.4e38)).astype(np.float32), exp["qmin"])
    allv = np.sort(np.concatenate([loc[g][:, 1:] for g in range(G)], axis=1), axis=1)
    assert np.array_equal(allv[:, r - 1], exp["qmax"])
    keys = torch.empty((G, nq, r), dtype=torch.int64, device="cuda")
    ids = torch.empty((G, nq, r), dtype=torch.int32, device="cuda")
    perm = torch.from_numpy(rng.permutation(G)).cuda()   # the gather order of the shards must not matter
    gathered = local[perm].contiguous()
    for g, ix in enumerate(shards):
        d_tmp = torch.empty((nq, r), dtype=torch.int8, device="cuda")
        c_tmp = torch.empty(nq, dtype=torch.int32, device="cuda")
        ix.search_bounded_device(gathered.data_ptr(), G, nq, ma, r, ids[g].data_ptr(), d_tmp.data_ptr(), c_tmp.data_ptr(),
                                 keys[g].data_ptr())
        ix.synchronize()
    with pytest.raises(qadc.QadcError):   # the second half without a first half of the same batch
        shards[0].search_bounded_device(gathered.data_ptr(), G, nq, ma, r, ids[0].data_ptr(), d_tmp.data_ptr(), c_tmp.data_ptr())
    for ix in shards:
        ix.close()
    mi = qadc.Index(0)
    o_ids = torch.empty((nq, r), dtype=torch.int32, device="cuda")
    o_d = torch.empty((nq, r), dtype=torch.int8, device="cuda")
    o_c = torch.empty(nq, dtype=torch.int32, device="cuda")
    mi.merge_shards_device(keys.data_ptr(), ids.data_ptr(), G, nq, r, o_ids.data_ptr(), o_d.data_ptr(), o_c.data_ptr())
    mi.synchronize()
    assert np.array_equal(o_ids.cpu().numpy().view(np.uint32), exp["ids"])
    assert np.array_equal(o_d.cpu().numpy(), exp["d"])
    assert np.array_equal(o_c.cpu().numpy(), exp["count"])
    mi.close()


@pytest.mark.parametrize("G,K,ma", [(2, 48, 10), (8, 1500, 64), (3, 5, 4)])
def test_sharded_coarse_assignment_equals_unsharded(qadc, oracle, G, K, ma):
    """Cells split into G contiguous ranges: per-range top-ma keys (qadc_coarse_partial_device),
    gathered [G][nq][ma], merged (qadc_coarse_merge_device) == the oracle's assignment; a search
    with that assignment (qadc_search_assigned_device) == the plain search."""
    import torch
    from qadc_b200 import sharding
    rng = np.random.default_rng(70 + G)
    dim, m, n, nq, r, keep = 64, 16, 30000, 13, 20, 0.2
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    cents[K // 2] = cents[0]                       # duplicate cell: distance tie resolved by index
    codes, labels, offsets = synth.make_ivf(rng, n, K, m)
    q = synth.make_queries(rng, nq, dim)
    ix = ivf_index(qadc, dim, m, cb, cents, codes, labels, offsets, keep)
    dq = torch.from_numpy(q).cuda()
    gathered = torch.empty((G, nq, ma), dtype=torch.int64, device="cuda")
    for g in range(G):
        first, count = sharding.coarse_range(K, g, G)
        ix.coarse_partial_device(dq.data_ptr(), nq, ma, first, count, gathered[g].data_ptr())
    assign = torch.empty((nq, ma), dtype=torch.int32, device="cuda")
    ix.coarse_merge_device(gathered.data_ptr(), G, nq, ma, assign.data_ptr())
    ix.synchronize()
    exp_assign, exp_dist = oracle.coarse_assign(q, cents, ma)
    assert np.array_equal(assign.cpu().numpy(), exp_assign)
    gk = gathered.cpu().numpy()
    for g in range(G):                              # every partial list: ascending, inside its range, ~0 padded
        first, count = sharding.coarse_range(K, g, G)
        real = gk[g] != -1
        assert np.all(real.sum(1) == min(ma, count))
        assert np.all(np.diff(gk[g].view(np.uint64), axis=1) >= 0)
        idx = (gk[g] & 0xffffffff)[real]
        assert np.all((idx >= first) & (idx < first + count))
    # search with the supplied assignment == search that computes it
    outs = []
    for supplied in (False, True):
        ids = torch.empty((nq, r), dtype=torch.int32, device="cuda")
        d = torch.empty((nq, r), dtype=torch.int8, device="cuda")
        cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
        if supplied:
            ix.search_assigned_device(dq.data_ptr(), assign.data_ptr(), nq, ma, r, ids.data_ptr(), d.data_ptr(), cnt.data_ptr())
        else:
            ix.search_device(dq.data_ptr(), nq, ma, r, ids.data_ptr(), d.data_ptr(), cnt.data_ptr())
        ix.synchronize()
        outs.append((ids.cpu().numpy(), d.cpu().numpy(), cnt.cpu().numpy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=keep,
                             offsets=offsets), q, ma, r, want_tables=False)
    assert np.array_equal(outs[1][0].view(np.uint32), exp["ids"]) and np.array_equal(outs[1][1], exp["d"])
    ix.close()


# ---- "next" row N1: PQ encoder on the GPU ----------------------------------------------------
@pytest.mark.parametrize("name", ["encode_m16", "encode_m32", "encode_8x8"])
def test_gpu_encoder_matches_reference_and_oracle(qadc, oracle, name):
    g = load(name)
    bits = int(g["bits"]) if "bits" in g else 4
    ix = qadc.Index(0)
    ix.set_pq(int(g["dim"]), int(g["m"]), g["codebooks"], bits=bits)
    codes = ix.encode(g["vectors"])
    assert np.array_equal(codes, g["ref_codes"])                                   # reference encoder output
    assert np.array_equal(codes, oracle.encode(g["vectors"], int(g["m"]), g["codebooks"], bits))
    ix.close()


def test_gpu_encode_8bit_ivf_then_adc_search(qadc, oracle):
    """8-bit quantiser end to end: coarse assignment + residual codes on the GPU (== oracle), loaded as
    a plain-ADC database, searched (== oracle)."""
    rng = np.random.default_rng(17)
    dim, m, bits, n, K, ma, nq, r = 64, 8, 8, 20000, 10, 3, 8, 25
    cb = (0.5 * rng.standard_normal((m, 256, dim // m))).astype(np.float32)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    base = (cents[rng.integers(0, K, n)] + rng.standard_normal((n, dim))).astype(np.float32)
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb, bits=bits)
    ix.set_coarse(cents)
    codes, assign = ix.encode(base)
    exp_assign, _ = oracle.coarse_assign(base, cents, 1)
    assert np.array_equal(assign, exp_assign[:, 0])
    assert np.array_equal(codes, oracle.encode((base - cents[assign]).astype(np.float32), m, cb, bits))
    order = np.argsort(assign, kind="stable")
    offsets = np.concatenate([[0], np.cumsum(np.bincount(assign, minlength=K))]).astype(np.int64)
    db = dict(dim=dim, m=m, bits=bits, codebooks=cb, centroids=cents, codes=codes[order], labels=order.astype(np.uint32),
              offsets=offsets)
    ix.adc_load(db["codes"], db["labels"], offsets)
    q = synth.make_queries(rng, nq, dim)
    ids, d, cnt = ix.adc_search(q, ma, r)
    exp = oracle.adc_search(db, q, ma, r)
    assert np.array_equal(ids, exp["ids"]) and np.array_equal(d.view(np.uint32), exp["d"].view(np.uint32))
    ix.close()


def test_gpu_encode_search_recall(qadc, oracle):
    """floats -> GPU encoder -> database -> search: queries that are small perturbations of database
    vectors find them (Recall@100 with t=1 as recall.hpp:45-54 defines it), OPQ rotation included."""
    rng = np.random.default_rng(9)
    dim, m, n, nq, r = 64, 16, 50000, 50, 100
    rot = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32)
    base = rng.standard_normal((n, dim)).astype(np.float32)
    cb = np.stack([rng.permutation(base @ rot.T)[:16, j * 4:(j + 1) * 4] for j in range(m)]).astype(np.float32)
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb, rotation=rot)
    codes = ix.encode(base)
    assert np.array_equal(codes, oracle.encode(oracle.rotate(base, rot), m, cb))
    ix.load_flat(codes, 0.02)
    truth = rng.integers(0, n, nq)
    q = base[truth] + 0.01 * rng.standard_normal((nq, dim)).astype(np.float32)
    ids, d, cnt = ix.search(q, 1, r)
    recall = np.mean([truth[i] in ids[i] for i in range(nq)])
    assert recall >= 0.9, recall
    ix.close()


def test_gpu_ivf_assign_encode(qadc, oracle):
    """index_db::add_vectors path: nearest coarse cell (k=1) + code of the residual."""
    rng = np.random.default_rng(10)
    dim, m, n, K = 128, 16, 5000, 300
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb)
    ix.set_coarse(cents)
    codes, assign = ix.encode(x)
    e_assign, _ = oracle.coarse_assign(x, cents, 1)
    assert np.array_equal(assign, e_assign[:, 0])
    resid = (x - cents[assign]).astype(np.float32)
    assert np.array_equal(codes, oracle.encode(resid, m, cb))
    ix.close()


@pytest.mark.parametrize("m,dim,K,ma,opq", [(16, 128, 300, 24, False), (32, 96, 64, 64, False), (16, 96, 500, 128, True),
                                            (16, 64, 40, 3, False)])
def test_ivf_fused_table_pipeline_equals_separate_kernels(qadc, oracle, m, dim, K, ma, opq):
    """ivf_prepare_kernel (tables, keep-prefix ADC, bounds, int8 tables and the shared-bound seed of a query in one
    kernel, tables resident in shared memory) against the separate kernels it replaces and against the oracle:
    assignment, float tables, qmin, qmax, int8 tables and search results bit for bit."""
    rng = np.random.default_rng(1000 + m + K)
    n, nq, r, keep = 60000, 33, 100, 0.08   # 3 probes, one of them possibly an empty list: still >= r prefix vectors
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(1, K - 1))
    q = synth.make_queries(rng, nq, dim)
    rot = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32) if opq else None
    ix = ivf_index(qadc, dim, m, cb, cents, codes, labels, offsets, keep, rot)
    outs = []
    for fused in (2, 0):   # 2: whenever the tables of a query fit shared memory (1 = only while two CTAs fit an SM)
        ix.set_option("ivf_fused", fused)
        t = ix.build_tables(q, ma, r)
        assert t["rc"] == 0
        outs.append((t, ix.search(q, ma, r)))
    (t1, s1), (t0, s0) = outs
    for k in ("assign", "tables", "qmin", "qmax", "qtables"):
        assert np.array_equal(t1[k], t0[k]), k
    for a, b in zip(s1, s0):
        assert np.array_equal(a, b)
    db = dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=keep, offsets=offsets)
    if opq:
        db["rotation"] = rot
    exp = oracle.search(db, q, ma, r)
    assert np.array_equal(t1["tables"], exp["tables"]) and np.array_equal(t1["qtables"], exp["qtables"])
    assert np.array_equal(s1[0], exp["ids"]) and np.array_equal(s1[1], exp["d"]) and np.array_equal(s1[2], exp["count"])
    ix.close()


@pytest.mark.parametrize("name", ["add_ivf_pq", "add_ivf_opq"])
def test_gpu_add_vectors_matches_reference_golden(qadc, oracle, name):
    """qadc_encode with a coarse quantiser (and an OPQ rotation: residual -> rotate -> encode) against what the
    reference's own index_db::add_vectors stored (tests/golden/add_ivf_*.npz) and against the oracle, bit for bit."""
    g = load(name)
    m, dim = int(g["m"]), int(g["dim"])
    rot = g["rotation"] if "rotation" in g else None
    ix = qadc.Index(0)
    ix.set_pq(dim, m, g["codebooks"], rotation=rot)
    ix.set_coarse(g["centroids"])
    codes, assign = ix.encode(g["vectors"])
    assert np.array_equal(assign, g["ref_assign"])
    resid = (g["vectors"] - g["centroids"][assign]).astype(np.float32)
    if rot is not None:
        resid = oracle.rotate(resid, rot)
    assert np.array_equal(codes, oracle.encode(resid, m, g["codebooks"]))
    diff = (codes != g["ref_codes"]).any(axis=1).mean()   # sgemm + -ffast-math on the reference side: see test_oracle
    assert diff == 0 if rot is None else diff <= 0.01, diff
    ix.close()


@pytest.mark.parametrize("r", [1, 2, 1024])
def test_extreme_r(qadc, oracle, r):
    """r = 1 and the largest supported r (1024), flat and IVF, against the oracle."""
    rng = np.random.default_rng(500 + r)
    dim, m, n, nq = 128, 16, 40000, 5
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    keep = 0.1
    ix = flat_index(qadc, dim, m, cb, codes, keep)
    ids, d, cnt = ix.search(q, 1, r)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=keep, offsets=np.array([0, n], np.int64)),
                        q, 1, r, want_tables=False)
    assert np.array_equal(ids, exp["ids"]) and np.array_equal(d, exp["d"]) and np.array_equal(cnt, exp["count"])
    ix.close()
    K, ma = 16, 6
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m)
    ix = ivf_index(qadc, dim, m, cb, cents, codes, labels, offsets, keep)
    ids, d, cnt = ix.search(q, ma, r)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=keep,
                             offsets=offsets), q, ma, r, want_tables=False)
    assert np.array_equal(ids, exp["ids"]) and np.array_equal(d, exp["d"]) and np.array_equal(cnt, exp["count"])
    with pytest.raises(qadc.QadcError):
        ix.search(q, ma, 1025)   # r > 1024 is refused
    ix.close()


def test_search_more_queries_than_one_sub_batch(qadc, oracle):
    """40 000 queries in one call (internally split into sub-batches of 32 768)."""
    rng = np.random.default_rng(4242)
    dim, m, n, nq, r = 128, 16, 20000, 40000, 10
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    ix = flat_index(qadc, dim, m, cb, codes, 0.05)
    ids, d, cnt = ix.search(q, 1, r)
    sel = np.array([0, 1, 32767, 32768, 32769, 39999])
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=0.05, offsets=np.array([0, n], np.int64)),
                        q[sel], 1, r, want_tables=False)
    assert np.array_equal(ids[sel], exp["ids"]) and np.array_equal(d[sel], exp["d"]) and np.array_equal(cnt[sel], exp["count"])
    ix.close()


def test_recall_at_100_matches_reference(qadc, oracle, ref):
    """Recall@100 (recall.hpp:45-54 with t = 1: the true nearest neighbour is among the returned
    ids) of the CUDA path vs the unmodified reference on the same encoded IVF database."""
    rng = np.random.default_rng(77)
    dim, m, n, K, ma, nq, r = 128, 16, 30000, 64, 8, 200, 100
    base = rng.standard_normal((n, dim)).astype(np.float32)
    cents = base[rng.permutation(n)[:K]].copy()
    # codebooks drawn from residual sub-vectors so that codes are meaningful
    assign, _ = oracle.coarse_assign(base, cents, 1)
    resid = (base - cents[assign[:, 0]]).astype(np.float32)
    cb = np.stack([resid[rng.permutation(n)[:16], j * 8:(j + 1) * 8] for j in range(m)]).astype(np.float32)
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb)
    ix.set_coarse(cents)
    codes, a = ix.encode(base)                                   # GPU encoder == reference encoder (golden test)
    order = np.argsort(a, kind="stable")
    offsets = np.zeros(K + 1, np.int64)
    offsets[1:] = np.cumsum(np.bincount(a, minlength=K))
    codes_s, labels = codes[order], order.astype(np.uint32)
    truth = rng.integers(0, n, nq)
    q = (base[truth] + 0.05 * rng.standard_normal((nq, dim))).astype(np.float32)
    gt = np.argmin(((q[:, None, :] - base[None, :, :]) ** 2).sum(-1), axis=1) if n * nq * dim < 2e9 else truth
    keep = 0.05
    ix.load_ivf(codes_s, labels, offsets, keep)
    ids, d, cnt = ix.search(q, ma, r)
    h = ref.ivf(dim, m, cb, cents, codes_s, labels, offsets)
    h.prepare(keep)
    res = h.search(q, ma, r, nthreads=4)
    h.close()
    rec_gpu = np.mean([gt[i] in ids[i][:cnt[i]] for i in range(nq)])
    rec_ref = np.mean([gt[i] in res["keys"][i][:res["sizes"][i]] for i in range(nq)])
    assert rec_gpu >= 0.5 and abs(rec_gpu - rec_ref) <= 0.02, (rec_gpu, rec_ref)
    ix.close()


def test_ivf_schedule_does_not_change_results(qadc, oracle):
    """The inverted-list scan hands superblock items to warps dynamically; any schedule must give
    the canonical result.  Long lists with many exact duplicates (distance ties at every rank),
    item sizes from one superblock to whole lists, repeated runs: always bit-equal to the oracle."""
    rng = np.random.default_rng(91)
    dim, m, n, K, ma, nq, r, keep = 64, 16, 300000, 16, 6, 25, 100, 0.01
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(3,))
    codes[n // 2:n // 2 + 40000] = codes[:40000]        # duplicates: ties between lists and inside lists
    q = synth.make_queries(rng, nq, dim)
    exp = oracle.search(dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=keep,
                             offsets=offsets), q, ma, r, want_tables=False)
    ix = ivf_index(qadc, dim, m, cb, cents, codes, labels, offsets, keep)
    for spi in (8, 1, 3, 64, 4096, 8):
        ix.set_option("ivf_sb_per_item", spi)
        for _ in range(2):
            ids, d, cnt = ix.search(q, ma, r)
            assert np.array_equal(cnt, exp["count"]) and np.array_equal(d, exp["d"]) and np.array_equal(ids, exp["ids"]), spi
    ix.close()


# ---- "next" row N4: the plain ADC scan of db_query on the GPU ----------------------------------
def adc_index(qadc, db):
    ix = qadc.Index(0)
    ix.set_pq(db["dim"], db["m"], db["codebooks"], rotation=db.get("rotation"), bits=db["bits"])
    if db.get("centroids") is not None:
        ix.set_coarse(db["centroids"])
        ix.adc_load(db["codes"], db["labels"], db["offsets"])
    else:
        ix.adc_load(db["codes"])
    return ix


@pytest.mark.parametrize("name", ADC)
def test_adc_search_golden(qadc, oracle, name):
    """GPU float ADC == oracle bit for bit (ids, float distances, counts) and within the stated
    1e-5 of the reference's db_query path on the golden inputs."""
    g = load(name)
    db = adc_db(g)
    ix = adc_index(qadc, db)
    ids, d, cnt = ix.adc_search(g["queries"], int(g["ma"]), int(g["r"]))
    exp = oracle.adc_search(db, g["queries"], int(g["ma"]), int(g["r"]))
    assert np.array_equal(cnt, exp["count"]) and np.array_equal(ids, exp["ids"])
    assert np.array_equal(d.view(np.uint32), exp["d"].view(np.uint32))
    check_adc_against_reference(ids, d, g)
    ix.close()


@pytest.mark.parametrize("m,bits,ivf", [(8, 8, False), (16, 8, False), (4, 8, True), (16, 4, False), (32, 4, True),
                                        (2, 16, False), (4, 16, True), (8, 16, False)])
def test_adc_search_larger(qadc, oracle, m, bits, ivf):
    """Splits, compaction rounds, duplicate codes (distance ties broken by scan order), r = 100; every (nsq, bits) pair
    of get_scan_func (query_common.hpp:122-147), the 16-bit ones with their 65 536-entry tables in global memory."""
    rng = np.random.default_rng(200 + m + bits)
    dim, n, nq, r = 8 * m, 150000, 9, 100
    db = dict(dim=dim, m=m, bits=bits, codebooks=rng.standard_normal((m, 1 << bits, dim // m)).astype(np.float32))
    ma = 1
    if ivf:
        K, ma = 20, 6
        sizes = rng.multinomial(n, np.ones(K) / K)
        sizes[4] = 0
        db.update(offsets=np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64),
                  centroids=(2 * rng.standard_normal((K, dim))).astype(np.float32),
                  labels=rng.permutation(n).astype(np.uint32))
    else:
        db["offsets"] = np.array([0, n], np.int64)
    codes = rng.integers(0, 256, (n, m * bits // 8), dtype=np.uint8)
    codes[n // 2:n // 2 + 3000] = codes[:3000]          # exact duplicates -> equal distances
    db["codes"] = codes
    q = synth.make_queries(rng, nq, dim)
    ix = adc_index(qadc, db)
    ids, d, cnt = ix.adc_search(q, ma, r)
    exp = oracle.adc_search(db, q, ma, r)
    assert np.array_equal(cnt, exp["count"]) and np.array_equal(ids, exp["ids"])
    assert np.array_equal(d.view(np.uint32), exp["d"].view(np.uint32))
    ix.close()


def test_adc_short_database_and_errors(qadc, oracle):
    rng = np.random.default_rng(3)
    db = dict(dim=32, m=4, bits=8, codebooks=rng.standard_normal((4, 256, 8)).astype(np.float32),
              codes=rng.integers(0, 256, (5, 4), dtype=np.uint8), offsets=np.array([0, 5], np.int64))
    q = synth.make_queries(rng, 2, 32)
    ix = qadc.Index(0)
    ix.set_pq(32, 4, db["codebooks"], bits=8)
    with pytest.raises(qadc.QadcError) as e:
        ix.adc_search(q, 1, 8)                           # nothing loaded yet
    assert e.value.code == qadc.QADC_ESTATE
    with pytest.raises(qadc.QadcError):
        ix.begin_database(np.array([5], np.uint32), False)   # Quick ADC needs 4-bit codes (db_query_4.cpp:393-402)
    ix.adc_load(db["codes"])
    ids, d, cnt = ix.adc_search(q, 1, 8)
    exp = oracle.adc_search(db, q, 1, 8)
    assert np.array_equal(cnt, exp["count"]) and np.array_equal(ids, exp["ids"]) and np.array_equal(d, exp["d"])
    with pytest.raises(qadc.QadcError):
        ix.adc_search(q, 2, 8)                           # flat database: ma must be 1
    ix.close()
    with pytest.raises(qadc.QadcError):
        qadc.Index(0).set_pq(24, 3, rng.standard_normal((3, 256, 8)).astype(np.float32), bits=8)   # (3,8): not a get_scan_func pair


def test_gpu_encoder_16bit_matches_oracle(qadc, oracle):
    """16-bit sub-quantisers: nearest of 65 536 centroids per sub-vector, codes as little-endian uint16
    (multiple_set_bits_native<std::uint16_t>, quantizers.hpp:36-47) == the oracle's encoder, then searched as a plain-ADC
    database == the oracle.  (The reference's own encoder goes through find_k_neighbors, whose stride is wrong for more
    than 256 centroids, neighbors.cpp:64 — no golden codes exist for this case.)"""
    rng = np.random.default_rng(1616)
    m, dim, n, nq, r = 4, 8, 3000, 6, 30
    cb = synth.lattice_codebook16(m)
    base = (3.0 * rng.standard_normal((n, dim))).astype(np.float32)
    ix = qadc.Index(0)
    ix.set_pq(dim, m, cb, bits=16)
    codes = ix.encode(base)
    assert codes.shape == (n, 2 * m) and np.array_equal(codes, oracle.encode(base, m, cb, 16))
    ix.adc_load(codes)
    q = (3.0 * synth.make_queries(rng, nq, dim)).astype(np.float32)
    ids, d, cnt = ix.adc_search(q, 1, r)
    exp = oracle.adc_search(dict(dim=dim, m=m, bits=16, codebooks=cb, codes=codes, offsets=np.array([0, n], np.int64)), q, 1, r)
    assert np.array_equal(cnt, exp["count"]) and np.array_equal(ids, exp["ids"]) and np.array_equal(d, exp["d"])
    ix.close()
