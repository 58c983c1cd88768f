"""CPU tests: the oracle (oracle/qadc_oracle.c) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), and against the live reference build
when oracle/_ref/libqadc_ref.so is present.  No GPU needed."""
import os

import numpy as np
import pytest

import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLAT = ["flat_m16", "flat_m32"]
IVF = ["ivf_m16", "ivf_m32"]
OPQ = ["ivf_opq_m16"]   # inverted lists + OPQ rotation (quantizers.hpp:286-301)
FLOAT_RTOL = 1e-5   # north_star: float tables within 1e-5 relative before quantisation


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def golden_db(g):
    """Oracle database dict (oracle.search) of a golden fixture."""
    ivf = "centroids" in g
    db = dict(dim=int(g["dim"]), m=int(g["m"]), codebooks=g["codebooks"], codes=g["codes"], keep=float(g["keep"]),
              offsets=g["offsets"] if ivf else np.array([0, g["codes"].shape[0]], np.int64))
    if ivf:
        db.update(centroids=g["centroids"], labels=g["labels"])
    if "rotation" in g:
        db["rotation"] = g["rotation"]
    return db


def rotated(g, vecs):
    """X * R^T in float64 (the reference: sgemm NoTrans/Trans), identity without OPQ."""
    if "rotation" not in g:
        return vecs
    return (vecs.astype(np.float64) @ g["rotation"].astype(np.float64).T).astype(np.float32)


def rel_err(a, b, floor):
    """max |a-b| / max(|b|, floor)"""
    return float(np.max(np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), floor)))


def blas_scale(vecs, codebooks, m):
    """||x_j||^2 + ||c_jc||^2 per table entry, shape (count, m, 16)."""
    x = vecs.reshape(vecs.shape[0], m, -1).astype(np.float64)
    return (x ** 2).sum(-1)[:, :, None] + (codebooks.astype(np.float64) ** 2).sum(-1)[None]


@pytest.mark.parametrize("name", FLAT)
def test_layout_matches_reference(oracle, name):
    g = load(name)
    assert np.array_equal(oracle.interleave(g["codes"]), g["ref_interleaved"])   # simd_layout.hpp:55-65


@pytest.mark.parametrize("name", FLAT)
def test_float_tables_within_tolerance(oracle, name):
    g = load(name)
    m = int(g["m"])
    mine = oracle.tables_direct(g["queries"], m, g["codebooks"])
    # vs compute_dists_single_simd_cg (the builder `db_query_4 -b1 -m1` uses)
    assert rel_err(mine, g["ref_tables_direct"], 1e-30) <= FLOAT_RTOL
    # vs the BLAS form (||x||^2 + ||c||^2 - 2 x.c): its rounding error scales with the terms
    # being cancelled, so the 1e-5 is taken relative to max(|entry|, ||x_j||^2 + ||c||^2)
    floor = blas_scale(g["queries"], g["codebooks"], m)
    assert rel_err(mine, g["ref_tables_blas"], floor) <= FLOAT_RTOL
    assert rel_err(oracle.tables_blasform(g["queries"], m, g["codebooks"]), g["ref_tables_blas"], floor) <= FLOAT_RTOL


@pytest.mark.parametrize("name", FLAT)
def test_bounds_and_quantiser(oracle, name):
    g = load(name)
    r, m = int(g["r"]), int(g["m"])
    ss = oracle.start_size(g["codes"].shape[0], g["keep"])
    assert ss == int(g["ref_start_size"][0])                                    # db_query_4.cpp:125-126
    mism = total = 0
    for q in range(g["queries"].shape[0]):
        t = g["ref_tables_used"][q, 0]
        qmax = oracle.prefix_qmax([g["codes"][:ss]], [t], r)                    # query_scan_start
        assert abs(qmax - g["ref_qmax"][q]) <= FLOAT_RTOL * g["ref_qmax"][q]
        assert abs(max(float(t.min()), 0.0) - g["ref_qmin"][q]) <= 1e-6 * max(1.0, g["ref_qmin"][q])
        # same bounds in -> QuantizerMAX out; -ffast-math may move an entry by one LSB
        qt = oracle.quantize(t, g["ref_qmin"][q], g["ref_qmax"][q])
        diff = qt.astype(int) - g["ref_qtables"][q, 0].astype(int)
        assert np.abs(diff).max() <= 1
        mism += int((diff != 0).sum())
        total += diff.size
    assert mism <= 0.01 * total, f"int8 table LSB mismatch rate {mism}/{total}"


@pytest.mark.parametrize("name", FLAT)
def test_distances_bit_exact(oracle, name):
    g = load(name)
    m, n = int(g["m"]), g["codes"].shape[0]
    for q in range(g["queries"].shape[0]):
        qt = g["ref_qtables"][q, 0]
        d = oracle.distances(g["codes"], qt)                                    # min(127, sum)
        assert np.array_equal(d, g["ref_distances"][q])
        lit = oracle.distances_interleaved(g["ref_interleaved"], n, m, qt)[:n]  # literal vpaddsb order
        assert np.array_equal(lit, d)


@pytest.mark.parametrize("name", FLAT)
def test_reference_heap_emulation_exact(oracle, name):
    """scan_avx_4 + kv_binheap restated: raw heap arrays identical to the reference's."""
    g = load(name)
    m, n, r = int(g["m"]), g["codes"].shape[0], int(g["r"])
    for q in range(g["queries"].shape[0]):
        keys, vals, hs = oracle.scan_ref_heap(g["ref_interleaved"], None, n, m, g["ref_qtables"][q, 0], r)
        assert hs == g["ref_heap_sizes"][q]
        assert np.array_equal(keys[:hs], g["ref_heap_keys"][q, :hs])
        assert np.array_equal(vals[:hs], g["ref_heap_vals"][q, :hs])


def tie_class_check(ids, d, cnt, ref_keys, ref_vals, all_d_parts, labels_parts):
    """Stage R: canonical result vs the raw reference heap.
    Returns False if the query is excluded (reference heap holds duplicate ids, F5b)."""
    r = len(ref_keys)
    real = ref_vals < 127
    if len(np.unique(ref_keys[real])) != real.sum():
        return False
    assert np.array_equal(np.sort(d), np.sort(ref_vals)), "distance multisets differ"
    vstar = int(d[cnt - 1]) if cnt == r else 127
    mine_lt = set(ids[:cnt][d[:cnt] < vstar].tolist())
    ref_lt = set(ref_keys[ref_vals < vstar].tolist())
    assert mine_lt == ref_lt, "ids below the r-th distance differ"
    # ids at d == v*: both sides must come from the tie class
    tie = set()
    for dd, lab in zip(all_d_parts, labels_parts):
        pos = np.nonzero(dd == vstar)[0]
        tie |= set((pos if lab is None else lab[pos]).tolist())
    assert set(ids[:cnt][d[:cnt] == vstar].tolist()) <= tie
    assert set(ref_keys[(ref_vals == vstar) & real].tolist()) <= tie
    return True


@pytest.mark.parametrize("name", FLAT)
def test_canonical_rule_vs_reference_heap_flat(oracle, name):
    g = load(name)
    r = int(g["r"])
    offsets = np.array([0, g["codes"].shape[0]], np.int64)
    checked = 0
    for q in range(g["queries"].shape[0]):
        qt = g["ref_qtables"][q]
        ids, d, cnt, _ = oracle.scan_with_tables(g["codes"], None, offsets, np.zeros(1, np.int32), qt, r)
        dist = g["ref_distances"][q]
        e_ids, e_d, e_cnt = synth.canonical_from_distances([dist], [None], r)   # rule evaluated by numpy
        assert cnt == e_cnt and np.array_equal(ids, e_ids) and np.array_equal(d, e_d)
        checked += tie_class_check(ids, d, cnt, g["ref_heap_keys"][q], g["ref_heap_vals"][q], [dist], [None])
    assert checked >= 1


@pytest.mark.parametrize("name", IVF + OPQ)
def test_ivf_against_reference(oracle, name):
    g = load(name)
    r, m, ma = int(g["r"]), int(g["m"]), int(g["ma"])
    codes, labels, offsets = g["codes"], g["labels"], g["offsets"]
    K = len(offsets) - 1
    for p in range(K):
        assert oracle.start_size(int(offsets[p + 1] - offsets[p]), g["keep"]) == int(g["ref_start_size"][p])
    # fixed coarse assignment == find_k_neighbors as shipped when K <= 256
    assign, _ = oracle.coarse_assign(g["queries"], g["centroids"], ma)
    assert np.array_equal(assign, g["ref_assign"])
    assert np.array_equal(assign, g["ref_assign_fkn"])
    checked = 0
    for q in range(g["queries"].shape[0]):
        a = g["ref_assign"][q]
        # residual tables (direct form) vs the reference's blas-form tables
        resid = rotated(g, g["queries"][q][None, :] - g["centroids"][a])
        mine = oracle.tables_direct(resid, m, g["codebooks"])
        assert rel_err(mine, g["ref_tables_used"][q], blas_scale(resid, g["codebooks"], m)) <= FLOAT_RTOL
        # bounds from the reference's tables
        t = g["ref_tables_used"][q]
        prefixes = [codes[offsets[p]:offsets[p] + g["ref_start_size"][p]] for p in a]
        qmax = oracle.prefix_qmax(prefixes, list(t), r)
        assert abs(qmax - g["ref_qmax"][q]) <= FLOAT_RTOL * g["ref_qmax"][q]
        qt = oracle.quantize(t, g["ref_qmin"][q], g["ref_qmax"][q])
        assert np.abs(qt.astype(int) - g["ref_qtables"][q].astype(int)).max() <= 1
        # reference heap emulated over the probed lists with the reference's int8 tables
        heap = None
        d_parts, l_parts = [], []
        for rank, p in enumerate(a):
            pc, pl = codes[offsets[p]:offsets[p + 1]], labels[offsets[p]:offsets[p + 1]]
            d_parts.append(oracle.distances(pc, g["ref_qtables"][q, rank]) if len(pc) else np.zeros(0, np.int8))
            l_parts.append(pl)
            if len(pc) == 0:
                continue
            heap = oracle.scan_ref_heap(oracle.interleave(pc), np.ascontiguousarray(pl), len(pc), m,
                                        g["ref_qtables"][q, rank], r, heap)
        keys, vals, hs = heap
        assert hs == g["ref_heap_sizes"][q]
        assert np.array_equal(keys[:hs], g["ref_heap_keys"][q, :hs])
        assert np.array_equal(vals[:hs], g["ref_heap_vals"][q, :hs])
        # canonical rule
        ids, d, cnt, _ = oracle.scan_with_tables(codes, labels, offsets, a, g["ref_qtables"][q], r)
        e_ids, e_d, e_cnt = synth.canonical_from_distances(d_parts, l_parts, r)
        assert cnt == e_cnt and np.array_equal(ids, e_ids) and np.array_equal(d, e_d)
        checked += tie_class_check(ids, d, cnt, g["ref_heap_keys"][q], g["ref_heap_vals"][q], d_parts, l_parts)
    assert checked >= 1


@pytest.mark.parametrize("name", FLAT + IVF + OPQ)
def test_full_pipeline_self_consistent(oracle, name):
    """qo_search (the whole canonical pipeline) agrees with its own stages and stays within
    tolerance of the reference's float results."""
    g = load(name)
    r, m = int(g["r"]), int(g["m"])
    ivf = "centroids" in g
    ma = int(g["ma"]) if ivf else 1
    db = golden_db(g)
    res = oracle.search(db, g["queries"], ma, r)
    assert res["rc"] == 0
    assert np.all(np.abs(res["qmax"] - g["ref_qmax"]) <= FLOAT_RTOL * g["ref_qmax"])
    lsb = np.abs(res["qtables"].astype(int) - g["ref_qtables"].astype(int))
    assert lsb.max() <= 1 and (lsb != 0).mean() <= 0.02
    for q in range(g["queries"].shape[0]):
        ids, d, cnt, _ = oracle.scan_with_tables(g["codes"], g["labels"] if ivf else None, db["offsets"],
                                                 res["assign"][q], res["qtables"][q], r)
        assert np.array_equal(ids, res["ids"][q]) and np.array_equal(d, res["d"][q]) and cnt == res["count"][q]


# db_query (plain ADC): 8-bit flat / IVF, 4-bit IVF + OPQ, 16-bit flat / IVF (lattice codebooks, synth.lattice_codebook16)
ADC = ["adc_flat_8x8", "adc_ivf_16x8", "adc_ivf_opq_32x4", "adc_flat_2x16", "adc_ivf_8x16"]


def adc_db(g):
    cb = g["codebooks"] if "codebooks" in g else synth.lattice_codebook16(int(g["m"]))
    db = dict(dim=int(g["dim"]), m=int(g["m"]), bits=int(g["bits"]), codebooks=cb, codes=g["codes"],
              offsets=g["offsets"])
    if "centroids" in g:
        db.update(centroids=g["centroids"], labels=g["labels"])
    if "rotation" in g:
        db["rotation"] = g["rotation"]
    return db


def check_adc_against_reference(ids, d, g):
    """Float ADC results vs the reference's sorted heap (scanner_simple): distances within 1e-5
    relative (the -ffast-math build may reassociate the per-vector sum), same ids wherever the
    r-th and (r+1)-th candidates are not closer than that."""
    assert np.max(np.abs(d - g["ref_dists"]) / np.maximum(np.abs(g["ref_dists"]), 1e-30)) <= FLOAT_RTOL
    for q in range(d.shape[0]):
        assert set(ids[q].tolist()) == set(g["ref_ids"][q].tolist())


@pytest.mark.parametrize("name", ADC)
def test_adc_oracle_vs_reference_golden(oracle, name):
    g = load(name)
    res = oracle.adc_search(adc_db(g), g["queries"], int(g["ma"]), int(g["r"]))
    assert np.all(res["count"] == int(g["r"]))
    assert np.all(np.diff(res["d"], axis=1) >= 0)
    check_adc_against_reference(res["ids"], res["d"], g)


def test_adc_oracle_short_database_pads_like_the_reference(oracle):
    """Fewer vectors than r: the tail is (0, FLT_MAX), the reference's pre-filled heap slots."""
    rng = np.random.default_rng(3)
    db = dict(dim=32, m=4, bits=8, codebooks=rng.standard_normal((4, 256, 8)).astype(np.float32),
              codes=rng.integers(0, 256, (5, 4), dtype=np.uint8), offsets=np.array([0, 5], np.int64))
    res = oracle.adc_search(db, synth.make_queries(rng, 2, 32), 1, 8)
    assert np.all(res["count"] == 5) and np.all(res["ids"][:, 5:] == 0)
    assert np.all(res["d"][:, 5:] == np.finfo(np.float32).max)
    assert all(sorted(res["ids"][q, :5].tolist()) == [0, 1, 2, 3, 4] for q in range(2))


def test_prefix_too_small_is_reported(oracle):
    """Fewer than r prefix vectors -> qmax = FLT_MAX -> the reference exits (db_query_4.cpp:271-274)."""
    rng = np.random.default_rng(5)
    cb = synth.make_pq(rng, 128, 16)
    codes = synth.make_codes(rng, 500, 16)
    db = dict(dim=128, m=16, codebooks=cb, codes=codes, keep=0.01, offsets=np.array([0, 500], np.int64))
    res = oracle.search(db, synth.make_queries(rng, 2, 128), 1, 10)
    assert res["rc"] == 1 and np.all(res["count"] == -1)


def test_heap_restatement_properties(oracle):
    """kv_binheap (binheap.hpp:75-116): keeps the k smallest values, root = max."""
    import ctypes as C
    rng = np.random.default_rng(7)
    for cap in (1, 2, 7, 100):
        keys, vals, hs = np.zeros(cap, np.uint32), np.zeros(cap, np.int8), C.c_int(0)
        stream = rng.integers(0, 127, 1000).astype(np.int8)
        for i, v in enumerate(stream):
            oracle.lib.qo_heap_push_i8(keys, vals, cap, C.byref(hs), i, int(v))
        assert hs.value == cap
        assert np.array_equal(np.sort(vals), np.sort(stream)[:cap])
        assert vals[0] == vals.max()
        assert np.array_equal(stream[keys], vals)


# ---- live cross-checks against the compiled reference (skipped where it is not built) ------
@pytest.mark.parametrize("n,m,dim", [(1, 16, 128), (15, 16, 128), (16, 32, 256), (17, 16, 96), (1000, 32, 96),
                                     (4099, 16, 128)])
def test_live_reference_random_shapes(oracle, ref, n, m, dim):
    rng = np.random.default_rng(n * 31 + m)
    codes = synth.make_codes(rng, n, m)
    inter = ref.interleave(codes)
    assert np.array_equal(oracle.interleave(codes), inter)
    for _ in range(3):
        qt = synth.make_qtables(rng, (), m)
        assert np.array_equal(oracle.distances(codes, qt), ref.dump_distances(inter, n, m, qt))
        r = int(rng.integers(1, 40))
        k1, v1, n1 = oracle.scan_ref_heap(inter, None, n, m, qt, r)
        k2, v2, n2 = ref.scan_avx_4(inter, None, n, m, qt, r)
        assert n1 == n2 and np.array_equal(k1[:n1], k2[:n2]) and np.array_equal(v1[:n1], v2[:n2])
    cb = synth.make_pq(rng, dim, m)
    q = synth.make_queries(rng, 4, dim)
    assert rel_err(oracle.tables_direct(q, m, cb), ref.tables(q, m, cb, False), 1e-30) <= FLOAT_RTOL


def test_live_reference_coarse_bug_documented(oracle, ref):
    """with more than 256 centroids the shipped find_k_neighbors strides wrongly;
    the oracle's fixed statement equals brute force, the reference does not."""
    rng = np.random.default_rng(11)
    dim, K, nq = 32, 600, 40
    cents = rng.standard_normal((K, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    brute = np.argsort(((q[:, None, :] - cents[None]) ** 2).sum(-1), axis=1, kind="stable")[:, :4]
    mine, _ = oracle.coarse_assign(q, cents, 4)
    assert np.array_equal(mine, brute)
    assert not np.array_equal(ref.find_k_neighbors(q, cents, 4), brute)
    small = cents[:200]
    brute_s = np.argsort(((q[:, None, :] - small[None]) ** 2).sum(-1), axis=1, kind="stable")[:, :4]
    assert np.array_equal(ref.find_k_neighbors(q, small, 4), brute_s)


@pytest.mark.parametrize("name", ["encode_m16", "encode_m32", "encode_8x8"])
def test_encoder_matches_reference(oracle, name):
    """PQ encoder (quantizers.hpp:222-245, :36-68): same codes as the reference's encoder."""
    g = load(name)
    bits = int(g["bits"]) if "bits" in g else 4
    mine = oracle.encode(g["vectors"], int(g["m"]), g["codebooks"], bits)
    assert np.array_equal(mine, g["ref_codes"])
    # code format: 4-bit: byte b = idx[2b] | idx[2b+1] << 4; 8-bit: byte b = idx[b]
    x, cb, m = g["vectors"], g["codebooks"], int(g["m"])
    dsq = x.shape[1] // m
    idx = np.stack([np.argmin(((x[:, None, j * dsq:(j + 1) * dsq] - cb[j][None]) ** 2).sum(-1), axis=1) for j in range(m)], 1)
    packed = (idx[:, 0::2] | (idx[:, 1::2] << 4)).astype(np.uint8) if bits == 4 else idx.astype(np.uint8)
    assert (packed != mine).mean() < 1e-3   # float64 argmin vs float32 direct form: only near-ties may differ


@pytest.mark.parametrize("name", ["add_ivf_pq", "add_ivf_opq"])
def test_add_vectors_restatement_matches_reference(oracle, name):
    """index_db::add_vectors (databases.hpp:270-298): nearest cell, residual, [OPQ: rotate], encode — the oracle's
    stages chained the same way reproduce the reference's cells and codes (golden = the reference's own add_vectors)."""
    g = load(name)
    m = int(g["m"])
    assign, _ = oracle.coarse_assign(g["vectors"], g["centroids"], 1)
    assert np.array_equal(assign[:, 0], g["ref_assign"])
    resid = (g["vectors"] - g["centroids"][assign[:, 0]]).astype(np.float32)
    if "rotation" in g:
        resid = oracle.rotate(resid, g["rotation"])
    codes = oracle.encode(resid, m, g["codebooks"])
    # the reference rotates with sgemm (-ffast-math): a residual component within float noise of a cell boundary may
    # land in the neighbouring centroid; bound the rate instead of demanding equality for the OPQ case
    diff = (codes != g["ref_codes"]).any(axis=1).mean()
    assert diff == 0 if "rotation" not in g else diff <= 0.01, diff
