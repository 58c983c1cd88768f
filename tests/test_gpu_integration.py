"""The integration claim of INTEGRATION.md, executed: the UNMODIFIED reference translation unit (db_query_4.cpp, compiled
in place into oracle/_ref/libqadc_refint.so) drives its own process_queries loop (query_common.hpp:330-368) once with
its own nns_engine_batch<scanner_4> and once with the adaptor a maintainer would add
(quick-adc_b200/host/reference_adaptor.hpp: scanner_gpu_4 + nns_engine_gpu over libqadc_b200.so).  Both runs see
the same in-memory flat_db / index_db, the same query and ground-truth files, the same heap type; a recording
wrapper copies every query's heap out.  Comparison = SURVEY §8c Stage R (tie classes) + recall."""
import ctypes as C
import os

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libqadc_refint.so")


def _run(db, q_path, gt_path, nq, r, ma, batch, use_gpu, devices=(0,)):
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libqadc_refint.so not built (needs /root/reference)")
    lib = C.CDLL(LIB)
    ivf = "centroids" in db
    K = len(db["offsets"]) - 1 if ivf else 0
    keys = np.zeros((nq, r), np.uint32); vals = np.full((nq, r), 127, np.int8); sizes = np.zeros(nq, np.int32)
    recall = C.c_double(); metrics = np.zeros(4, np.float64)
    cb = np.ascontiguousarray(db["codebooks"], np.float32).reshape(-1)
    codes = np.ascontiguousarray(db["codes"], np.uint8)
    off = np.ascontiguousarray(db["offsets"], np.int64)
    dv = np.ascontiguousarray(devices, np.int32)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rot = db.get("rotation")
    rc = lib.refint_run(C.c_int(db["dim"]), C.c_int(db["m"]), p(cb), p(None if rot is None else np.ascontiguousarray(rot, np.float32)),
                        C.c_int(K), p(db.get("centroids")), p(codes), p(db.get("labels")), p(off), C.c_float(db["keep"]),
                        str(q_path).encode(), str(gt_path).encode(), C.c_int(r), C.c_int(ma), C.c_int(batch), C.c_int(use_gpu),
                        p(dv), C.c_int(len(dv)), p(keys), p(vals), p(sizes), C.byref(recall), p(metrics))
    assert rc == 0
    return keys, vals, sizes, recall.value, metrics


def _files(tmp_path, q, gt):
    from qadc_b200 import dbfile
    dbfile.write_vecs(tmp_path / "q.fvecs", q)
    dbfile.write_vecs(tmp_path / "gt.ivecs", gt)
    return tmp_path / "q.fvecs", tmp_path / "gt.ivecs"


def _flat_db(rng, n, dim, m, keep):
    return dict(dim=dim, m=m, codebooks=synth.make_pq(rng, dim, m), codes=synth.make_codes(rng, n, m), keep=keep,
                offsets=np.array([0, n], np.int64))


def _ivf_db(rng, n, dim, m, K, keep, opq=False):
    codes, labels, offsets = synth.make_ivf(rng, n, K, m)
    db = dict(dim=dim, m=m, codebooks=synth.make_pq(rng, dim, m), centroids=(2 * rng.standard_normal((K, dim))).astype(np.float32),
              codes=codes, labels=labels, offsets=offsets, keep=keep)
    if opq:
        db["rotation"] = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32)
    return db


def test_reference_engine_through_the_harness_equals_the_ref_library(ref, tmp_path):
    """CPU: the harness's reference leg (process_queries + nns_engine_batch<scanner_4>) returns the heaps the
    round-1 reference harness (oracle/_ref/libqadc_ref.so, nns_engine loop) returns with sgemm-form tables."""
    rng = np.random.default_rng(11)
    db = _flat_db(rng, 20000, 128, 16, 0.02)
    nq, r = 12, 50
    q = synth.make_queries(rng, nq, 128)
    qp, gp = _files(tmp_path, q, np.zeros((nq, 1), np.int32))
    keys, vals, sizes, recall, _ = _run(db, qp, gp, nq, r, 1, 5, 0)
    h = ref.flat(128, 16, db["codebooks"], db["codes"])
    h.prepare(db["keep"])
    exp = h.search(q, 1, r, blas_tables=True)
    h.close()
    assert np.array_equal(sizes, exp["sizes"]) and np.array_equal(keys, exp["keys"]) and np.array_equal(vals, exp["vals"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["flat16", "flat32", "ivf", "ivf_opq"])
def test_reference_process_queries_with_the_gpu_engine(qadc, oracle, tmp_path, kind):
    rng = np.random.default_rng({"flat16": 1, "flat32": 2, "ivf": 3, "ivf_opq": 4}[kind])
    if kind == "flat16":
        db, ma = _flat_db(rng, 120000, 128, 16, 0.01), 1
    elif kind == "flat32":
        db, ma = _flat_db(rng, 60000, 128, 32, 0.01), 1   # sq_dim 4: the reference's dispatchers exit(1) on the 96-d / sq_dim 3 shape
    else:
        db, ma = _ivf_db(rng, 80000, 128, 16, 200, 0.05, opq=(kind == "ivf_opq")), 16   # K <= 256: the reference's assignment is right
    nq, r, batch = 48, 100, 20
    q = synth.make_queries(rng, nq, db["dim"])
    # ground truth = the canonical best neighbour of half of the queries (recall must then be exactly 0.5 on both sides)
    exp = oracle.search(db, q, ma, r, want_tables=True)
    gt = exp["ids"][:, :1].astype(np.int32).copy()
    gt[::2] = db["codes"].shape[0] + 7
    qp, gp = _files(tmp_path, q, gt)
    rk, rv, rs, r_recall, _ = _run(db, qp, gp, nq, r, ma, batch, 0)
    gk, gv, gs, g_recall, g_met = _run(db, qp, gp, nq, r, ma, batch, 1)
    assert np.array_equal(gs, rs)                                    # same heap fill
    # the GPU engine hands the reference's loop exactly the canonical result
    for i in range(nq):
        order = np.lexsort((gk[i], gv[i]))
        assert np.array_equal(np.sort(gv[i]), np.sort(np.concatenate([exp["d"][i][:exp["count"][i]], np.full(r - exp["count"][i], 127, np.int8)])))
        assert set(gk[i][gv[i] < 127].tolist()) == set(exp["ids"][i][:exp["count"][i]].tolist())
    # Stage R against the reference's own heaps: identical distance multisets and identical ids below the r-th
    # distance wherever the two sides quantised the tables identically (the reference is a -ffast-math build with
    # sgemm-form tables: an entry may differ by one LSB) and the reference heap holds no duplicate ids (F5)
    n_dup = n_exact = 0
    for i in range(nq):
        real = rv[i] < 127
        if len(np.unique(rk[i][real])) != real.sum():
            n_dup += 1                      # pad lanes re-pushed by the reference (F5b): its heap is not a set
            continue
        sg, sr = np.sort(gv[i]).astype(int), np.sort(rv[i]).astype(int)
        assert np.abs(sg - sr).max() <= 2, (i, sg, sr)     # never more than table-LSB noise apart
        if not np.array_equal(sg, sr):
            continue
        n_exact += 1
        vstar = int(sg[-1]) if gs[i] == r else 127
        assert set(gk[i][gv[i] < vstar].tolist()) == set(rk[i][rv[i] < vstar].tolist())
    print(f"{kind}: {n_exact} of {nq} queries identical to the reference heap (multiset + ids below the r-th distance), "
          f"{n_dup} excluded for duplicate ids in the reference heap, {nq - n_exact - n_dup} within table-LSB noise")
    assert n_exact >= (0.9 if ma == 1 else 0.5) * (nq - n_dup) and n_dup <= nq // 2, (n_exact, n_dup)
    assert abs(g_recall - 0.5) < 1e-9 and abs(r_recall - g_recall) <= 1.0 / nq + 1e-9
    assert g_met[3] > 0                                              # scan_us came back through query_metrics


@pytest.mark.gpu
def test_reference_process_queries_sharded(qadc, tmp_path):
    """Same loop, the adaptor given several devices (-g 0,1,...): byte-identical heaps to the one-device run."""
    import torch
    n_dev = torch.cuda.device_count()
    devices = tuple(range(n_dev)) if n_dev > 1 else (0, 0)
    rng = np.random.default_rng(9)
    db = _ivf_db(rng, 50000, 96, 16, 120, 0.05)
    nq, r = 30, 100
    q = synth.make_queries(rng, nq, 96)
    qp, gp = _files(tmp_path, q, np.zeros((nq, 1), np.int32))
    one = _run(db, qp, gp, nq, r, 8, 16, 1)
    many = _run(db, qp, gp, nq, r, 8, 16, 1, devices)
    assert np.array_equal(one[0], many[0]) and np.array_equal(one[1], many[1]) and np.array_equal(one[2], many[2])
