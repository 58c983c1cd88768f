"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libqadc_ref.so,
built from /root/reference by oracle/Makefile).  Run in the CPU container:

    python tests/golden/make_golden.py

The reference ships no tests or fixtures, so these files are what pins the
oracle and the CUDA path to the reference's behaviour on the GPU box, where
/root/reference does not exist.  Every array below that starts with ``ref_`` was produced
by reference code; the rest are the seeded inputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.pyoracle import Ref, RefAdc  # noqa: E402
import synth  # noqa: E402


def flat_case(ref, name, seed, dim, m, n, nq, r, keep):
    rng = np.random.default_rng(seed)
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    queries = synth.make_queries(rng, nq, dim)
    out = dict(dim=dim, m=m, r=r, keep=np.float32(keep), codebooks=cb, codes=codes, queries=queries)
    out["ref_interleaved"] = ref.interleave(codes)
    out["ref_tables_direct"] = ref.tables(queries, m, cb, False)        # compute_dists_single_simd_cg
    out["ref_tables_blas"] = ref.tables(queries, m, cb, True)           # compute_dists_multiple_blas_cg
    h = ref.flat(dim, m, cb, codes)
    h.prepare(keep)
    out["ref_start_size"] = np.array([h.starts_size(0)], np.uint32)
    res = h.search(queries, 1, r, nthreads=1, blas_tables=False, want_tables=True)   # what `db_query_4 -b1` runs
    out["ref_heap_keys"], out["ref_heap_vals"], out["ref_heap_sizes"] = res["keys"], res["vals"], res["sizes"]
    out["ref_tables_used"] = res["tables"]                                # after the in-place clamp
    qmin = np.zeros(nq, np.float32)
    qmax = np.zeros(nq, np.float32)
    qt = np.zeros((nq, 1, m, 16), np.int8)
    dist = np.zeros((nq, n), np.int8)
    for q in range(nq):
        a = np.zeros(1, np.int32)
        qmin[q], qmax[q] = h.query_bounds(a, out["ref_tables_direct"][q], r)
        qmin[q] = max(qmin[q], 0.0)                                       # db_query_4.cpp:262-263
        qt[q, 0] = ref.quantize(res["tables"][q, 0], qmin[q], qmax[q])    # QuantizerMAX
        dist[q] = ref.dump_distances(out["ref_interleaved"], n, m, qt[q, 0])   # scan_avx_4 per-vector values
    out.update(ref_qmin=qmin, ref_qmax=qmax, ref_qtables=qt, ref_distances=dist)
    h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "heap sizes", res["sizes"], "d<127 per query", (dist < 127).sum(1))


def ivf_case(ref, name, seed, dim, m, n, K, ma, nq, r, keep, empty=(), opq=False):
    rng = np.random.default_rng(seed)
    cb = synth.make_pq(rng, dim, m)
    cents = (2.0 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty)
    queries = synth.make_queries(rng, nq, dim)
    out = dict(dim=dim, m=m, r=r, ma=ma, keep=np.float32(keep), codebooks=cb, centroids=cents, codes=codes,
               labels=labels, offsets=offsets, queries=queries)
    h = ref.ivf(dim, m, cb, cents, codes, labels, offsets)
    if opq:
        # opq::rotate_multiple_vectors (quantizers.hpp:286-301) rotates the residuals before the tables
        out["rotation"] = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32)
        h.set_rotation(out["rotation"])
    h.prepare(keep)
    out["ref_start_size"] = np.array([h.starts_size(p) for p in range(K)], np.uint32)
    # K <= 256, so find_k_neighbors is correct as shipped
    out["ref_assign_fkn"] = ref.find_k_neighbors(queries, cents, ma)
    res = h.search(queries, ma, r, nthreads=1, blas_tables=False, want_tables=True)
    out["ref_assign"] = res["assign"]
    out["ref_heap_keys"], out["ref_heap_vals"], out["ref_heap_sizes"] = res["keys"], res["vals"], res["sizes"]
    out["ref_tables_used"] = res["tables"]                                # blas form (ma > 1), clamped
    qmin = np.zeros(nq, np.float32)
    qmax = np.zeros(nq, np.float32)
    qt = np.zeros((nq, ma, m, 16), np.int8)
    for q in range(nq):
        qmin[q], qmax[q] = h.query_bounds(res["assign"][q], res["tables"][q], r)
        qmin[q] = max(qmin[q], 0.0)
        qt[q] = ref.quantize(res["tables"][q], qmin[q], qmax[q])
    out.update(ref_qmin=qmin, ref_qmax=qmax, ref_qtables=qt)
    h.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "heap sizes", res["sizes"])


def encode_case(ref, name, seed, dim, m, n, bits=4):
    """base_pq::encode_multiple_vectors (quantizers.hpp:222-245) on seeded vectors."""
    rng = np.random.default_rng(seed)
    cb = synth.make_pq(rng, dim, m) if bits == 4 else rng.standard_normal((m, 1 << bits, dim // m)).astype(np.float32)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), dim=dim, m=m, bits=bits, codebooks=cb, vectors=x,
                        ref_codes=ref.encode(x, m, cb, bits))
    print(name, "encoded", n)


def add_vectors_case(ref, name, seed, dim, m, n, K, opq):
    """index_db::add_vectors (databases.hpp:270-298) with a plain PQ or an OPQ (residual -> rotate -> encode)."""
    rng = np.random.default_rng(seed)
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    x = (cents[rng.integers(0, K, n)] + rng.standard_normal((n, dim))).astype(np.float32)
    out = dict(dim=dim, m=m, codebooks=cb, centroids=cents, vectors=x)
    if opq:
        out["rotation"] = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32)
    out["ref_assign"], out["ref_codes"] = ref.index_add_vectors(x, m, cb, cents, out.get("rotation"))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "cells used", len(np.unique(out["ref_assign"])))


def archive_case(ref, name, seed, dim, m, n, K):
    """Database files as the reference writes them (flatdb_create.cpp:49-53): flat/index x pq/opq.
    Field order = the reference's save() members; field bytes = oracle/shims/cereal (cereal 1.2.2's
    binary archive restated; the real library is not in this image)."""
    import tempfile
    rng = np.random.default_rng(seed)
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    rot = np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32)
    cents = rng.standard_normal((K, dim)).astype(np.float32)
    icodes, labels, offsets = synth.make_ivf(rng, n, K, m, (1,))
    out = dict(dim=dim, m=m, codebooks=cb, codes=codes, rotation=rot, centroids=cents, ivf_codes=icodes, labels=labels,
               offsets=offsets)
    with tempfile.TemporaryDirectory() as tmp:
        for ivf in (False, True):
            for opq in (False, True):
                h = ref.ivf(dim, m, cb, cents, icodes, labels, offsets) if ivf else ref.flat(dim, m, cb, codes)
                if opq:
                    h.set_rotation(rot)
                path = os.path.join(tmp, "db")
                h.save(path)
                h.close()
                out["ref_%s_%s" % ("index" if ivf else "flat", "opq" if opq else "pq")] = np.fromfile(path, np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.size for k, v in out.items() if k.startswith("ref_")})


def adc_case(name, seed, dim, m, bits, n, nq, r, K=0, ma=1, opq=False):
    """db_query (plain ADC, "next" row N4): scanner_simple over 4-, 8- or 16-bit codes, flat or IVF."""
    rng = np.random.default_rng(seed)
    cs = m * bits // 8
    # 16-bit quantisers: the codebooks (m x 65536 x 2 floats) are the lattice of synth.lattice_codebook16, regenerated by
    # the tests instead of being stored (test_oracle.adc_db)
    cb = synth.lattice_codebook16(m) if bits == 16 else rng.standard_normal((m, 1 << bits, dim // m)).astype(np.float32)
    out = dict(dim=dim, m=m, bits=bits, r=r, ma=ma, codebooks=cb, queries=synth.make_queries(rng, nq, dim) * (2.0 if bits == 16 else 1.0))
    if opq:
        out["rotation"] = np.ascontiguousarray(np.linalg.qr(rng.standard_normal((dim, dim)))[0].astype(np.float32))
    if K:
        sizes = rng.multinomial(n, np.ones(K) / K)
        sizes[K // 3] = 0
        out["offsets"] = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        n = int(out["offsets"][-1])
        out["centroids"] = (2 * rng.standard_normal((K, dim))).astype(np.float32)
        out["labels"] = rng.permutation(n).astype(np.uint32)
    else:
        out["offsets"] = np.array([0, n], np.int64)
    out["codes"] = rng.integers(0, 256, (n, cs), dtype=np.uint8)
    out["ref_ids"], out["ref_dists"] = RefAdc().search(out, out["queries"], ma, r)
    if bits == 16:
        del out["codebooks"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "first distances", out["ref_dists"][0, :3])


def main():
    """python tests/golden/make_golden.py [name ...]  (no names: every fixture)"""
    ref = Ref()
    only = set(sys.argv[1:])

    def want(name):
        return not only or name in only

    for name, kw in (("encode_m16", dict(seed=105, dim=128, m=16, n=400)), ("encode_m32", dict(seed=106, dim=256, m=32, n=200)),
                     ("encode_8x8", dict(seed=112, dim=64, m=8, n=300, bits=8))):
        if want(name):
            encode_case(ref, name, **kw)
    for name, kw in (("add_ivf_pq", dict(seed=120, dim=64, m=16, n=600, K=40, opq=False)),
                     ("add_ivf_opq", dict(seed=121, dim=64, m=16, n=600, K=40, opq=True))):
        if want(name):
            add_vectors_case(ref, name, **kw)
    # n = 3003: not a multiple of 16 -> exercises the pad-lane duplicate quirk
    if want("flat_m16"):
        flat_case(ref, "flat_m16", seed=101, dim=128, m=16, n=3003, nq=8, r=20, keep=0.05)
    # 96-d, m=32 -> sq_dim 3 (only reachable by direct template instantiation)
    if want("flat_m32"):
        flat_case(ref, "flat_m32", seed=102, dim=96, m=32, n=2048, nq=6, r=16, keep=0.05)
    if want("ivf_m16"):
        ivf_case(ref, "ivf_m16", seed=103, dim=128, m=16, n=6000, K=24, ma=5, nq=8, r=20, keep=0.08, empty=(3,))
    if want("ivf_m32"):
        ivf_case(ref, "ivf_m32", seed=104, dim=96, m=32, n=4000, K=12, ma=3, nq=6, r=10, keep=0.1)
    if want("adc_flat_8x8"):
        adc_case("adc_flat_8x8", seed=109, dim=64, m=8, bits=8, n=4000, nq=8, r=20)
    if want("adc_ivf_16x8"):
        adc_case("adc_ivf_16x8", seed=110, dim=128, m=16, bits=8, n=6000, nq=8, r=20, K=16, ma=5)
    if want("adc_ivf_opq_32x4"):
        adc_case("adc_ivf_opq_32x4", seed=111, dim=96, m=32, bits=4, n=5000, nq=6, r=16, K=12, ma=4, opq=True)
    if want("adc_flat_2x16"):
        adc_case("adc_flat_2x16", seed=113, dim=4, m=2, bits=16, n=4000, nq=8, r=20)
    if want("adc_ivf_8x16"):
        adc_case("adc_ivf_8x16", seed=114, dim=16, m=8, bits=16, n=6000, nq=8, r=20, K=16, ma=5)
    if want("archives"):
        archive_case(ref, "archives", seed=108, dim=32, m=16, n=300, K=6)
    # OPQ: the rotation sits between the residuals and the tables
    if want("ivf_opq_m16"):
        ivf_case(ref, "ivf_opq_m16", seed=107, dim=64, m=16, n=5000, K=16, ma=4, nq=8, r=20, keep=0.08, empty=(9,), opq=True)


if __name__ == "__main__":
    main()
