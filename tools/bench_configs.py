#!/usr/bin/env python
"""Throughput of BASELINE.json configs 1-3 (+ scaled 5, 6) on one B200 (parity-test shapes, not bench
lines), each with a spot check of a few queries: the canonical top-r recomputed with numpy from the
per-vector distances of an independent kernel (qadc_dump_distances), like bench.py --verify.  Parity
against the oracle / the reference is the job of tests/.  usage: tools/bench_configs.py [1] [2] [3] [5] [6]"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import qadc_b200

R = 100


def spot_check(ix, db, queries, ma, ids, d, cnt):
    """Canonical rule (d, probe rank, position) over d < 127, evaluated by numpy on dumped distances."""
    tabs = ix.build_tables(queries, ma, R)
    offsets, labels = db["offsets"], db.get("labels")
    ok = True
    for s in range(queries.shape[0]):
        dist, rank, pos = [], [], []
        for a, p in enumerate(tabs["assign"][s]):
            if offsets[p + 1] == offsets[p]:
                continue
            dd = ix.dump_distances(int(p), tabs["qtables"][s, a])
            keep = np.nonzero(dd < 127)[0]
            dist.append(dd[keep]); rank.append(np.full(len(keep), a)); pos.append(keep + offsets[p])
        dist, rank, pos = (np.concatenate(x) if x else np.zeros(0, np.int64) for x in (dist, rank, pos))
        order = np.lexsort((pos, rank, dist))[:R]
        e_ids = (labels[pos[order]] if labels is not None else pos[order]).astype(np.uint32)
        n = len(order)
        ok = ok and cnt[s] == n and np.array_equal(ids[s][:n], e_ids) and np.array_equal(d[s][:n], dist[order])
    return bool(ok)


def run(name, ix, db, queries, ma, check=6, reps=3, qb=None):
    nq = queries.shape[0]
    if qb is not None:
        ix.set_option("flat_qb", qb)
    ids, d, cnt, met = ix.search(queries, ma, R, want_metrics=True)   # warm-up + result
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ids, d, cnt, met = ix.search(queries, ma, R, want_metrics=True)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    sel = np.linspace(0, nq - 1, check).astype(int)
    ok = spot_check(ix, db, queries[sel], ma, ids[sel], d[sel], cnt[sel])
    scanned = db["scanned_per_query"]
    out = dict(config=name, nq=nq, ma=ma, seconds=t, queries_per_s=nq / t, vectors_scanned_per_s=scanned * nq / t,
               index_us=met.index_us, table_us=met.table_us, scan_us=met.scan_us, h2d_us=met.h2d_us, d2h_us=met.d2h_us,
               launches=ix.last_launch_count(), spot_check=ok, qb=qb)
    print(json.dumps(out), flush=True)


def config1(nq=10000):
    rng = np.random.default_rng(1235)
    n, dim, m = 10 ** 6, 128, 16
    cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
    codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.load_flat(codes, 0.01)
    db = dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=0.01, offsets=np.array([0, n], np.int64), scanned_per_query=n)
    for qb in (1, 2, 4):
        run("1: SIFT1M-shaped flat 16x4, 10k queries", ix, db, q, 1, qb=qb)
    ix.close()


def config2(nq=10000):
    rng = np.random.default_rng(1236)
    n, dim, m, K, ma = 10 ** 6, 128, 16, 4096, 64
    cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    sizes = rng.multinomial(n, np.ones(K) / K)
    offsets = np.zeros(K + 1, np.int64); offsets[1:] = np.cumsum(sizes)
    labels = rng.permutation(n).astype(np.uint32)
    codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.set_coarse(cents); ix.load_ivf(codes, labels, offsets, 0.01)
    db = dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=0.01, offsets=offsets,
              scanned_per_query=float(ma * n / K))
    run("2: SIFT1M-shaped IVF-4096 16x4, nprobe 64, 10k queries", ix, db, q, ma)
    ix.close()


def config3(nq=10000):
    rng = np.random.default_rng(1237)
    n, dim, m = 10 ** 7, 96, 32
    cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
    codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.load_flat(codes, 0.001)
    db = dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=0.001, offsets=np.array([0, n], np.int64), scanned_per_query=n)
    for qb in (1, 2):
        run("3: Deep10M-shaped flat 32x4, 10k queries", ix, db, q, 1, check=3, qb=qb)
    ix.close()


def config5(nq=2000, n=20 * 10 ** 6):
    """Deep1B-shaped IVF-65536, 96-d, 16x4 (sq_dim 6), nprobe 128 — at 20M vectors on one GPU."""
    rng = np.random.default_rng(1239)
    dim, m, K, ma = 96, 16, 65536, 128
    cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    sizes = rng.multinomial(n, np.ones(K) / K)
    offsets = np.zeros(K + 1, np.int64); offsets[1:] = np.cumsum(sizes)
    labels = rng.permutation(n).astype(np.uint32)
    codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    t0 = time.perf_counter()
    ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.set_coarse(cents); ix.load_ivf(codes, labels, offsets, 0.01)
    print("load_ivf seconds", time.perf_counter() - t0, flush=True)
    db = dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=0.01, offsets=offsets,
              scanned_per_query=float(ma * n / K))
    run("5 (scaled to 20M): Deep1B-shaped IVF-65536 16x4, nprobe 128", ix, db, q, ma, check=3)
    ix.close()


def config6(nq=2000, n=20 * 10 ** 6):
    """Long inverted lists (about 20k vectors each, like Deep1B / IVF-65536): IVF-1024, nprobe 16."""
    rng = np.random.default_rng(1240)
    dim, m, K, ma = 96, 16, 1024, 16
    cb = rng.standard_normal((m, 16, dim // m)).astype(np.float32)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    sizes = rng.multinomial(n, np.ones(K) / K)
    offsets = np.zeros(K + 1, np.int64); offsets[1:] = np.cumsum(sizes)
    labels = rng.permutation(n).astype(np.uint32)
    codes = rng.integers(0, 256, (n, m // 2), dtype=np.uint8)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    ix = qadc_b200.Index(0); ix.set_pq(dim, m, cb); ix.set_coarse(cents); ix.load_ivf(codes, labels, offsets, 0.001)
    db = dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=0.001, offsets=offsets,
              scanned_per_query=float(ma * n / K))
    run("6: long lists, IVF-1024 x 20k vectors, nprobe 16", ix, db, q, ma, check=3)
    ix.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "3"]
    for w in which:
        {"1": config1, "2": config2, "3": config3, "5": config5, "6": config6}[w]()
