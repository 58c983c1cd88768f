"""Seeded synthetic inputs shared by the tests (SURVEY §8d shapes, scaled down)."""
import numpy as np


def make_pq(rng, dim, m):
    """codebooks m x 16 x (dim/m), N(0,1)."""
    return rng.standard_normal((m, 16, dim // m)).astype(np.float32)


def lattice_codebook16(m):
    """A 16-bit product quantiser whose codebooks need no storage: sub-quantiser j (2-d sub-vectors) has the 65 536
    centroids of a 256 x 256 lattice, centroid c = ((c & 255) - 127.5) / 32 + j / 64, ((c >> 8) - 127.5) / 32 - j / 64
    (exactly representable in float32).  Shape m x 65536 x 2."""
    c = np.arange(65536)
    cb = np.empty((m, 65536, 2), np.float32)
    for j in range(m):
        cb[j, :, 0] = ((c & 255) - 127.5) / 32 + j / 64
        cb[j, :, 1] = ((c >> 8) - 127.5) / 32 - j / 64
    return cb


def make_codes(rng, n, m):
    """i.i.d. uniform nibbles, generated directly as bytes (row-major n x m/2)."""
    return rng.integers(0, 256, (n, m // 2), dtype=np.uint8)


def make_queries(rng, nq, dim):
    return rng.standard_normal((nq, dim)).astype(np.float32)


def make_qtables(rng, shape_prefix, m, lo=0, hi=40, p_sat=0.15):
    """Random int8 tables in [0,127] with a share of saturated (127) entries, like real ones."""
    t = rng.integers(lo, hi, tuple(shape_prefix) + (m, 16)).astype(np.int8)
    sat = rng.random(t.shape) < p_sat
    t[sat] = 127
    return t


def make_ivf(rng, n, K, m, empty=()):
    """Random list membership: sizes ~ multinomial(n, uniform), labels a permutation of 0..n-1.
    Partitions in `empty` get size 0."""
    p = np.ones(K)
    for e in empty:
        p[e] = 0
    sizes = rng.multinomial(n, p / p.sum())
    offsets = np.zeros(K + 1, np.int64)
    offsets[1:] = np.cumsum(sizes)
    labels = rng.permutation(n).astype(np.uint32)
    codes = make_codes(rng, n, m)
    return codes, labels, offsets


def canonical_from_distances(d_parts, labels_parts, r):
    """SURVEY §8c Stage S evaluated with numpy: d_parts[a] = int8 distances of probe a's
    partition (scan order), labels_parts[a] = its labels or None. Returns ids, dists, count."""
    keys, ids = [], []
    for a, d in enumerate(d_parts):
        pos = np.nonzero(d < 127)[0]
        keys.append((d[pos].astype(np.uint64) << np.uint64(48)) | (np.uint64(a) << np.uint64(32)) | pos.astype(np.uint64))
        lab = labels_parts[a]
        ids.append(pos.astype(np.uint32) if lab is None else lab[pos])
    keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    ids = np.concatenate(ids) if ids else np.zeros(0, np.uint32)
    order = np.argsort(keys, kind="stable")[:r]
    n = len(order)
    out_ids = np.zeros(r, np.uint32)
    out_d = np.full(r, 127, np.int8)
    out_ids[:n] = ids[order]
    out_d[:n] = (keys[order] >> np.uint64(48)).astype(np.int8)
    return out_ids, out_d, n
