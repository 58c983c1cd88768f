"""CPU check of the arithmetic behind the keep-prefix pre-selection of long flat prefixes (csrc/qadc_flatprep.cuh): the
provisional int8 tables  P[j][c] = floor((T[j][c] - min_j) * 127 / (U0 - sum_j min_j) - 0.05)  clamped to [0, 127]  must
never exclude a vector whose float32 ADC distance (sequential sum in sub-quantiser order, as the kernels and the oracle
compute it) is at or below U0 — otherwise qmax would no longer be the reference's.  The kernel's float32 operations are
restated with numpy float32 scalars; the GPU tests (test_flat_long_prefix_*) assert the end result bit for bit."""
import numpy as np
import pytest

F = np.float32


def provisional_tables(T, u0):
    """T: [m,16] float32 tables of one query, u0: the provisional bound.  Returns P [m,16] int32 (all zero when the scale
    is rejected, as flat_prefix_bound_kernel does)."""
    mn = T.min(axis=1)
    smin = F(0)
    for j in range(T.shape[0]):
        smin = F(smin + mn[j])
    rng_ = F(u0 - smin)
    ok = u0 < F(1e30) and rng_ > 0 and rng_ >= F(1e-2) * u0
    if not ok:
        return np.zeros(T.shape, np.int32)
    scale = F(F(127.0) / rng_)
    x = ((T - mn[:, None]).astype(F) * scale).astype(F) - F(0.05)
    return np.where(x <= 0, 0, np.where(x >= 127, 127, np.floor(x))).astype(np.int32)


def float_distances(T, nib):
    d = np.zeros(nib.shape[0], F)
    for j in range(T.shape[0]):
        d = (d + T[j][nib[:, j]]).astype(F)
    return d


@pytest.mark.parametrize("kind", ["gaussian", "offset", "tiny_spread", "huge_dynamic_range", "many_zero_entries"])
@pytest.mark.parametrize("m", [16, 32])
def test_no_vector_at_or_below_u0_is_excluded(kind, m):
    rng = np.random.default_rng(sum(map(ord, kind)) * 100 + m)
    n, r = 200000, 100
    nib = rng.integers(0, 16, (n, m), dtype=np.uint8)
    for trial in range(6):
        T = (rng.standard_normal((m, 16)) ** 2).astype(F) * F(4)
        if kind == "offset":                 # a large common term: the usable range is a small part of u0
            T = (T + F(30)).astype(F)
        elif kind == "tiny_spread":          # nearly constant sub-quantisers: the scale must be rejected, never wrong
            T = (F(1000) + T * F(1e-4)).astype(F)
        elif kind == "huge_dynamic_range":
            T = (T * (F(10) ** rng.integers(-3, 4, (m, 1)).astype(F))).astype(F)
        elif kind == "many_zero_entries":
            T[rng.random(T.shape) < 0.4] = F(0)
        d = float_distances(T, nib)
        for sample in (2048, 65536):
            u0 = np.sort(d[:sample])[r - 1]
            P = provisional_tables(T, u0)
            L = np.zeros(n, np.int64)
            for j in range(m):
                L += P[j][nib[:, j]]
            assert L[d <= u0].max() <= 127, (kind, m, trial, sample)
            # and the candidates always contain r vectors (the sample's own r smallest)
            assert (L <= 127).sum() >= r
