"""qadc_multi_* (-m gpu): one database sharded over several GPUs by one process == the single-GPU result, bit for
bit.  On a box with several GPUs the exchange is NCCL (asserted); on a one-GPU box the same layer runs with
virtual shards on GPU 0."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    n = torch.cuda.device_count()
    return (list(range(n)), True) if n > 1 else ([0, 0, 0], False)


@pytest.mark.parametrize("m,dim", [(16, 128), (32, 96)])
def test_multi_flat_equals_single(qadc, oracle, m, dim):
    rng = np.random.default_rng(500 + m)
    n, nq, r, keep = 333333, 40, 100, 0.01
    cb = synth.make_pq(rng, dim, m)
    codes = synth.make_codes(rng, n, m)
    q = synth.make_queries(rng, nq, dim)
    one = qadc.Index(0)
    one.set_pq(dim, m, cb)
    one.load_flat(codes, keep)
    exp = one.search(q, 1, r)
    one.close()
    devs, nccl = _devices()
    mi = qadc.MultiIndex(devs)
    assert mi.uses_nccl == nccl
    mi.set_pq(dim, m, cb)
    mi.load_flat(codes, keep)
    for _ in range(2):   # twice: scratch reuse
        ids, d, cnt, met = mi.search(q, 1, r, want_metrics=True)
        assert np.array_equal(ids, exp[0]) and np.array_equal(d, exp[1]) and np.array_equal(cnt, exp[2])
    assert met.scan_us > 0
    o = oracle.search(dict(dim=dim, m=m, codebooks=cb, codes=codes, keep=keep, offsets=np.array([0, n], np.int64)), q[:6], 1, r,
                      want_tables=False)
    assert np.array_equal(ids[:6], o["ids"]) and np.array_equal(d[:6], o["d"])
    mi.close()


def test_multi_ivf_equals_single(qadc, oracle):
    rng = np.random.default_rng(77)
    dim, m, n, K, ma, nq, r, keep = 96, 16, 200000, 300, 24, 50, 100, 0.02
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(0, 17, 299))
    q = synth.make_queries(rng, nq, dim)
    one = qadc.Index(0)
    one.set_pq(dim, m, cb); one.set_coarse(cents)
    one.load_ivf(codes, labels, offsets, keep)
    exp = one.search(q, ma, r)
    one.close()
    devs, nccl = _devices()
    mi = qadc.MultiIndex(devs)
    mi.set_pq(dim, m, cb); mi.set_coarse(cents)
    mi.load_ivf(codes, labels, offsets, keep)
    ids, d, cnt = mi.search(q, ma, r)
    assert np.array_equal(ids, exp[0]) and np.array_equal(d, exp[1]) and np.array_equal(cnt, exp[2])
    o = oracle.search(dict(dim=dim, m=m, codebooks=cb, centroids=cents, codes=codes, labels=labels, keep=keep, offsets=offsets),
                      q[:8], ma, r, want_tables=False)
    assert np.array_equal(ids[:8], o["ids"]) and np.array_equal(d[:8], o["d"]) and np.array_equal(cnt[:8], o["count"])
    mi.close()


def test_multi_reports_errors(qadc):
    with pytest.raises(qadc.QadcError):
        qadc.MultiIndex([0, 99])
    rng = np.random.default_rng(3)
    mi = qadc.MultiIndex([0, 0])
    mi.set_pq(128, 16, synth.make_pq(rng, 128, 16))
    with pytest.raises(qadc.QadcError) as e:
        mi.search(synth.make_queries(rng, 2, 128), 1, 10)
    assert e.value.code == qadc.QADC_ESTATE
    mi.load_flat(synth.make_codes(rng, 900, 16), 0.01)      # prefix of 9 vectors < r
    with pytest.raises(qadc.QadcError) as e:
        mi.search(synth.make_queries(rng, 2, 128), 1, 10)
    assert e.value.code == qadc.QADC_EBOUND
    mi.close()


def test_multi_ivf_more_than_one_sub_batch_and_bound_error(qadc):
    """The owner-computes exchange of qadc_multi_search runs in sub-batches of 32768 queries: 33000 queries (the first 200
    repeated) return the single-GPU result for every copy; a keep so small that the probed prefixes hold fewer than r
    vectors reports the reference's "Max quantization bound too high" (QADC_EBOUND) through the sharded bounds too."""
    rng = np.random.default_rng(91)
    dim, m, n, K, ma, r, keep = 64, 16, 40000, 40, 6, 20, 0.05
    cb = synth.make_pq(rng, dim, m)
    cents = (2 * rng.standard_normal((K, dim))).astype(np.float32)
    codes, labels, offsets = synth.make_ivf(rng, n, K, m, empty=(3,))
    base_q = synth.make_queries(rng, 200, dim)
    q = np.tile(base_q, (165, 1))          # 33000 queries
    one = qadc.Index(0)
    one.set_pq(dim, m, cb); one.set_coarse(cents)
    one.load_ivf(codes, labels, offsets, keep)
    exp = one.search(base_q, ma, r)
    one.close()
    devs, _ = _devices()
    mi = qadc.MultiIndex(devs)
    mi.set_pq(dim, m, cb); mi.set_coarse(cents)
    mi.load_ivf(codes, labels, offsets, keep)
    ids, d, cnt = mi.search(q, ma, r)
    for name, got, want in (("ids", ids, exp[0]), ("d", d, exp[1]), ("cnt", cnt, exp[2])):
        assert np.array_equal(got.reshape((165, 200) + got.shape[1:]), np.broadcast_to(want, (165,) + want.shape)), name
    mi.close()
    mi = qadc.MultiIndex(devs)
    mi.set_pq(dim, m, cb); mi.set_coarse(cents)
    mi.load_ivf(codes, labels, offsets, 0.0001)     # one prefix vector per list: 6 probes x 1 < r
    with pytest.raises(qadc.QadcError) as e:
        mi.search(base_q[:4], ma, r)
    assert e.value.code == qadc.QADC_EBOUND
    mi.close()
