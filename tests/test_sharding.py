"""CPU tests of the multi-GPU host logic: shard planning, and the all-gather layout exercised
with two gloo ranks (the merge kernel itself is covered by -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flat_shards_partition_the_database(qadc):
    from qadc_b200 import sharding
    for n in (1, 255, 256, 257, 100000, 10 ** 9):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for g in range(world):
                lo, hi = sharding.flat_shard_range(n, g, world)
                assert lo == prev and lo <= hi <= n
                assert lo % 256 == 0 or lo == n
                prev = hi
            assert prev == n
    lo, hi = sharding.flat_shard_range(10 ** 9, 3, 8)
    assert abs((hi - lo) - 10 ** 9 / 8) <= 256


def test_start_size_matches_reference_arithmetic(qadc, oracle):
    from qadc_b200 import sharding
    rng = np.random.default_rng(0)
    for size in [0, 1, 2, 99, 100, 244, 10 ** 6, 10 ** 9, 2 ** 31] + rng.integers(1, 2 ** 31, 50).tolist():
        for keep in (0.01, 0.0005, 0.00213, 1.0, 0.05):
            assert sharding.start_size(int(size), keep) == oracle.start_size(int(size), keep)
    assert sharding.start_size(10 ** 6, 0.01) == 10000      # SURVEY a9 [probe]


def test_ivf_owner_is_balanced(qadc):
    from qadc_b200 import sharding
    rng = np.random.default_rng(1)
    sizes = rng.multinomial(10 ** 6, np.ones(4096) / 4096)
    owner = sharding.ivf_list_owner(sizes, 8)
    load = np.bincount(owner, weights=sizes, minlength=8)
    assert load.sum() == sizes.sum() and load.max() - load.min() <= sizes.max()


def test_coarse_ranges_partition_the_cells(qadc):
    from qadc_b200 import sharding
    for K in (1, 7, 37, 4096, 65536):
        for world in (1, 2, 3, 8):
            nxt = 0
            for rank in range(world):
                first, count = sharding.coarse_range(K, rank, world)
                assert first == nxt and count >= 0
                nxt = first + count
            assert nxt == K


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import qadc_b200  # noqa: F401
    from qadc_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    nq, r = 5, 7
    # each rank's local sorted top-r (distance << 48 | position), positions inside its shard range
    lo, hi = sharding.flat_shard_range(10000, rank, world)
    d = np.sort(rng.integers(0, 30, (nq, r)), axis=1).astype(np.int64)
    pos = np.sort(rng.integers(lo, hi, (nq, r)), axis=1).astype(np.int64)
    keys = np.sort((d << 48) | pos, axis=1)
    ids = (keys & 0xffffffff).astype(np.int32)
    gk, gi = sharding.all_gather_topk(torch.from_numpy(keys), torch.from_numpy(ids))
    # coarse candidates: each rank's ma best cells of its own cell range, one all-gather
    first, count = sharding.coarse_range(37, rank, world)
    ck = np.sort((rng.integers(1, 1 << 30, (nq, 4)).astype(np.int64) << 32) | rng.integers(first, first + count, (nq, 4)), axis=1)
    gc = sharding.all_gather_keys(torch.from_numpy(ck))
    if rank == 0:
        q.put((gk.numpy(), gi.numpy(), gc.numpy()))
    else:
        q.put((keys, ids, ck))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_layout_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gathered = [g for g in got if g[0].ndim == 3][0]
    local1 = [g for g in got if g[0].ndim == 2][0]
    gk, gi, gc = gathered
    assert gc.shape == (2, 5, 4) and np.array_equal(gc[1], local1[2])          # [G][nq][ma], rank order
    assert np.all((gc[0] & 0xffffffff) < 19) and np.all((gc[1] & 0xffffffff) >= 19)   # cell ranges 0..18 | 19..36
    assert gk.shape == (2, 5, 7) and gi.shape == (2, 5, 7)
    assert np.array_equal(gk[1], local1[0]) and np.array_equal(gi[1], local1[1])   # rank order, [G][nq][r]
    # merging the gathered lists by key reproduces the global top-r of the union
    merged = np.sort(gk.transpose(1, 0, 2).reshape(5, -1), axis=1)[:, :7]
    assert np.all(np.diff(merged, axis=1) >= 0)
    assert np.all(merged[:, 0] == np.minimum(gk[0][:, 0], gk[1][:, 0]))
