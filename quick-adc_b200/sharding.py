"""Host-side sharding of a database across the GPUs of one box.

Flat: contiguous ranges of whole 256-vector superblocks, so that global position =
shard offset + local position and the canonical order is preserved.  IVF: whole lists are
assigned to GPUs (greedy by size); a list is never split.  A flat shard keeps a replica of
the keep-prefix, so all shards derive identical quantisation bounds without a collective.
Inverted lists use "owner computes": a shard builds tables and scans keep-prefixes only for the
probes whose lists it holds and the shards exchange (min table entry, r smallest prefix
distances) per query, from which every shard derives the same qmin/qmax (owner_computes_search).
The other exchange steps are one all-gather of the per-shard top-r (key, id) lists and one
all-gather of the per-shard coarse candidates (the coarse quantizer's cells are split into
contiguous ranges, so ranking the queries against K cells costs K/G per GPU).
"""
import numpy as np

SB = 256


def start_size(size, keep):
    """starts_sizes[p] = max(1u, (unsigned)(size * keep)) in float32 (db_query_4.cpp:125-126)."""
    if size == 0:
        return 0
    return max(1, int(np.float32(size) * np.float32(keep)))


def flat_shard_range(n, rank, world):
    """[lo, hi) of shard `rank`: superblocks split as evenly as possible, in order."""
    n_sb = (n + SB - 1) // SB
    base, extra = divmod(n_sb, world)
    sb_lo = rank * base + min(rank, extra)
    sb_hi = sb_lo + base + (1 if rank < extra else 0)
    return min(sb_lo * SB, n), min(sb_hi * SB, n)


def ivf_list_owner(sizes, world):
    """Greedy longest-first assignment of whole lists to shards. Returns owner[p]."""
    sizes = np.asarray(sizes, np.int64)
    owner = np.zeros(len(sizes), np.int32)
    load = np.zeros(world, np.int64)
    for p in np.argsort(-sizes, kind="stable"):
        g = int(np.argmin(load))
        owner[p] = g
        load[g] += sizes[p]
    return owner


def coarse_range(K, rank, world):
    """[first, first + count) of the coarse cells rank `rank` ranks the queries against."""
    base, extra = divmod(K, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def all_gather_keys(keys, group=None):
    """All-gather of one int64 tensor [nq, k] -> [G, nq, k] in rank order (coarse candidates)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * keys.shape[0],) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
    dist.all_gather_into_tensor(out, keys.contiguous(), group=group)
    return out.view((world,) + tuple(keys.shape))


def sharded_coarse_assign(index, d_queries, nq, ma, K, rank, world, d_part_keys, d_assign, group=None):
    """Coarse assignment with the cells split over the ranks: partial ranking on this GPU
    (qadc_coarse_partial_device), one all-gather, merge (qadc_coarse_merge_device).  d_part_keys:
    int64 CUDA tensor [nq, ma] (scratch), d_assign: int32 CUDA tensor [nq, ma] (result, identical on
    every rank and identical to the unsharded assignment)."""
    first, count = coarse_range(K, rank, world)
    index.coarse_partial_device(d_queries.data_ptr(), nq, ma, first, count, d_part_keys.data_ptr())
    gathered = all_gather_keys(d_part_keys, group) if world > 1 else d_part_keys.view((1,) + tuple(d_part_keys.shape))
    index.coarse_merge_device(gathered.data_ptr(), world, nq, ma, d_assign.data_ptr())
    return d_assign


def all_gather_topk(keys, ids, group=None):
    """One all-gather of the per-shard top-r lists: keys int64 [nq, r], ids int32 [nq, r]
    -> ([G, nq, r], [G, nq, r]) in rank order, the layout qadc_merge_shards_device reads.
    NCCL on CUDA tensors (NVLink/NVSwitch), gloo on CPU tensors (tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nq = keys.shape[0]
    gk = torch.empty((world * nq,) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
    gi = torch.empty((world * nq,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
    k, i = keys.contiguous(), ids.contiguous()
    coalesce = getattr(dist, "_coalescing_manager", None) if keys.is_cuda else None
    done = False
    if coalesce is not None:
        try:   # both all-gathers in ONE NCCL group launch (ncclGroupStart/End): one kernel, one latency
            with coalesce(group=group, device=keys.device, async_ops=False):
                dist.all_gather_into_tensor(gk, k, group=group)   # concatenation along dim 0
                dist.all_gather_into_tensor(gi, i, group=group)
            done = True
        except (TypeError, RuntimeError, AssertionError):
            done = False
    if not done:
        dist.all_gather_into_tensor(gk, k, group=group)
        dist.all_gather_into_tensor(gi, i, group=group)
    return gk.view((world,) + tuple(keys.shape)), gi.view((world,) + tuple(ids.shape))


def owner_computes_search(index, d_queries, d_assign, nq, ma, r, d_local, d_ids, d_dists, d_counts, d_keys, group=None):
    """Sharded inverted lists without replicated work (qadc_tables_local_device -> all-gather ->
    qadc_search_bounded_device): this shard's tables and prefix distances for the probes it owns, one
    all-gather of nq x (r + 1) floats per shard, then bounds, int8 tables, scan and local top-r.
    d_local: float32 CUDA tensor [nq, r + 1] (scratch); the other tensors as for search_device.
    At most 32768 queries per call."""
    import torch
    import torch.distributed as dist
    index.tables_local_device(d_queries.data_ptr(), d_assign.data_ptr(), nq, ma, r, d_local.data_ptr())
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        gathered = torch.empty((world * nq, r + 1), dtype=torch.float32, device=d_local.device)
        dist.all_gather_into_tensor(gathered, d_local, group=group)
    else:
        gathered = d_local
    index.search_bounded_device(gathered.data_ptr(), world, nq, ma, r, d_ids.data_ptr(), d_dists.data_ptr(),
                                d_counts.data_ptr(), None if d_keys is None else d_keys.data_ptr())
