#!/usr/bin/env python
"""BASELINE config 5 end to end: Deep1B-shaped IVF-65536, PQ 16x4 (96-d, sq_dim 6), nprobe 128,
inverted lists sharded over the GPUs of one box (whole lists per GPU, greedy by size), replicated
keep-prefixes, one NCCL all-gather of the per-shard top-r + merge on every rank.

    torchrun --nnodes=1 --nproc-per-node N tools/bench_ivf_sharded.py [--n-vectors 1000000000] [--queries 10000]

Not a bench line (bench.py measures the flat 1e9 scan); it prints one JSON line with queries/s
and vectors scanned/s, and cross-checks a few queries against a single-GPU unsharded run when the
database fits (--check).  Codes are the bench's counter-based hash of the global vector index; list p
owns the contiguous index range [offsets[p], offsets[p+1]) and labels are the global indices."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (code generator)

DIM, M, K, MA, R, KEEP = 96, 16, 65536, 128, 100, 0.0005


def main():
    import torch
    import torch.distributed as dist
    import qadc_b200
    from qadc_b200 import sharding

    ap = argparse.ArgumentParser()
    ap.add_argument("--n-vectors", type=int, default=10 ** 9)
    ap.add_argument("--queries", type=int, default=10000)
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--ma", type=int, default=MA)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", type=int, default=0, help="queries compared with an unsharded single-GPU index on rank 0")
    ap.add_argument("--replicated-coarse", action="store_true",
                    help="every rank ranks the queries against all K cells (no coarse exchange), for comparison")
    ap.add_argument("--replicated-tables", action="store_true",
                    help="round-1 pipeline: every rank builds the tables of all probes from replicated keep-prefixes")
    ap.add_argument("--as-rank-of", type=int, default=0, metavar="W",
                    help="single process: hold only rank 0's lists of a W-way sharding (per-GPU work of the W-GPU run, no exchange)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    N, nq, Kc, ma = args.n_vectors, args.queries, args.k, args.ma

    rng = np.random.default_rng(77)
    cb = rng.standard_normal((M, 16, DIM // M)).astype(np.float32)
    cents = (2 * rng.standard_normal((Kc, DIM))).astype(np.float32)
    queries = rng.standard_normal((nq, DIM)).astype(np.float32)
    sizes = rng.multinomial(N, np.ones(Kc) / Kc).astype(np.int64)
    offsets = np.zeros(Kc + 1, np.int64); offsets[1:] = np.cumsum(sizes)
    shards = args.as_rank_of if (args.as_rank_of and world == 1) else world
    owner = sharding.ivf_list_owner(sizes, shards)

    def build(index, mine):
        """mine: boolean mask of lists this index owns; prefixes are replicated for all lists."""
        index.set_pq(DIM, M, cb)
        index.set_coarse(cents)
        index.begin_database(np.where(mine, sizes, 0).astype(np.uint32), True)
        index.set_owned_partitions(mine)
        t0 = time.perf_counter()
        for p in range(Kc):
            n_p = int(sizes[p])
            if n_p == 0:
                continue
            npre = sharding.start_size(n_p, KEEP)
            if mine[p]:
                codes = bench.codes_torch(int(offsets[p]), int(offsets[p + 1]), dev)
                labels = torch.arange(int(offsets[p]), int(offsets[p + 1]), dtype=torch.int64, device=dev).to(torch.int32)
                stream.synchronize()
                index.upload_codes_device(p, 0, n_p, codes.data_ptr(), labels.data_ptr())
                if shards > 1:
                    index.set_prefix_device(p, codes.data_ptr(), npre)
            elif shards > 1:
                pre = bench.codes_torch(int(offsets[p]), int(offsets[p]) + npre, dev)
                stream.synchronize()
                index.set_prefix_device(p, pre.data_ptr(), npre)
        index.finalize(KEEP)
        return time.perf_counter() - t0

    ix = qadc_b200.Index(local, stream.cuda_stream)
    t_build = build(ix, owner == rank)
    d_q = torch.from_numpy(queries).to(dev)
    d_ids = torch.empty((nq, R), dtype=torch.int32, device=dev); d_d = torch.empty((nq, R), dtype=torch.int8, device=dev)
    d_cnt = torch.empty(nq, dtype=torch.int32, device=dev); d_keys = torch.empty((nq, R), dtype=torch.int64, device=dev)
    o_ids, o_d, o_cnt = torch.empty_like(d_ids), torch.empty_like(d_d), torch.empty_like(d_cnt)

    # coarse assignment: cells split over the ranks, one all-gather of the partial rankings, merge
    split_coarse = shards > 1 and not args.replicated_coarse
    d_assign = torch.empty((nq, ma), dtype=torch.int32, device=dev)
    d_part = torch.empty((nq, ma), dtype=torch.int64, device=dev)
    d_gath = None
    if split_coarse and world == 1:
        # --as-rank-of: the other ranks' partial rankings are computed once, outside the timed steps
        d_gath = torch.empty((shards, nq, ma), dtype=torch.int64, device=dev)
        for g in range(shards):
            first, count = sharding.coarse_range(Kc, g, shards)
            ix.coarse_partial_device(d_q.data_ptr(), nq, ma, first, count, d_gath[g].data_ptr())
        ix.synchronize()

    d_local = torch.empty((nq, R + 1), dtype=torch.float32, device=dev)
    d_local_all = torch.empty((max(shards, 1), nq, R + 1), dtype=torch.float32, device=dev)

    def step():
        if split_coarse:
            if world > 1:
                sharding.sharded_coarse_assign(ix, d_q, nq, ma, Kc, rank, world, d_part, d_assign)
            else:
                first, count = sharding.coarse_range(Kc, 0, shards)
                ix.coarse_partial_device(d_q.data_ptr(), nq, ma, first, count, d_gath[0].data_ptr())
                ix.coarse_merge_device(d_gath.data_ptr(), shards, nq, ma, d_assign.data_ptr())
            if args.replicated_tables:
                ix.search_assigned_device(d_q.data_ptr(), d_assign.data_ptr(), nq, ma, R, d_ids.data_ptr(), d_d.data_ptr(),
                                          d_cnt.data_ptr(), d_keys.data_ptr())
            elif world > 1:
                sharding.owner_computes_search(ix, d_q, d_assign, nq, ma, R, d_local, d_ids, d_d, d_cnt, d_keys)
            else:
                # --as-rank-of: the other shards' bound shares are stood in for by copies of this shard's (same work per
                # GPU as the real run; the results are not those of the full database and are not checked)
                ix.tables_local_device(d_q.data_ptr(), d_assign.data_ptr(), nq, ma, R, d_local_all[0].data_ptr())
                d_local_all[1:] = d_local_all[0]
                ix.search_bounded_device(d_local_all.data_ptr(), shards, nq, ma, R, d_ids.data_ptr(), d_d.data_ptr(),
                                         d_cnt.data_ptr(), d_keys.data_ptr())
        else:
            ix.search_device(d_q.data_ptr(), nq, ma, R, d_ids.data_ptr(), d_d.data_ptr(), d_cnt.data_ptr(), d_keys.data_ptr())
        if world > 1:
            gk, gi = sharding.all_gather_topk(d_keys, d_ids)
            ix.merge_shards_device(gk.data_ptr(), gi.data_ptr(), world, nq, R, o_ids.data_ptr(), o_d.data_ptr(), o_cnt.data_ptr())
        else:
            o_ids.copy_(d_ids); o_d.copy_(d_d); o_cnt.copy_(d_cnt)

    for _ in range(2):
        step()
    ix.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = os.environ.get("QADC_PROFILE_RANGE") == "1"   # ncu --profile-from-start off: only the timed steps
    if prof:
        torch.cuda.profiler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if prof:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    ix.synchronize()
    stages = None
    if rank == 0:   # per-stage CUDA-event times of one more batch through the host-buffer entry point
        _, _, _, met = ix.search(queries, ma, R, want_metrics=True)
        stages = {n: float(getattr(met, n)) for n, _ in met._fields_}

    ok = None
    if args.check and rank == 0 and N * (M // 2) < 60e9:
        one = qadc_b200.Index(local, stream.cuda_stream)
        build(one, np.ones(Kc, bool))
        c_ids = torch.empty((args.check, R), dtype=torch.int32, device=dev); c_d = torch.empty((args.check, R), dtype=torch.int8, device=dev)
        c_cnt = torch.empty(args.check, dtype=torch.int32, device=dev)
        one.search_device(d_q.data_ptr(), args.check, ma, R, c_ids.data_ptr(), c_d.data_ptr(), c_cnt.data_ptr())
        one.synchronize()
        ok = bool(torch.equal(c_ids, o_ids[:args.check]) and torch.equal(c_d, o_d[:args.check]) and torch.equal(c_cnt, o_cnt[:args.check]))
        one.close()
    if rank == 0:
        scanned = float(ma) * N / Kc
        print(json.dumps({"config": "5: Deep1B-shaped IVF-%d PQ 16x4, nprobe %d, sharded lists" % (Kc, ma), "n_vectors": N,
                          "n_gpus": world, "shards": shards, "coarse": "split" if split_coarse else "replicated",
                          "tables": "replicated" if (args.replicated_tables or not split_coarse) else "owner computes", "queries": nq, "ms_per_batch": ms, "queries_per_s": nq / (ms * 1e-3),
                          "vectors_scanned_per_s": scanned * nq / (ms * 1e-3), "build_seconds": t_build, "stage_metrics": stages,
                          "matches_unsharded": ok}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ix.close()


if __name__ == "__main__":
    main()
