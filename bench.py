#!/usr/bin/env python
"""bench.py — vectors scanned/s of the Quick ADC search path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[3], the one the metric is quoted on; it fits one GPU):
SIFT1B-shaped synthetic flat database, 1e9 vectors x PQ m=16 x 4 bit (8-byte codes),
r = 100, keep = 0.05 %, sharded contiguously over the N GPUs of the box (strong scaling).
A step = one batch of `--queries` queries through the whole hot path (float tables, keep-prefix
scan, bounds, int8 tables, 4-bit scan, top-r, shard merge).  Codes are a counter-based hash of
the global vector index, so every sharding sees the same database.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the reference's own AVX2 scan (oracle/_ref, the unmodified sources
compiled by oracle/Makefile) on the host cores with OpenMP over queries, on a bounded sample
of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM, M, R, KEEP = 128, 16, 100, 0.0005
SEED = 1234 + 3
CODE_BYTES = M // 2
METRIC = "vectors scanned/s/GPU & HBM roofline % (16x4 PQ); queries/s at 1/2/4/8 GPU"


def mix64_np(idx):
    """splitmix64 finaliser on uint64 (one 8-byte code per vector)."""
    z = (idx + np.uint64(0x9E3779B97F4A7C15) * np.uint64(SEED)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def codes_np(lo, hi):
    with np.errstate(over="ignore"):
        return mix64_np(np.arange(lo, hi, dtype=np.uint64)).view(np.uint8).reshape(-1, CODE_BYTES)


def codes_torch_at(idx):
    """Same generator on the device for an int64 tensor of vector indices (int64 arithmetic wraps like uint64)."""
    import torch

    def c(v):  # uint64 constant as int64
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(x, s):  # logical shift right on int64
        return (x >> s) & ((1 << (64 - s)) - 1)

    z = idx + c((0x9E3779B97F4A7C15 * SEED) & ((1 << 64) - 1))
    z = (z ^ lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    return z.view(torch.uint8).view(-1, CODE_BYTES)


def codes_torch(lo, hi, device):
    import torch
    return codes_torch_at(torch.arange(lo, hi, dtype=torch.int64, device=device))


def make_quantizer_and_queries(nq):
    rng = np.random.default_rng(SEED)
    cb = rng.standard_normal((M, 16, DIM // M)).astype(np.float32)
    queries = rng.standard_normal((nq, DIM)).astype(np.float32)
    return cb, queries


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev = device_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(self.dev)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for rw in rows:
            if len(rw) < 8:
                continue
            try:
                sm.append(float(rw[1])); mx.append(float(rw[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), rw[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(n_local, nq, qb):
    """(bytes, source): dram__bytes_read.sum + dram__bytes_write.sum of one scan launch.  NOT measured in this run (ncu
    cannot run inside a timed bench): read from the committed `ncu --set full` capture of exactly the default workload
    (1e9 local vectors, 16 queries per step, 1 query per pass); (None, None) for any other configuration."""
    if (n_local, nq, qb) != (10 ** 9, 16, 1):
        return None, None
    for name in ("r02c_scan_flat_16x4_1B_ncu_full.txt", "r02b_scan_flat_16x4_1B_ncu_full.txt", "r02_scan_flat_16x4_1B_ncu_full.txt", "r01_scan_flat_16x4_1B_ncu_full.txt"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            vals = {}
            for line in open(p):
                t = line.split()
                if len(t) >= 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    vals[t[0]] = float(t[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[t[2]]
            return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], f"committed ncu capture profiles/{name} (same workload, not this run)"
        except Exception:
            continue
    return None, None


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(n_cpu, nq, threads, steps, warmup, log=None):
    """The reference's own CPU path (oracle/_ref) on vectors [0, n_cpu) of the same database."""
    from oracle.pyoracle import Ref
    if not Ref.available():
        return None
    ref = Ref()
    cb, queries = make_quantizer_and_queries(nq)
    h = ref.flat(DIM, M, cb, codes_np(0, n_cpu))
    h.prepare(KEEP if n_cpu * KEEP >= 2 * R else (2.0 * R) / n_cpu)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        h.search(queries, 1, R, nthreads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if log:
            log(f"cpu step {it}: {dt:.3f}s")
    h.close()
    t = float(np.mean(times))
    return dict(value=n_cpu * nq / t, seconds_per_step=t)


def recall_check(qadc_b200, torch, dev, stream, nq=400):
    """Recall@100 (recall.hpp:45-54, t = 1: the true nearest neighbour is among the returned ids) on a Deep1B-shaped
    (config 5) database scaled to one GPU and a few CPU-seconds: 1e6 x 96-d clustered vectors, IVF-4096, PQ 16x4
    (sq_dim 6) ENCODED on the GPU, nprobe 64, top-100, keep 5 %.  The CUDA path against (a) the unmodified reference
    scanner fed the same coarse assignment, (b) the reference exactly as shipped, whose find_k_neighbors mis-strides
    for more than 256 cells (neighbors.cpp:64).  Checker use of oracle/_ref, outside every timed region."""
    from oracle.pyoracle import Ref
    if not Ref.available():
        return None
    ref = Ref()
    rng = np.random.default_rng(2025)
    n, dim, m, K, ma, r, keep = 10 ** 6, 96, 16, 4096, 64, 100, 0.05
    centres = rng.standard_normal((2048, dim)).astype(np.float32) * 2.0
    base = (centres[rng.integers(0, 2048, n)] + rng.standard_normal((n, dim)).astype(np.float32)).astype(np.float32)
    cents = base[rng.permutation(n)[:K]].copy()                      # coarse quantizer: a sample (k-means is out of scope)
    ix = qadc_b200.Index(dev.index, stream.cuda_stream)
    ix.set_pq(dim, m, np.zeros((m, 16, dim // m), np.float32))
    ix.set_coarse(cents)
    _, a0 = ix.encode(base[:20000])                                  # residual sample for the codebooks
    resid = base[:20000] - cents[a0]
    cb = np.stack([resid[rng.permutation(20000)[:16], j * 6:(j + 1) * 6] for j in range(m)]).astype(np.float32)
    ix.set_pq(dim, m, cb)
    ix.set_coarse(cents)
    codes, assign = ix.encode(base)
    order = np.argsort(assign, kind="stable")
    offsets = np.zeros(K + 1, np.int64); offsets[1:] = np.cumsum(np.bincount(assign, minlength=K))
    codes_s, labels = codes[order], order.astype(np.uint32)
    truth = rng.integers(0, n, nq)
    q = (base[truth] + 0.7 * rng.standard_normal((nq, dim))).astype(np.float32)
    tb, tq = torch.from_numpy(base).to(dev), torch.from_numpy(q).to(dev)      # exact ground truth (plumbing, not the path)
    gt = torch.cat([torch.cdist(tq[i:i + 100], tb).argmin(1) for i in range(0, nq, 100)]).cpu().numpy()
    del tb, tq
    ix.load_ivf(codes_s, labels, offsets, keep)
    ids, d, cnt = ix.search(q, ma, r)
    tabs = ix.build_tables(q, ma, r)
    ix.close()
    ours = float(np.mean([gt[i] in ids[i][:cnt[i]] for i in range(nq)]))
    # The reference legs run in a child process: the reference's error path is exit(1) ("Max quantization bound too
    # high", db_query_4.cpp:271-274), which must not be able to take the bench line down.
    out = {"shape": f"config-5-shaped, scaled: {n} x {dim}-d clustered vectors encoded on the GPU, IVF-{K}, PQ 16x4, nprobe {ma}, "
                    f"top-{r}, keep {keep * 100:g}%, {nq} queries, ground truth = exact nearest neighbour",
           "recall_at_100": ours}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "recall_in.npz")
        np.savez(path, cb=cb, cents=cents, codes=codes_s, labels=labels, offsets=offsets, q=q, gt=gt, assign=tabs["assign"],
                 meta=np.array([dim, m, ma, r], np.int64), keep=np.float32(keep))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--recall-ref-worker", path], capture_output=True, text=True)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if p.returncode == 0 and lines:
            out.update(json.loads(lines[-1]))
        else:
            out["reference_error"] = (p.stderr or p.stdout)[-300:]
    out["note"] = ("reference_as_shipped uses the reference's own coarse assignment, which mis-strides for K > 256 "
                   "(neighbors.cpp:64); with the same (fixed) assignment the two sides differ only in tie-breaking")
    return out


def recall_ref_worker(path):
    """Child process of recall_check: the unmodified reference (oracle/_ref) on the database the parent built."""
    from oracle.pyoracle import Ref
    z = np.load(path)
    dim, m, ma, r = (int(v) for v in z["meta"])
    cb, cents, codes, labels, offsets, q, gt, assign = (z[k] for k in ("cb", "cents", "codes", "labels", "offsets", "q", "gt", "assign"))
    keep = float(z["keep"])
    nq = q.shape[0]
    ref = Ref()
    h = ref.ivf(dim, m, cb, cents, codes, labels, offsets)
    h.prepare(keep)
    sizes = np.diff(offsets)
    starts = np.where(sizes > 0, np.maximum(1, (sizes.astype(np.float32) * np.float32(keep)).astype(np.int64)), 0)
    hit = used = 0
    for i in range(nq):
        a = assign[i]
        if starts[a].sum() < r:          # the reference would exit(1) on this query
            continue
        t = ref.tables((q[i][None, :] - cents[a]).astype(np.float32), m, cb, blas_form=True)
        keys, vals, sz = h.query_scan(a, t, r)
        hit += int(gt[i] in keys[:sz]); used += 1
    res = {"reference_same_assignment": hit / max(used, 1), "reference_queries": used}
    print(json.dumps(res), flush=True)
    shipped = h.search(q, ma, r, nthreads=os.cpu_count() or 1)   # may exit(1): printed separately, after the first result
    res["reference_as_shipped"] = float(np.mean([gt[i] in shipped["keys"][i][:shipped["sizes"][i]] for i in range(nq)]))
    print(json.dumps(res), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.pyoracle import Ref
    threads = os.cpu_count() or 1
    if not Ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libqadc_ref.so not built"}))
        return
    n_cpu = min(args.n_vectors, 1 << 25)
    nq = max(threads, 32) * 2
    res = cpu_reference_run(n_cpu, nq, threads, args.steps, args.warmup)
    sample = (f"first {n_cpu} vectors of the same synthetic database ({n_cpu * CODE_BYTES >> 20} MiB of codes, larger than the "
              f"host's last-level cache share per thread but far smaller than the 8 GB the GPU arm streams), {nq} queries "
              f"per step, OpenMP over queries around the reference's scanner_4::query_scan; a step of the full "
              f"{args.n_vectors}-vector workload would take {args.n_vectors / n_cpu:.0f}x longer at the same rate")
    cfg = workload_config(args, nq, 1)
    cfg.update(sampled_n_vectors=n_cpu, queries_per_step=nq, sharding="none (host cores)",
               reference_build="unmodified sources, the reference's own flags (CMakeLists.txt:7) except -march=haswell "
                               "instead of -march=native so that the library built in the CPU container runs on this host; "
                               "the scan is explicit AVX2 intrinsics (simd_scan.hpp:125-187)",
               l2="n/a (CPU)")
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "vectors/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": res["value"], "unit": "vectors/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": res["value"], "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, nq, qb):
    return {"workload": f"SIFT1B-shaped flat PQ m=16x4bit ({args.n_vectors} vectors x 8 B), top-{R}, keep {KEEP * 100:g}%",
            "n_vectors": args.n_vectors, "dim": DIM, "m": M, "bits": 4, "r": R, "keep": KEEP, "queries_per_step": nq,
            "queries_per_pass": qb, "sharding": f"flat contiguous x{args.gpus}",
            "l2": "database per GPU >> 126 MB L2 (no flush needed)" if args.n_vectors // args.gpus * 8 > 512e6
                  else "L2 flushed between steps"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import qadc_b200
    from qadc_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, nq = args.n_vectors, args.queries
    cb, queries = make_quantizer_and_queries(nq)

    # one explicit stream for everything (library kernels, torch copies, NCCL, timing events)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ix = qadc_b200.Index(local, stream.cuda_stream)
    ix.set_pq(DIM, M, cb)
    lo, hi = sharding.flat_shard_range(N, rank, world)
    n_local = hi - lo
    ix.begin_database([n_local], False)
    chunk = 1 << 24
    for c0 in range(lo, hi, chunk):
        c1 = min(c0 + chunk, hi)
        t = codes_torch(c0, c1, dev)
        torch.cuda.synchronize(dev)
        ix.upload_codes_device(0, c0 - lo, c1 - c0, t.data_ptr())
        del t
    n_prefix = sharding.start_size(N, KEEP)
    if world > 1:
        ix.set_position_base(0, lo)
        pre = codes_torch(0, n_prefix, dev)
        torch.cuda.synchronize(dev)
        ix.set_prefix_device(0, pre.data_ptr(), n_prefix)
        del pre
    ix.finalize(KEEP)
    if args.qb:
        ix.set_option("flat_qb", args.qb)
    ix.set_option("time_scan", 1)
    if args.flat_filter is not None:
        ix.set_option("flat_filter", args.flat_filter)
    if args.flat_ring is not None:
        ix.set_option("flat_ring", args.flat_ring)

    # device-resident buffers (the `value` leg)
    d_q = torch.from_numpy(queries).to(dev)
    d_ids = torch.empty((nq, R), dtype=torch.int32, device=dev)
    d_d = torch.empty((nq, R), dtype=torch.int8, device=dev)
    d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    d_keys = torch.empty((nq, R), dtype=torch.int64, device=dev)   # the only buffer exchanged: a flat id is the key's low 32 bits
    g_keys = torch.empty((world, nq, R), dtype=torch.int64, device=dev) if world > 1 else None
    o_ids, o_d, o_cnt = torch.empty_like(d_ids), torch.empty_like(d_d), torch.empty_like(d_cnt)
    flush = None
    if n_local * CODE_BYTES <= 512e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    launches = [0]

    def step_device():
        if flush is not None:
            flush.fill_(1)
        ix.search_device(d_q.data_ptr(), nq, 1, R, d_ids.data_ptr(), d_d.data_ptr(), d_cnt.data_ptr(), d_keys.data_ptr())
        launches[0] += ix.last_launch_count()
        if world > 1:
            dist.all_gather_into_tensor(g_keys.view(world * nq, R), d_keys)   # ONE NCCL all-gather per step (nq*r*8 B per rank)
            ix.merge_shards_device(g_keys.data_ptr(), None, world, nq, R, o_ids.data_ptr(), o_d.data_ptr(), o_cnt.data_ptr())
            launches[0] += 1

    # host-buffer leg (`e2e`): pinned queries in, pinned results out, copies inside the timed region
    h_q = torch.from_numpy(queries).pin_memory()
    h_ids = torch.empty((nq, R), dtype=torch.int32).pin_memory()
    h_d = torch.empty((nq, R), dtype=torch.int8).pin_memory()
    h_cnt = torch.empty(nq, dtype=torch.int32).pin_memory()
    h_keys = torch.empty((nq, R), dtype=torch.int64).pin_memory()

    def step_e2e():
        if world == 1:
            ix.search_host_buffers(h_q.data_ptr(), nq, 1, R, h_ids.data_ptr(), h_d.data_ptr(), h_cnt.data_ptr())
        else:
            d_q.copy_(h_q, non_blocking=True)
            step_device()
            h_ids.copy_(o_ids, non_blocking=True); h_d.copy_(o_d, non_blocking=True); h_cnt.copy_(o_cnt, non_blocking=True)
            stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    def per_rank(x):
        """[x of rank 0, ..., x of rank world-1] (diagnostic: skew between the GPUs of a box)"""
        if world == 1:
            return [float(x)]
        t = torch.zeros(world, device=dev)
        t[rank] = float(x)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    for _ in range(args.warmup):
        step_device()
    ix.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches[0] = 0

    prof = os.environ.get("QADC_PROFILE_RANGE") == "1"   # ncu --profile-from-start off: only the timed steps
    if prof:
        torch.cuda.profiler.start()
    ms_step = timed(step_device, args.steps)
    # CUDA events recorded on the launching stream around every scan launch of the timed steps
    scan_ms = ix.scan_ms_history(min(args.steps, 64))
    scan_ms_ranks = per_rank(np.mean(scan_ms))
    if prof:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    n_launch = launches[0]
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    ix.synchronize()
    # diagnostic (N > 1): where a step's time goes on every rank — the local search vs the top-r exchange (NCCL all-gather
    # + shard merge; a rank that finishes its scan early waits here for the slowest GPU of the box)
    exchange = None
    if world > 1:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        acc = [0.0, 0.0]
        for _ in range(5):
            barrier()
            ev[0].record(stream)
            ix.search_device(d_q.data_ptr(), nq, 1, R, d_ids.data_ptr(), d_d.data_ptr(), d_cnt.data_ptr(), d_keys.data_ptr())
            ev[1].record(stream)
            dist.all_gather_into_tensor(g_keys.view(world * nq, R), d_keys)
            ix.merge_shards_device(g_keys.data_ptr(), None, world, nq, R, o_ids.data_ptr(), o_d.data_ptr(), o_cnt.data_ptr())
            ev[2].record(stream)
            stream.synchronize()
            acc[0] += ev[0].elapsed_time(ev[1]) / 5
            acc[1] += ev[1].elapsed_time(ev[2]) / 5
        exchange = {"search_ms_per_rank": per_rank(acc[0]), "allgather_and_merge_ms_per_rank": per_rank(acc[1]),
                    "note": "5 untimed steps after the timed region, each behind a barrier + synchronize: the search time here includes "
                            "the host's launch latency of its short kernels (in the timed loop the launches of step i+1 are queued "
                            "while step i scans), the exchange includes waiting for the slowest rank"}
    # informational: the same step with 4 queries sharing every pass over the codes.  A quarter of
    # the HBM traffic, so the scan is bound by integer issue instead of HBM (and draws less power);
    # not the configuration `value` / `roofline` are quoted on.
    ms_batched = None
    if args.qb == 1 and nq >= 4:
        ix.set_option("flat_qb", 4)
        for _ in range(2):
            step_device()
        ms_batched = timed(step_device, min(args.steps, 10))
        ix.set_option("flat_qb", args.qb)
        ix.synchronize()

    # roofline of the dominant kernel (the 4-bit scan): algorithmic bytes = passes x N_local x 8
    qb_used = args.qb if args.qb else (2 if nq >= 2 else 1)
    passes = -(-nq // qb_used)
    t_scan = float(np.mean(scan_ms)) * 1e-3
    achieved = passes * n_local * CODE_BYTES / t_scan / 1e9
    peak, peak_src = measured_peak_hbm()

    verify = None
    if args.verify:
        # full-size cross-check outside the timed region, on EVERY rank: the (merged) result of `--verify` queries
        # must equal the canonical rule evaluated with numpy on the per-vector distances of an independent kernel
        # (qadc_dump_distances over this rank's shard); with several GPUs the per-shard candidates are gathered and
        # merged on the host, so the device-side all-gather + merge is what is being checked.
        fin_ids, fin_d, fin_c = (o_ids, o_d, o_cnt) if world > 1 else (d_ids, d_d, d_cnt)
        res_ids = fin_ids.cpu().numpy().view(np.uint32); res_d = fin_d.cpu().numpy(); res_c = fin_c.cpu().numpy()
        tabs = ix.build_tables(queries[:args.verify], 1, R)
        cand = np.full((args.verify, R), np.iinfo(np.int64).max, np.int64)   # (distance << 32 | global position), ascending
        for s_ in range(args.verify):
            dd = ix.dump_distances(0, tabs["qtables"][s_, 0])
            thr = res_d[s_][res_c[s_] - 1] if res_c[s_] == R else 126
            pos = np.nonzero(dd <= thr)[0]
            order = np.lexsort((pos, dd[pos]))[:R]
            cand[s_, :len(order)] = (dd[pos[order]].astype(np.int64) << 32) | (pos[order].astype(np.int64) + lo)
            del dd
        if world > 1:
            gc = sharding.all_gather_keys(torch.from_numpy(cand).to(dev)).cpu().numpy()   # [world, verify, R]
        else:
            gc = cand[None]
        okv = True
        for s_ in range(args.verify):
            k = np.sort(gc[:, s_].reshape(-1))[:R]
            k = k[k != np.iinfo(np.int64).max]
            n_ = len(k)
            okv = okv and res_c[s_] == n_ and np.array_equal(res_ids[s_][:n_], (k & 0xffffffff).astype(np.uint32)) \
                and np.array_equal(res_d[s_][:n_].astype(np.int64), k >> 32)
        if world > 1:
            t_ok = torch.tensor([1 if okv else 0], device=dev)
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            okv = bool(t_ok.item())
        verify = {"queries": args.verify, "ok": bool(okv),
                  "method": "canonical top-r recomputed from qadc_dump_distances (independent kernel) + numpy on every shard, "
                            "gathered and merged on the host, compared with the device result on every rank"}

    configs = None
    if not args.no_configs:
        from tools import bench_legs
        sm_mhz = float((clocks or {}).get("sm_max_mhz") or 1965.0)
        configs = {}
        try:
            if world == 1:
                ix.close(); ix = None
                torch.cuda.empty_cache()
                configs["1"] = bench_legs.leg_flat(qadc_b200, torch, dev, stream, "1: SIFT1M-shaped flat PQ 16x4, 10k queries, top-100",
                                                   10 ** 6, 128, 16, 0.01, 10000, (1, 2, 4), 1235, sm_mhz)
                configs["2"] = bench_legs.leg_ivf(qadc_b200, torch, dev, stream, "2: SIFT1M-shaped IVF-4096 residual PQ 16x4, nprobe 64, 10k queries, top-100",
                                                  10 ** 6, 128, 16, 4096, 64, 0.01, 10000, 1236)
                configs["3"] = bench_legs.leg_flat(qadc_b200, torch, dev, stream, "3: Deep10M-shaped flat PQ 32x4 (96-d), batched 10k queries, top-100",
                                                   10 ** 7, 96, 32, 0.001, 10000, (1, 2), 1237, sm_mhz, check=2)
            else:
                ix.close(); ix = None
                torch.cuda.empty_cache()
                n5 = 10 ** 9 if world == 8 else 125 * 10 ** 6 * world   # the full Deep1B shape on 8 GPUs, the same per-GPU share otherwise
                configs["5"] = bench_legs.leg_config5_sharded(qadc_b200, torch, dist, sharding, codes_torch, codes_torch_at, dev, stream,
                                                              rank, world, n5, 10000)
        except Exception as e:   # an informational leg must never take the headline down
            configs["error"] = repr(e)[:500]

    if rank == 0:
        traffic, traffic_src = ncu_traffic_per_launch(n_local, nq, qb_used)
        value = N * nq / (ms_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "vectors/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": workload_config(args, nq, qb_used),
            "queries_per_s": nq / (ms_step * 1e-3),
            "e2e": {"value": N * nq / (ms_e2e * 1e-3), "unit": "vectors/s", "h2d_bytes_per_step": int(queries.nbytes),
                    "d2h_bytes_per_step": int(nq * R * 5 + nq * 4), "ms_per_step": ms_e2e},
            "gpu_launches": n_launch,
            "batched": None if ms_batched is None else {
                "queries_per_pass": 4, "value": N * nq / (ms_batched * 1e-3), "unit": "vectors/s", "ms_per_step": ms_batched,
                "note": "informational: same step, 4 queries share each pass over the codes (integer-issue bound)"},
            "verify": verify,
            "exchange": exchange,
            "configs": configs,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "scan_flat_wr_kernel<16,4> (one query per pass, per-warp TMA rings)" if qb_used == 1 else "scan_flat_wrq_kernel (several queries per pass)",
                         "kernel_ms": t_scan * 1e3, "kernel_ms_per_rank": scan_ms_ranks, "kernel_share_of_step": t_scan * 1e3 / ms_step,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": passes * n_local * CODE_BYTES},
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            n_cpu = min(N, 1 << 24)
            nq_cpu = max(2 * threads, 32)
            res = cpu_reference_run(n_cpu, nq_cpu, threads, 2, 1)
            if res:
                line["cpu_baseline"] = {"value": res["value"], "unit": "vectors/s", "cores": threads, "kind": "reference",
                                        "sample": f"reference scanner_4 (AVX2, oracle/_ref) on the first {n_cpu} vectors of "
                                                  f"the same database, {nq_cpu} queries, OpenMP over queries"}
            if not args.no_configs:
                try:
                    line["recall_check"] = recall_check(qadc_b200, torch, dev, stream)
                except Exception as e:
                    line["recall_check"] = {"error": repr(e)[:300]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if ix is not None:
        ix.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-vectors", type=int, default=int(os.environ.get("QADC_BENCH_N", 10 ** 9)))
    ap.add_argument("--queries", type=int, default=16)
    ap.add_argument("--qb", type=int, default=1, help="queries per pass of the flat scan (0 = library default)")
    ap.add_argument("--flat-filter", type=int, default=None, help="0: exact lookup core only (A/B against the pre-filter)")
    ap.add_argument("--flat-ring", type=int, default=None, help="1: per-warp TMA rings (scan_flat_wr_kernel), 0: CTA-wide ring")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--verify", type=int, default=2, help="queries cross-checked at full size after timing (every N)")
    ap.add_argument("--no-configs", action="store_true", help="skip the informational `configs` legs (BASELINE configs 1-3 / 5)")
    ap.add_argument("--recall-ref-worker", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.recall_ref_worker:
        recall_ref_worker(args.recall_ref_worker)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
