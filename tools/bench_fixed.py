#!/usr/bin/env python
"""Per-step fixed cost of the flat search (everything that is not the 4-bit scan kernel) as ONE of G GPUs sees it:
rank 0's contiguous shard of the 1e9 x 16x4 bench database with the replicated keep-prefix (500 000 vectors), 16 queries
per step, one query per pass.  The scan shrinks with G, the table / keep-prefix / bound / merge stages do not: at G = 8
they are what separates the measured scaling from linear.

usage: [G=8] [STEPS=40] [OPT=flat_prep] [QADC_LIB=build_ab/x.so] python tools/bench_fixed.py [out.npz]
Prints one line per setting (1, 0, 1, 0) of the option OPT (skipped for libraries that do not know it) and, when a path is
given, stores the result arrays so that two builds can be compared bit for bit (tools/bench_fixed.py a.npz b.npz = compare)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) == 3:
    a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
    same = all(np.array_equal(a[k], b[k]) for k in a.files)
    print("results identical:", same, sorted(a.files))
    sys.exit(0 if same else 1)

import torch
import bench
import qadc_b200
from qadc_b200 import sharding

G = int(os.environ.get("G", "8"))
STEPS = int(os.environ.get("STEPS", "40"))
N, nq, R = 10 ** 9, 16, bench.R
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
cb, queries = bench.make_quantizer_and_queries(nq)
ix = qadc_b200.Index(0, stream.cuda_stream)
ix.set_pq(bench.DIM, bench.M, cb)
lo, hi = sharding.flat_shard_range(N, 0, G)
n_local = hi - lo
ix.begin_database([n_local], False)
for c0 in range(lo, hi, 1 << 24):
    c1 = min(c0 + (1 << 24), hi)
    t = bench.codes_torch(c0, c1, dev)
    torch.cuda.synchronize(dev)
    ix.upload_codes_device(0, c0 - lo, c1 - c0, t.data_ptr())
    del t
n_prefix = sharding.start_size(N, bench.KEEP)
if G > 1:
    ix.set_position_base(0, lo)
    pre = bench.codes_torch(0, n_prefix, dev)
    torch.cuda.synchronize(dev)
    ix.set_prefix_device(0, pre.data_ptr(), n_prefix)
    del pre
ix.finalize(bench.KEEP)
ix.set_option("flat_qb", 1)
ix.set_option("time_scan", 1)
d_q = torch.from_numpy(queries).to(dev)
d_ids = torch.empty((nq, R), dtype=torch.int32, device=dev)
d_d = torch.empty((nq, R), dtype=torch.int8, device=dev)
d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
d_keys = torch.empty((nq, R), dtype=torch.int64, device=dev)


def step():
    ix.search_device(d_q.data_ptr(), nq, 1, R, d_ids.data_ptr(), d_d.data_ptr(), d_cnt.data_ptr(), d_keys.data_ptr())


out = {}
OPT = os.environ.get("OPT", "flat_prep")
for seed in (1, 0, 1, 0):
    try:
        ix.set_option(OPT, seed)
    except qadc_b200.QadcError:
        if seed == 0:
            continue
    for _ in range(5):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(STEPS):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / STEPS
    scan = float(np.mean(ix.scan_ms_history(min(STEPS, 64))))
    print(f"G={G} n_local={n_local} prefix={n_prefix} {OPT}={seed}: step {ms:.4f} ms, scan kernel {scan:.4f} ms, "
          f"fixed {1e3 * (ms - scan):.1f} us, launches/step {ix.last_launch_count()}", flush=True)
    out[f"ids_seed{seed}"] = d_ids.cpu().numpy()
    out[f"d_seed{seed}"] = d_d.cpu().numpy()
    out[f"cnt_seed{seed}"] = d_cnt.cpu().numpy()
if "ids_seed0" in out:
    same = all(np.array_equal(out[f"{k}_seed0"], out[f"{k}_seed1"]) for k in ("ids", "d", "cnt"))
    print(f"{OPT} 0 == {OPT} 1:", same, flush=True)
if len(sys.argv) == 2:
    np.savez(sys.argv[1], ids=out["ids_seed1"], d=out["d_seed1"], cnt=out["cnt_seed1"])
ix.close()
