#!/usr/bin/env python
"""Throughput of the plain ADC path (db_query, "next" row N4) on one B200: SIFT1M-shaped flat
database, PQ 8x8 (the reference README's ADC baseline configuration: 2594 us/query on one CPU
thread, README.md:275-278), top-100.  Not a bench line.  usage: tools/bench_adc.py [m bits]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qadc_b200  # noqa: E402


def main():
    m, bits = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 8)
    n, dim, nq, r = 10 ** 6, 128, 2000, 100
    rng = np.random.default_rng(4)
    cb = rng.standard_normal((m, 1 << bits, dim // m)).astype(np.float32)
    codes = rng.integers(0, 256, (n, m * bits // 8), dtype=np.uint8)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    ix = qadc_b200.Index(0)
    ix.set_pq(dim, m, cb, bits=bits)
    ix.adc_load(codes)
    ix.adc_search(q[:64], 1, r)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        ids, d, cnt = ix.adc_search(q, 1, r)
        best = min(best, time.perf_counter() - t0)
    assert np.all(cnt == r) and np.all(np.diff(d, axis=1) >= 0)
    print(json.dumps({"config": "plain ADC, SIFT1M-shaped flat %dx%d, top-%d, %d queries" % (m, bits, r, nq),
                      "seconds": best, "us_per_query": best / nq * 1e6, "queries_per_s": nq / best,
                      "vectors_scanned_per_s": nq * n / best, "launches": ix.last_launch_count()}))
    ix.close()


if __name__ == "__main__":
    main()
