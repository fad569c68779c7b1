#!/usr/bin/env python
"""Device-resident timings of the BASELINE.json config SHAPES other than the bench workload
(configs[0], [2], [3]; scaled where noted) — kernel phases from the library's CUDA events.
Run under gpurun:  python tools/bench_configs.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import SpikingKmerCounter  # noqa: E402
from neurokmer_b200.devmem import copy_h2d  # noqa: E402


def run(name, k, pool, lens, seed, flags, reps=5, canonical=True, exact=False):
    import time
    lens = np.asarray(lens, np.int64)
    n = int(lens.sum())
    c = SpikingKmerCounter(k, 1.0, 0.95, 2, 1.0, pool, canonical)
    if exact:
        c.enable_exact_counts(True)
    db, do = c.stage_reserve(n, lens.size)
    c.synth_fill(db, seed, 0, n, flags)
    offs = np.zeros(lens.size + 1, np.uint64); offs[1:] = np.cumsum(lens)
    copy_h2d(do, offs); c.synchronize()
    best = None
    for _ in range(reps):
        c.synchronize(); t0 = time.perf_counter()
        c.reset(); c.stream_begin(); c.process_staged(n, lens.size, 1); c.stream_finish(); top = c.top_abundant_neurons(20)
        wall = (time.perf_counter() - t0) * 1e3
        t = c.timings(); t["wall_ms"] = wall
        if best is None or t["count_ms"] < best["count_ms"]:
            best = t
    kmers = best["kmers"]
    out = dict(config=name, bases=n, kmers=kmers, k=k, pool=pool,
               mark_ms=round(best["mark_ms"], 4), count_ms=round(best["count_ms"], 4), post_ms=round(best["lif_ms"], 4),
               count_gkmers_s=round(kmers / best["count_ms"] / 1e6, 2), job_wall_ms=round(best["wall_ms"], 3),
               total_spikes=c.energy.total_spikes(), top1=top[0])
    print(json.dumps(out), flush=True)
    c.close()


if __name__ == "__main__":
    run("config1: 10 Mbp, k=21, pool 1M", 21, 1_000_000, [10_000_000], 1, 0)
    run("config2: 113 Mbp, k=31, pool 2M (bench workload)", 31, 2_000_000, [30e6, 25e6, 20e6, 15e6, 10e6, 8e6, 5e6], 2, 3)
    c2 = [30e6, 25e6, 20e6, 15e6, 10e6, 8e6, 5e6]
    run("f4: config2 NON-canonical (pack_kmer path; N runs take the skip rule)", 31, 2_000_000, c2, 2, 3, canonical=False)
    run("f1: config2 with the exact side tables on (words appended, sorted, run-length encoded)", 31, 2_000_000, c2, 2, 3, reps=3, exact=True)
    nreads = 10_000_000
    lens = np.full(nreads, 150, np.int64); lens[::1000] = 20
    run("config3 shape: 10 M reads x 150 bp (1/5 of the 50 M), k=31, pool 2M", 31, 2_000_000, lens, 3, 0)
    run("config4: 1 Gbp, k=15, pool 65536 (contention)", 15, 65536, [1_000_000_000], 4, 0)
    run("config5 shard shape: 1.25 Gbp (1/8 of 10 Gbp), k=31, pool 16M", 31, 16_000_000, [100_000_000] * 12 + [50_000_000], 5, 1)
    if "--full" in sys.argv:
        nreads = 50_000_000
        lens = np.full(nreads, 150, np.int64); lens[::1000] = 20
        run("config3 FULL: 50 M reads x 150 bp (7.5 Gbase, two launches of < 2^32 starts), k=31, pool 2M",
            31, 2_000_000, lens, 3, 0, reps=2)
