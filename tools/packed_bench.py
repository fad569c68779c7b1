#!/usr/bin/env python
"""Pre-packed input path on the bench workload (configs[1]): device-resident count-kernel rate with
packed input, and the end-to-end time of nk_stream_push_packed from pinned host memory for several
H2D granules (NK_PACKED_CHUNK_MBASES).  Run under gpurun:  python tools/packed_bench.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import PinnedBuffer, SpikingKmerCounter, pack_bases  # noqa: E402
from neurokmer_b200.devmem import copy_h2d, device_to_numpy  # noqa: E402

K, POOL = 31, 2_000_000
LENS = np.array([30e6, 25e6, 20e6, 15e6, 10e6, 8e6, 5e6], np.int64)
N = int(LENS.sum())

c = SpikingKmerCounter(K, 1.0, 0.95, 2, 1.0, POOL, True)
db, do = c.stage_reserve(N, LENS.size)
c.synth_fill(db, 2, 0, N, 3)
offs = np.zeros(LENS.size + 1, np.uint64); offs[1:] = np.cumsum(LENS)
copy_h2d(do, offs); c.synchronize()
pinned = PinnedBuffer(N); pinned.array[:] = device_to_numpy(db, N)
codes = PinnedBuffer(4 * ((N + 15) // 16), np.uint32); other = PinnedBuffer(4 * ((N + 31) // 32), np.uint32)
for th in (1, 4, 8, 16, 0):
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); _, _, n_other = pack_bases(pinned.array, threads=th, out_codes=codes.array, out_other=other.array)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps(dict(host_pack_threads=th or os.cpu_count(), ms=round(best * 1e3, 3), gb_s=round(N / best / 1e9, 1), n_other=n_other)), flush=True)


def resident(packed, has_other=True, reps=7):
    best = None
    for _ in range(reps):
        c.synchronize(); t0 = time.perf_counter()
        c.reset(); c.stream_begin()
        if packed:
            c.process_staged_packed(N, LENS.size, 1, has_other)
        else:
            c.process_staged(N, LENS.size, 1)
        c.stream_finish(); top = c.top_abundant_neurons(20)
        wall = (time.perf_counter() - t0) * 1e3
        t = c.timings(); t["wall_ms"] = wall
        if best is None or t["count_ms"] < best["count_ms"]:
            best = t
    return best, top, c.energy.total_spikes()


b, top_a, sp_a = resident(False)
print(json.dumps(dict(input="ASCII resident", count_ms=round(b["count_ms"], 4), gkmers_s=round(b["kmers"] / b["count_ms"] / 1e6, 1), wall_ms=round(b["wall_ms"], 3))), flush=True)
dc, dx, do2 = c.stage_reserve_packed(N, LENS.size)
copy_h2d(dc, codes.array); copy_h2d(dx, other.array); copy_h2d(do2, offs); c.synchronize()
b, top_p, sp_p = resident(True)
assert top_p == top_a and sp_p == sp_a
print(json.dumps(dict(input="packed resident (codes + other)", count_ms=round(b["count_ms"], 4), gkmers_s=round(b["kmers"] / b["count_ms"] / 1e6, 1), wall_ms=round(b["wall_ms"], 3))), flush=True)


def e2e(job, reps=30, check=True):
    ts = []
    for _ in range(reps):
        c.synchronize(); t0 = time.perf_counter()
        c.reset(); c.stream_begin(); job(); c.stream_finish(); top = c.top_abundant_neurons(20)
        ts.append((time.perf_counter() - t0) * 1e3)
    assert not check or top == top_a
    return float(np.median(ts)), float(np.min(ts))


med, mn = e2e(lambda: c.stream_push(pinned.array, offs))
print(json.dumps(dict(e2e="ASCII pinned", median_ms=round(med, 3), min_ms=round(mn, 3))), flush=True)
os.environ["NK_ZEROCOPY"] = "1"
med, mn = e2e(lambda: c.stream_push_packed(codes.array, other.array, offs))
t = c.timings()
print(json.dumps(dict(e2e="packed pinned, zero-copy body (kernel TMA-loads host memory)", median_ms=round(med, 3), min_ms=round(mn, 3),
                      count_ms=round(t["count_ms"], 4), launches=t["launches"])), flush=True)
os.environ["NK_ZEROCOPY"] = "0"
for mb in (16, 32, 64):
    os.environ["NK_PACKED_CHUNK_MBASES"] = str(mb)
    med, mn = e2e(lambda: c.stream_push_packed(codes.array, other.array, offs))
    med2, mn2 = e2e(lambda: c.stream_push_packed(codes.array, None, offs), check=False)
    print(json.dumps(dict(e2e="packed pinned", chunk_mbases=mb, median_ms=round(med, 3), min_ms=round(mn, 3),
                          no_other_median_ms=round(med2, 3), note="no_other: other=NULL (wrong results on this input, timing only)")), flush=True)
