"""Exact side tables on the bench workload: time of one whole job with nk_enable_exact_counts
(count kernel appending (word, index) + bucket partition + per-bucket shared-memory dedup) against the plain job."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from neurokmer_b200 import SpikingKmerCounter  # noqa: E402
from neurokmer_b200.devmem import copy_h2d  # noqa: E402

c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, 2_000_000, True)
n = bench.NBASES
offsets = np.concatenate([[0], np.cumsum(bench.SEQ_LENS)]).astype(np.uint64)
db, do = c.stage_reserve(n, 7)
c.synth_fill(db, 2, 0, n, int(os.environ.get('NK_SYNTH_FLAGS', 3))); copy_h2d(do, offsets); c.synchronize()
for exact in (False, True):
    c.enable_exact_counts(exact)
    ts = []
    for it in range(8):
        c.reset()
        t0 = time.perf_counter()
        c.process_staged(n, 7, 0)
        spikes = c.energy.total_spikes()   # observes the result: the job is complete
        ts.append((time.perf_counter() - t0) * 1e3)
    extra = ""
    if exact:
        keys, counts = c.exact_table()
        extra = f", {keys.size} distinct k-mers, {int(counts.sum())} windows, max count {int(counts.max())}, uniques sum {int(c.kmer_per_neuron().sum())}"
    print(f"exact={exact}: whole job {min(ts[2:]):.3f} ms (wall, best of {len(ts) - 2}){extra}", flush=True)
