#!/usr/bin/env python
"""Second job on a non-reset counter (carried neuron state): memoised LIF (one simulation per distinct
(state, count) key) against the direct kernel (one per neuron).  Run under gpurun."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import SpikingKmerCounter  # noqa: E402
from neurokmer_b200.devmem import copy_h2d  # noqa: E402

for pool, lens, seed in ((2_000_000, [30e6, 25e6, 20e6, 15e6, 10e6, 8e6, 5e6], 2), (16_000_000, [100e6] * 5, 5)):
    lens = np.array(lens, np.int64); n = int(lens.sum())
    offs = np.zeros(lens.size + 1, np.uint64); offs[1:] = np.cumsum(lens)
    for mode, name in ((0, "memoised"), (1, "direct")):
        c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, pool, True)
        c.debug_set_lif_path(mode)
        db, do = c.stage_reserve(n, lens.size); c.synth_fill(db, seed, 0, n, 3); copy_h2d(do, offs); c.synchronize()
        rows = []
        for job in range(4):
            c.stream_begin(); c.process_staged(n, lens.size, 1); c.stream_finish(); c.synchronize()
            t = c.timings()
            rows.append((t["lif_path"], round(t["lif_ms"], 4)))
        print(json.dumps(dict(pool=pool, mode=name, jobs_path_ms=rows, total_spikes=c.energy.total_spikes())), flush=True)
        c.close()
