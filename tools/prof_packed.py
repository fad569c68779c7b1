#!/usr/bin/env python
"""Two jobs of the bench workload through nk_stream_push_packed from pinned memory (zero-copy body):
the only count_kernel launches of this process are the PACKED instantiation, so that
`ncu -k regex:count_kernel --launch-skip 2 -c 1` captures the second job's body launch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import PinnedBuffer, SpikingKmerCounter, pack_bases  # noqa: E402
from oracle.synth import synth_bases  # noqa: E402  (host twin of the device generator: input only)

LENS = np.array([30e6, 25e6, 20e6, 15e6, 10e6, 8e6, 5e6], np.int64)
N = int(LENS.sum())
offs = np.zeros(LENS.size + 1, np.uint64); offs[1:] = np.cumsum(LENS)
bases = synth_bases(2, 0, N, 3)
codes = PinnedBuffer(4 * ((N + 15) // 16), np.uint32); other = PinnedBuffer(4 * ((N + 31) // 32), np.uint32)
pack_bases(bases, out_codes=codes.array, out_other=other.array)
c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, 2_000_000, True)
for _ in range(3):
    c.reset(); c.stream_begin(); c.stream_push_packed(codes.array, other.array, offs); c.stream_end()
    print(c.energy.total_spikes(), c.timings()["count_ms"])
