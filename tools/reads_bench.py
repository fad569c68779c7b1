#!/usr/bin/env python3
"""Short-read batches (BASELINE configs[2] shape: 150 bp reads) through the count kernel's compaction mode (mode 3):
device-resident, count_ms from CUDA events, L2 flushed before every launch.  Prints k-mers/s next to the long-sequence
number of the same number of bases."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import SpikingKmerCounter  # noqa: E402
from neurokmer_b200.devmem import copy_h2d  # noqa: E402


def run(read_len, nreads, k=31, pool=2_000_000):
    nb = read_len * nreads
    offsets = (np.arange(nreads + 1, dtype=np.uint64) * np.uint64(read_len))
    c = SpikingKmerCounter(k, 1.0, 0.95, 2, 1.0, pool, True)
    db, do = c.stage_reserve(nb, nreads)
    c.synth_fill(db, 3, 0, nb, 0)
    copy_h2d(do, offsets)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(c.cuda_stream())
    ms = []
    for it in range(13):
        with torch.cuda.stream(stream):
            flush.fill_(it & 0xFF)
        c.reset(); c.stream_begin(); c.process_staged(nb, nreads, 1); c.stream_finish(); c.synchronize()
        if it >= 3:
            ms.append(c.timings()["count_ms"])
    kmers = c.timings()["kmers"]
    c.close()
    return float(np.mean(ms)), kmers


if __name__ == "__main__":
    for read_len, nreads in ((150, 800_000), (100, 1_200_000), (250, 480_000), (120_000_000, 1)):
        ms, kmers = run(read_len, nreads)
        print(f"{nreads} reads x {read_len} bp: count {ms:.4f} ms, {kmers} k-mers, {kmers / ms / 1e6:.1f} G k-mers/s "
              f"({read_len * nreads / ms / 1e6:.1f} G bases/s)", flush=True)
