#!/usr/bin/env python3
"""What does each part of the count kernel cost?  Device-resident 113 Mbase job (the bench workload), count_ms of
`nk_last_timings` (CUDA events around the kernel), L2 flushed before every launch.

    python tools/count_ablate.py            # rows: input form x pool kind x canonical
    NEUROKMER_LIB=.../lib_nored.so python tools/count_ablate.py

pool 2,000,000 = the general modulo (13 FP64 ops per k-mer); pool 2,097,152 = a power of two (one AND);
packed = the pre-packed "nk2" input (no ASCII classification in the kernel).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import SpikingKmerCounter, pack_bases  # noqa: E402
from neurokmer_b200.devmem import copy_h2d  # noqa: E402

LENS = [20_000_000] * 5 + [10_000_000, 3_000_000]
STEPS = int(os.environ.get("NK_STEPS", 20))


def run(pool, canonical, packed, synth=None):
    nb, nseq = sum(LENS), len(LENS)
    offsets = np.concatenate([[0], np.cumsum(LENS)]).astype(np.uint64)
    c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, pool, canonical)
    rng = np.random.default_rng(7)
    host = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, nb, dtype=np.uint8)]
    if packed:
        codes, other, _ = pack_bases(host)
        dc, dx, do = c.stage_reserve_packed(nb, nseq)
        copy_h2d(dc, codes); copy_h2d(dx, other); copy_h2d(do, offsets)
    else:
        db, do = c.stage_reserve(nb, nseq)
        if synth is None:
            copy_h2d(db, host)
        else:
            c.synth_fill(db, 2, 0, nb, synth)   # the bench's generator: bit0 = runs of N (0.5 %), bit1 = lower-case blocks
        copy_h2d(do, offsets)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(c.cuda_stream())
    ms = []
    for it in range(3 + STEPS):
        with torch.cuda.stream(stream):
            flush.fill_(it & 0xFF)
        c.reset(); c.stream_begin()
        if packed:
            c.process_staged_packed(nb, nseq, 1, False)
        else:
            c.process_staged(nb, nseq, 1)
        c.stream_finish()
        c.synchronize()
        if it >= 3:
            ms.append(c.timings()["count_ms"])
    k = c.timings()["kmers"]
    c.close()
    return float(np.mean(ms)), float(np.min(ms)), k


def main():
    print("lib:", os.environ.get("NEUROKMER_LIB", "(in-tree)"))
    if os.environ.get("NK_ABLATE", "synth") == "synth":
        for flags in (0, 1, 2, 3):
            mean, mn, k = run(2_000_000, True, False, synth=flags)
            print(f"synthetic stream, flags {flags} (bit0: N runs, bit1: lower case), pool 2000000 canonical: "
                  f"count {mean:.4f} ms (min {mn:.4f})  kmers {k}")
        return
    for packed in (False, True):
        for pool in (2_000_000, 2_097_152):
            for canonical in (True, False):
                mean, mn, k = run(pool, canonical, packed)
                print(f"{'packed' if packed else 'ascii ':6s} pool {pool:>9d} {'canonical' if canonical else 'forward  '}: "
                      f"count {mean:.4f} ms (min {mn:.4f})  kmers {k}")


if __name__ == "__main__":
    main()
