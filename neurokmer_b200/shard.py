"""Sequence-chunk sharding for multi-GPU runs (host logic, no device code).

The reference parallelises over whole sequences (rayon `par_iter`, src/spiking_hash.rs:94-95;
one channel message per sequence, :415).  A k-mer is a pure function of its own k bytes, so a
batch can instead be cut at ANY base: the shard that owns window starts [a, b) reads bytes
[a, min(b + k - 1, end_of_sequence)).  Every window is counted by exactly one rank; per-rank
u64 currents are then summed with one all-reduce (integer sum: order-independent, bit-exact).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_bounds(nbytes: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Window-start range [lo, hi) of `rank` over the concatenated batch (balanced by bytes)."""
    per = -(-nbytes // world)
    if align > 1:
        per = -(-per // align) * align
    lo = min(nbytes, rank * per)
    hi = min(nbytes, lo + per)
    return lo, hi


def shard_batch(bases: np.ndarray, offsets: np.ndarray, k: int, world: int, rank: int) -> Tuple[np.ndarray, np.ndarray]:
    """(bases_r, offsets_r) such that the windows of all ranks partition the windows of the batch.

    Sequence i = bases[offsets[i]:offsets[i+1]].  A sequence cut by a shard boundary contributes
    the piece [max(start, lo), min(end, hi + k - 1)) to the shard owning starts [lo, hi)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    nseq = offsets.size - 1
    nbytes = int(offsets[-1]) if nseq >= 0 and offsets.size else 0
    lo, hi = shard_bounds(nbytes, world, rank)
    if hi <= lo or nseq <= 0:
        return np.zeros(0, np.uint8), np.zeros(1, np.uint64)
    starts, ends = offsets[:-1].astype(np.int64), offsets[1:].astype(np.int64)
    # sequences with end > lo and start < hi
    first = int(np.searchsorted(ends, lo, side="right"))
    last = int(np.searchsorted(starts, hi, side="left"))
    p0 = np.maximum(starts[first:last], lo)
    p1 = np.minimum(ends[first:last], hi + k - 1)
    keep = p1 > p0
    p0, p1 = p0[keep], p1[keep]
    lens = p1 - p0
    out_off = np.zeros(lens.size + 1, np.uint64)
    out_off[1:] = np.cumsum(lens)
    if lens.size == 0:
        return np.zeros(0, np.uint8), out_off
    # pieces are consecutive in memory except for the k-1 overlap of the last one, so one slice + fix-up
    out = np.empty(int(out_off[-1]), np.uint8)
    for i in range(lens.size):
        out[int(out_off[i]):int(out_off[i + 1])] = bases[int(p0[i]):int(p1[i])]
    return out, out_off
