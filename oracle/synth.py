"""numpy twin of the library's synthetic base generator (neurokmer_b200/csrc/nk_misc.cu,
`synth_kernel`).  TEST INFRASTRUCTURE: lets the CPU oracle / reference arm
materialise the same position-addressable stream without a GPU.

    base(p) = "ACGT"[(splitmix64(seed*G + (p>>5)) >> 2(p&31)) & 3]
    flags bit1: lower-case if splitmix64((seed ^ 0x6C6F7765)*G + (p>>12)) % 100 == 0
    flags bit0: 'N' inside one run of 100..10000 bases per 2^20-base block
"""
from __future__ import annotations

import numpy as np

G = np.uint64(0x9E3779B97F4A7C15)
M64 = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + G).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _mul(a: int, b: int) -> np.uint64:
    return np.uint64((a * b) & M64)


def synth_bases(seed: int, start: int, n: int, flags: int = 0, block: int = 1 << 24) -> np.ndarray:
    out = np.empty(n, np.uint8)
    lut = np.frombuffer(b"ACGT", np.uint8)
    with np.errstate(over="ignore"):
        for b0 in range(0, n, block):
            m = min(block, n - b0)
            p = np.arange(start + b0, start + b0 + m, dtype=np.uint64)
            h = splitmix64(_mul(seed, int(G)) + (p >> np.uint64(5)))
            c = lut[((h >> (np.uint64(2) * (p & np.uint64(31)))) & np.uint64(3)).astype(np.intp)].copy()
            if flags & 2:
                hl = splitmix64(_mul(seed ^ 0x6C6F7765, int(G)) + (p >> np.uint64(12)))
                c[(hl % np.uint64(100)) == 0] |= 0x20
            if flags & 1:
                blk = p >> np.uint64(20)
                isn = np.zeros(m, bool)
                for d in (0, 1):
                    ok = blk >= np.uint64(d)
                    b = blk - np.uint64(d)
                    hh = splitmix64(_mul(seed ^ 0x4E52554E, int(G)) + b)
                    s = (b << np.uint64(20)) + (hh & np.uint64(0xFFFFF))
                    ln = np.uint64(100) + ((hh >> np.uint64(20)) % np.uint64(9901))
                    isn |= ok & (p >= s) & (p < s + ln)
                c[isn] = ord("N")
            out[b0:b0 + m] = c
    return out
