"""The count kernel's one-stage remainder (neurokmer_b200/csrc/nk_device.cuh: fastmod_single_dev) re-derived with exact
integer arithmetic on the CPU: for every pool size the host picks a split (`kind`), and for adversarial hashes the
device's three floating-point steps — x = fma(A, m, B); Q = fma_rz(x, inv_dn, 2^52); r = fma(-q, p, x) — are emulated
with Python integers / exact binary fractions (an IEEE double is mantissa * 2^exponent), so the claims in the header
comment are checked, not assumed:  x < 2^53 (exact), q in {floor(x/p) - 1, floor(x/p)}, 0 <= r < 2p <= 2^32, and
min(r, r - p) on unsigned 32-bit words == h mod p.  (The GPU test test_exact_modulo_all_pool_sizes runs the real kernel.)"""
import math
import struct
from fractions import Fraction

import numpy as np
import pytest


def next_down(x: float) -> float:
    (u,) = struct.unpack("<Q", struct.pack("<d", x))
    return struct.unpack("<d", struct.pack("<Q", u - 1))[0]


def plan(p: int):
    """make_fastmod's choice (nk_device.cuh): kind 3 = split at bit 32, kind 2 = split at bit 44, 0 = two-stage form"""
    if p & (p - 1) == 0 or p < 4 or p > 2**31:
        return 0, 0, 0
    m32, m44 = 2**32 % p, 2**44 % p
    if m32 < 2**21 - 1:
        return 3, 32, m32
    return 2, 44, m44


def device_steps(h: int, p: int):
    kind, s, m = plan(p)
    assert kind in (2, 3)
    A, B = h >> s, h & ((1 << s) - 1)
    # kind 2 takes A in place as A * 2^12 and multiplies by m * 2^-12: the product is the same integer
    x = A * m + B
    assert x < 2**53                       # one fma is exact
    inv_dn = next_down(1.0 / p)            # <= 1/p
    assert Fraction(inv_dn) <= Fraction(1, p)
    # fma_rz(x, inv_dn, 2^52): the exact product plus 2^52, truncated to an integer (ulp is 1 in [2^52, 2^53))
    Q = math.floor(Fraction(x) * Fraction(inv_dn) + 2**52)
    assert 2**52 <= Q < 2**53
    q = Q - 2**52
    assert q in (x // p, x // p - 1)
    r = x - q * p
    assert 0 <= r < 2 * p <= 2**32
    ri = r & 0xFFFFFFFF
    return min(ri, (ri - p) & 0xFFFFFFFF)


POOLS = [4, 5, 6, 7, 10, 1000, 65535, 65537, 999_983, 1_000_000, 2_000_000, 2**21 - 3, 2**21 - 1, 2**21 + 1, 15_625, 16_000_000,
         2**24 - 1, 1_431_655_766, 2**31 - 1, 2**31 - 19, 3 * 2**29 + 1, 2**30 + 1, 2**31]


@pytest.mark.parametrize("p", POOLS)
def test_one_stage_remainder_is_exact(p):
    kind, s, m = plan(p)
    if kind == 0:
        assert p & (p - 1) == 0      # powers of two take the AND; everything else in the list has a one-stage plan
        return
    rng = np.random.default_rng(p % 2**32)
    hs = [0, 1, p - 1, p, p + 1, 2**32 - 1, 2**32, 2**44 - 1, 2**44, 2**63, 2**64 - 1, 2**64 - 2, (2**64 - 1) // p * p,
          (2**64 - 1) // p * p - 1, 0xFFFFFFFF00000000, 0xFFFFFFFFFFFFF000, 0xFFFFF00000000000 | (p - 1)]
    hs += [int(x) for x in rng.integers(0, 2**64, size=300, dtype=np.uint64)]
    qs = rng.integers(0, (2**64 - 1) // p, size=300, dtype=np.uint64)
    hs += [int(q) * p for q in qs] + [int(q) * p + p - 1 for q in qs] + [int(q) * p + 1 for q in qs]
    # values whose x = A*m + B is an exact multiple of p, or one below (the quotient estimate's worst cases)
    for A in (0, 1, (1 << (64 - s)) - 1, int(rng.integers(0, 1 << (64 - s)))):
        am = A * m
        for t in (0, 1, 2, 3):
            B = (-am) % p + t * p
            for d in (0, -1, 1):
                b = B + d
                if 0 <= b < (1 << s):
                    hs.append((A << s) | b)
    for h in hs:
        assert device_steps(h, p) == h % p, (h, p)


def test_plan_covers_every_pool_size_class():
    assert plan(2_000_000)[0] == 3 and plan(16_000_000)[0] == 2 and plan(2**20)[0] == 0 and plan(3)[0] == 0
    assert plan(2**31 + 1)[0] == 0 and plan(2**31 - 1)[0] in (2, 3)
    for p in range(4, 3000):
        k, s, m = plan(p)
        if p & (p - 1):
            assert k == 3 and m == 2**32 % p      # every p < 2^21 splits at bit 32
