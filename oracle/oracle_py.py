"""Pure-Python / numpy twin of oracle/nk_oracle.c (TEST INFRASTRUCTURE ONLY).

Nothing in the product (neurokmer_b200/) imports this module; only tests/,
``__graft_entry__.smoke()`` and bench.py's cpu_baseline leg may.

PARITY UNPINNED (see nk_oracle.c header and DESIGN.md §3).  This twin is
written from the *closed-form* definitions rather than from
the rolling recurrences that nk_oracle.c follows, so agreement between the two
checks the restatement of `RollingKmerHash` (reference src/models.rs:186-286)
against the direct windowed definition.

Also wraps libnk_oracle.so through ctypes (class ``COracle``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Iterable, List, Sequence, Tuple

import numpy as np

M64 = (1 << 64) - 1

# --- 2-bit codes: reference src/models.rs:231-251 ---------------------------
_F = {ord("A"): 0, ord("a"): 0, ord("C"): 1, ord("c"): 1, ord("G"): 2, ord("g"): 2, ord("T"): 3, ord("t"): 3}
_C = {ord("A"): 3, ord("a"): 3, ord("C"): 2, ord("c"): 2, ord("G"): 1, ord("g"): 1, ord("T"): 0, ord("t"): 0}


def fwd_word(window: bytes) -> int:
    """Σ F(s[j])·4^(k-1-j); non-ACGT → 0."""
    w = 0
    for b in window:
        w = (w << 2) | _F.get(b, 0)
    return w & M64


def rc_word(window: bytes) -> int:
    """Σ C(s[j])·4^j; non-ACGT → 0 (NOT 3) (reference src/models.rs:243-251)."""
    w = 0
    for j, b in enumerate(window):
        w |= _C.get(b, 0) << (2 * j)
    return w & M64


def pack_kmer(window: bytes) -> int:
    """Non-canonical pack: non-ACGT skipped, no mask (reference src/utils.rs:26-39)."""
    w = 0
    for b in window:
        if b in _F:
            w = ((w << 2) | _F[b]) & M64
    return w


def pack_nk2(bases) -> Tuple[np.ndarray, np.ndarray, int]:
    """numpy restatement of the product's pre-packed input layout (include/neurokmer.h, "nk2"):
    codes u32 (16 bases per word, first base in bits 31:30, F codes of src/models.rs:231-239),
    other u32 (bit p%32 of word p//32: byte p is not ACGTacgt), number of such bytes.
    Test infrastructure: the checker of nk_pack_bases, never called by the product."""
    a = np.frombuffer(bytes(bases), np.uint8) if not isinstance(bases, np.ndarray) else bases.astype(np.uint8)
    n = a.size
    lut_c = np.zeros(256, np.uint32)
    lut_o = np.ones(256, np.uint32)
    for b, c in _F.items():
        lut_c[b] = c
        lut_o[b] = 0
    c = np.zeros((n + 15) // 16 * 16, np.uint32)
    c[:n] = lut_c[a]
    codes = (c.reshape(-1, 16) << (30 - 2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    o = np.zeros((n + 31) // 32 * 32, np.uint32)
    o[:n] = lut_o[a]
    other = (o.reshape(-1, 32) << np.arange(32, dtype=np.uint32)).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return codes, other, int(lut_o[a].sum())


def kmer_words(seq: bytes, k: int, canonical: bool = True) -> List[int]:
    out = []
    for i in range(0, len(seq) - k + 1):
        win = seq[i : i + k]
        out.append(min(fwd_word(win), rc_word(win)) if canonical else pack_kmer(win))
    return out


# --- SipHash (published algorithm; siphasher 1.0.2 is not vendored) ----------
def _rotl(x: int, b: int) -> int:
    return ((x << b) | (x >> (64 - b))) & M64


def siphash(c: int, d: int, k0: int, k1: int, data: bytes) -> int:
    v0 = 0x736F6D6570736575 ^ k0
    v1 = 0x646F72616E646F6D ^ k1
    v2 = 0x6C7967656E657261 ^ k0
    v3 = 0x7465646279746573 ^ k1

    def rnd(v0, v1, v2, v3):
        v0 = (v0 + v1) & M64; v1 = _rotl(v1, 13); v1 ^= v0; v0 = _rotl(v0, 32)
        v2 = (v2 + v3) & M64; v3 = _rotl(v3, 16); v3 ^= v2
        v0 = (v0 + v3) & M64; v3 = _rotl(v3, 21); v3 ^= v0
        v2 = (v2 + v1) & M64; v1 = _rotl(v1, 17); v1 ^= v2; v2 = _rotl(v2, 32)
        return v0, v1, v2, v3

    n = len(data)
    full = n - n % 8
    for off in range(0, full, 8):
        m = int.from_bytes(data[off : off + 8], "little")
        v3 ^= m
        for _ in range(c):
            v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
        v0 ^= m
    b = ((n & 0xFF) << 56) | int.from_bytes(data[full:], "little")
    v3 ^= b
    for _ in range(c):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    v0 ^= b
    v2 ^= 0xFF
    for _ in range(d):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    return (v0 ^ v1 ^ v2 ^ v3) & M64


def siphash13_u64(x: int) -> int:
    """SipHasher13::new_with_keys(0,0) fed `x.hash()` = 8 LE bytes (reference src/spiking_hash.rs:78-82)."""
    return siphash(1, 3, 0, 0, int(x).to_bytes(8, "little"))


def neuron_index(word: int, pool_size: int) -> int:
    return siphash13_u64(word) % pool_size


def currents(seqs: Iterable[bytes], k: int, pool_size: int, canonical: bool = True) -> np.ndarray:
    cur = np.zeros(pool_size, dtype=np.uint64)
    for s in seqs:
        for w in kmer_words(s, k, canonical):
            cur[neuron_index(w, pool_size)] += np.uint64(1)
    return cur


# --- LIF: reference src/models.rs:34-51, src/spiking_hash.rs:187-200 / 544-659 ---
def lif_neuron(count: int, steps: int, thr: float, leak: float, period: int,
               v: float = 0.0, r: int = 0, skip_zero: bool = True) -> Tuple[int, float, int]:
    """One neuron, `steps` ticks, numpy float32 with separate multiply and add.

    skip_zero=True  → in-memory driver (zero-current neurons untouched, :189-191)
    skip_zero=False → streaming/SIMD driver (all neurons stepped, :561-647)
    Returns (new_spikes, v, r).
    """
    f32 = np.float32
    if steps == 0:
        return 0, float(f32(v)), r
    if skip_zero and count == 0:
        return 0, float(f32(v)), r
    inp = f32(np.float64(count) / np.float64(steps))
    vv, lk, th = f32(v), f32(leak), f32(thr)
    n = 0
    for _ in range(steps):
        if r > 0:
            r -= 1
            continue
        vv = f32(f32(vv * lk) + inp)
        if vv >= th:
            vv = f32(0.0)
            r = period
            n += 1
    return n, float(vv), r


def top_n(spikes: Sequence[int], n: int) -> List[Tuple[int, int]]:
    """Stable sort desc by spikes ⇒ ties by ascending idx (reference src/spiking_hash.rs:661-673)."""
    order = sorted(range(len(spikes)), key=lambda i: -int(spikes[i]))  # sorted() is stable
    return [(i, int(spikes[i])) for i in order[:n]]


def cost_fixed(spike_cost: float) -> int:
    """`(cost * 1000.0) as u64` (reference src/models.rs:162, src/spiking_hash.rs:649): Rust's float -> int
    `as` cast truncates toward zero and SATURATES (NaN -> 0, negative -> 0, >= 2^64 -> u64::MAX)."""
    x = float(spike_cost) * 1000.0
    if x != x or x <= 0.0:
        return 0
    if x >= 18446744073709551616.0:
        return M64
    return int(x)


def energy(total_spikes: int, spike_cost: float) -> float:
    """reference src/models.rs:159-172 + src/spiking_hash.rs:649-655 (wrapping u64 arithmetic)."""
    return float((total_spikes * cost_fixed(spike_cost)) & M64) / 1000.0


# --- ctypes wrapper around libnk_oracle.so -----------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnk_oracle.so")


def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "nk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _SO


class COracle:
    """ctypes view of oracle/libnk_oracle.so."""

    def __init__(self) -> None:
        self.lib = ctypes.CDLL(build_c_oracle())
        L = self.lib
        u8p, u64p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint64)
        f32p, u32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint32)
        U64, U32, F32, I = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float, ctypes.c_int
        L.nko_pack_kmer.restype = U64; L.nko_pack_kmer.argtypes = [u8p, U64]
        L.nko_siphash.restype = U64; L.nko_siphash.argtypes = [U32, U32, U64, U64, u8p, U64]
        L.nko_siphash13_u64.restype = U64; L.nko_siphash13_u64.argtypes = [U64]
        L.nko_neuron_index.restype = U64; L.nko_neuron_index.argtypes = [U64, U64]
        L.nko_kmer_words.restype = U64; L.nko_kmer_words.argtypes = [u8p, U64, U32, I, u64p]
        L.nko_kmer_fwd_rc.restype = U64; L.nko_kmer_fwd_rc.argtypes = [u8p, U64, U32, u64p, u64p]
        L.nko_accumulate.restype = U64; L.nko_accumulate.argtypes = [u8p, u64p, U64, U32, U64, I, u64p]
        L.nko_accumulate_mt.restype = U64; L.nko_accumulate_mt.argtypes = [u8p, u64p, U64, U32, U64, I, u64p, U32]
        lif_args = [u64p, U64, U64, U64, F32, F32, U32, f32p, u32p, u64p]
        L.nko_lif_scalar.restype = U64; L.nko_lif_scalar.argtypes = lif_args
        L.nko_lif_simd_semantics.restype = U64; L.nko_lif_simd_semantics.argtypes = lif_args
        L.nko_lif_mt.restype = U64; L.nko_lif_mt.argtypes = [u64p, U64, U64, F32, F32, U32, f32p, u32p, u64p, I, U32]
        L.nko_energy_fixed.restype = U64; L.nko_energy_fixed.argtypes = [U64, ctypes.c_double]
        L.nko_energy_total.restype = ctypes.c_double; L.nko_energy_total.argtypes = [U64]
        L.nko_top_n.restype = U64; L.nko_top_n.argtypes = [u64p, U64, U64, u64p, u64p]
        L.nko_process_sequence.restype = U64
        L.nko_process_sequence.argtypes = [u8p, U64, U32, U64, I, F32, F32, U32, u64p, f32p, u32p, u64p]

    # helpers ---------------------------------------------------------------
    @staticmethod
    def _p(a: np.ndarray, ct):
        return a.ctypes.data_as(ctypes.POINTER(ct))

    @staticmethod
    def _bases(seq) -> np.ndarray:
        if isinstance(seq, (bytes, bytearray)):
            return np.frombuffer(bytes(seq), dtype=np.uint8)
        return np.ascontiguousarray(seq, dtype=np.uint8)

    def siphash(self, c: int, d: int, k0: int, k1: int, data: bytes) -> int:
        a = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        return self.lib.nko_siphash(c, d, k0, k1, self._p(a, ctypes.c_uint8), len(data))

    def siphash13_u64(self, x: int) -> int:
        return self.lib.nko_siphash13_u64(x)

    def neuron_index(self, w: int, pool: int) -> int:
        return self.lib.nko_neuron_index(w, pool)

    def pack_kmer(self, win: bytes) -> int:
        a = self._bases(win) if len(win) else np.zeros(1, np.uint8)
        return self.lib.nko_pack_kmer(self._p(a, ctypes.c_uint8), len(win))

    def kmer_words(self, seq, k: int, canonical: bool = True) -> np.ndarray:
        a = self._bases(seq)
        n = max(0, a.size - k + 1)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        m = self.lib.nko_kmer_words(self._p(a if a.size else np.zeros(1, np.uint8), ctypes.c_uint8),
                                    a.size, k, int(canonical), self._p(out, ctypes.c_uint64))
        return out[:m]

    def kmer_fwd_rc(self, seq, k: int):
        a = self._bases(seq)
        n = max(0, a.size - k + 1)
        f = np.zeros(max(n, 1), dtype=np.uint64); r = np.zeros(max(n, 1), dtype=np.uint64)
        m = self.lib.nko_kmer_fwd_rc(self._p(a if a.size else np.zeros(1, np.uint8), ctypes.c_uint8),
                                     a.size, k, self._p(f, ctypes.c_uint64), self._p(r, ctypes.c_uint64))
        return f[:m], r[:m]

    def indices(self, words: np.ndarray, pool: int) -> np.ndarray:
        return np.array([self.lib.nko_neuron_index(int(w), pool) for w in words], dtype=np.uint64)

    def accumulate(self, bases: np.ndarray, offsets: np.ndarray, k: int, pool: int,
                   canonical: bool = True, currents: np.ndarray | None = None, threads: int = 1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if currents is None:
            currents = np.zeros(pool, dtype=np.uint64)
        b = bases if bases.size else np.zeros(1, np.uint8)
        nseq = offsets.size - 1
        if threads <= 1:
            tot = self.lib.nko_accumulate(self._p(b, ctypes.c_uint8), self._p(offsets, ctypes.c_uint64),
                                          nseq, k, pool, int(canonical), self._p(currents, ctypes.c_uint64))
        else:
            tot = self.lib.nko_accumulate_mt(self._p(b, ctypes.c_uint8), self._p(offsets, ctypes.c_uint64),
                                             nseq, k, pool, int(canonical),
                                             self._p(currents, ctypes.c_uint64), threads)
        return currents, tot

    def accumulate_exact(self, bases: np.ndarray, offsets: np.ndarray, k: int, pool: int, canonical: bool = True,
                         threads: int = 1, want_uniques: bool = True):
        """accumulate + the reference's exact side tables (per-task HashMap, merged counts, kmer_per_neuron):
        returns (currents, windows, n_distinct, uniques or None)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        currents = np.zeros(pool, dtype=np.uint64)
        uniques = np.zeros(pool, dtype=np.uint32) if want_uniques else None
        b = bases if bases.size else np.zeros(1, np.uint8)
        nd = ctypes.c_uint64()
        fn = self.lib.nko_accumulate_exact_mt
        fn.restype = ctypes.c_uint64
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint64, ctypes.c_int,
                       ctypes.c_void_p, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint64), ctypes.c_void_p]
        tot = fn(b.ctypes.data, offsets.ctypes.data, offsets.size - 1, k, pool, int(canonical), currents.ctypes.data,
                 max(1, threads), ctypes.byref(nd), None if uniques is None else uniques.ctypes.data)
        return currents, int(tot), int(nd.value), uniques

    def lif(self, currents: np.ndarray, steps: int, thr: float, leak: float, period: int,
            v: np.ndarray | None = None, r: np.ndarray | None = None, spikes: np.ndarray | None = None,
            simd_semantics: bool = False, threads: int = 1):
        pool = currents.size
        currents = np.ascontiguousarray(currents, dtype=np.uint64)
        v = np.zeros(pool, np.float32) if v is None else v
        r = np.zeros(pool, np.uint32) if r is None else r
        spikes = np.zeros(pool, np.uint64) if spikes is None else spikes
        fired = self.lib.nko_lif_mt(self._p(currents, ctypes.c_uint64), pool, steps, thr, leak, period,
                                    self._p(v, ctypes.c_float), self._p(r, ctypes.c_uint32),
                                    self._p(spikes, ctypes.c_uint64), int(simd_semantics), max(1, threads))
        return fired, v, r, spikes

    def top_n(self, spikes: np.ndarray, n: int):
        spikes = np.ascontiguousarray(spikes, dtype=np.uint64)
        m = min(n, spikes.size)
        oi = np.zeros(max(m, 1), np.uint64); os_ = np.zeros(max(m, 1), np.uint64)
        got = self.lib.nko_top_n(self._p(spikes, ctypes.c_uint64), spikes.size, n,
                                 self._p(oi, ctypes.c_uint64), self._p(os_, ctypes.c_uint64))
        return oi[:got], os_[:got]

    def energy(self, spikes: int, cost: float) -> float:
        return self.lib.nko_energy_total(self.lib.nko_energy_fixed(spikes, cost))

    def process_sequence(self, seq, k, pool, canonical, thr, leak, period, scratch, v, r, spikes) -> int:
        a = self._bases(seq)
        return self.lib.nko_process_sequence(self._p(a if a.size else np.zeros(1, np.uint8), ctypes.c_uint8),
                                             a.size, k, pool, int(canonical), thr, leak, period,
                                             self._p(scratch, ctypes.c_uint64), self._p(v, ctypes.c_float),
                                             self._p(r, ctypes.c_uint32), self._p(spikes, ctypes.c_uint64))


class OracleCounter:
    """CPU mirror of `SpikingKmerCounter`'s two batch entry points + read-outs
    (reference src/spiking_hash.rs:84-201, 277-486, 661-695) on top of COracle."""

    def __init__(self, k: int, threshold: float, leak: float, refractory: int, spike_cost: float,
                 pool_size: int, use_canonical: bool, steps: int = 1000, threads: int = 1) -> None:
        self.c = COracle()
        self.k, self.threshold, self.leak = k, np.float32(threshold), np.float32(leak)
        self.refractory, self.spike_cost, self.pool_size = refractory, spike_cost, pool_size
        self.use_canonical, self.steps, self.threads = use_canonical, steps, threads
        self.v = np.zeros(pool_size, np.float32)
        self.r = np.zeros(pool_size, np.uint32)
        self.spikes = np.zeros(pool_size, np.uint64)
        self.currents = np.zeros(pool_size, np.uint64)
        self.total_spikes = 0
        self.energy_fixed = 0

    def _simulate(self, simd: bool) -> None:
        fired, *_ = self.c.lif(self.currents, self.steps, float(self.threshold), float(self.leak),
                               self.refractory, self.v, self.r, self.spikes, simd, self.threads)
        self.total_spikes += fired
        self.energy_fixed = (self.energy_fixed + fired * cost_fixed(self.spike_cost)) & M64

    def process_parallel(self, bases: np.ndarray, offsets: np.ndarray) -> None:
        self.currents, _ = self.c.accumulate(bases, offsets, self.k, self.pool_size, self.use_canonical,
                                             None, self.threads)
        self._simulate(simd=False)

    def process_streaming(self, batches) -> None:
        cur = np.zeros(self.pool_size, np.uint64)
        for bases, offsets in batches:
            self.c.accumulate(bases, offsets, self.k, self.pool_size, self.use_canonical, cur, self.threads)
        self.currents = cur
        self._simulate(simd=True)

    def top_abundant_neurons(self, n: int):
        return self.c.top_n(self.spikes, n)

    def energy_used(self) -> float:
        return self.energy_fixed / 1000.0
