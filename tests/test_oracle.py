"""CPU tests (-m "not gpu"): pins the ORACLE (oracle/nk_oracle.c + oracle/oracle_py.py).

The reference ships no golden vectors and cannot be built here (PARITY UNPINNED, see
nk_oracle.c).  What pins the oracle instead:
  * SipHash-2-4's published test vectors (reference C implementation / paper appendix) for the
    round function + padding rule the 1-3 variant shares;
  * CPython's own SipHash-1-3 (PYTHONHASHSEED=0 => key 0,0 — the reference's keys) for the 1-3 variant;
  * the independent pure-Python twin (closed-form windowing vs the C oracle's rolling recurrences);
  * the survey's known-answer table and algebraic invariants;
  * tests/golden/golden_small.json.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import random_dna
from oracle import oracle_py as op

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.json")


def test_siphash24_published_vectors(coracle):
    key = bytes(range(16))
    k0, k1 = int.from_bytes(key[:8], "little"), int.from_bytes(key[8:], "little")
    vec = [0x726FDB47DD0E0E31, 0x74F839C593DC67FD, 0x0D6C8009D9A94F5A, 0x85676696D7FB7E2D, 0xCF2794E0277187B7,
           0x18765564CD99A68D, 0xCBC9466E58FEE3CE, 0xAB0200F58B01D137, 0x93F5F5799A932462]
    for n, want in enumerate(vec):
        msg = bytes(range(n))
        assert coracle.siphash(2, 4, k0, k1, msg) == want, n
        assert op.siphash(2, 4, k0, k1, msg) == want, n
    # the worked example of the SipHash paper (15-byte message)
    assert coracle.siphash(2, 4, k0, k1, bytes(range(15))) == 0xA129CA6149BE45E5


def test_siphash13_against_cpython():
    """CPython >= 3.11 hashes bytes with SipHash-1-3; PYTHONHASHSEED=0 zeroes the key."""
    if sys.hash_info.algorithm != "siphash13":
        pytest.skip("this CPython does not use siphash13")
    xs = [0, 1, 2, 0x1B, 0xDEADBEEF, 2**63, 2**64 - 1, 4**31 - 1] + [int(x) for x in
          np.random.default_rng(1).integers(0, 2**64, size=300, dtype=np.uint64)]
    code = "import struct,sys;print(' '.join(str(hash(struct.pack('<Q',int(x)))&(2**64-1)) for x in sys.argv[1:]))"
    out = subprocess.check_output([sys.executable, "-c", code] + [str(x) for x in xs],
                                  env=dict(os.environ, PYTHONHASHSEED="0")).split()
    from oracle.oracle_py import COracle
    c = COracle()
    checked = 0
    for x, h in zip(xs, out):
        ours = c.siphash13_u64(x)
        assert ours == op.siphash13_u64(x)
        if ours in (2**64 - 1, 2**64 - 2):  # CPython remaps hash -1 -> -2
            continue
        assert int(h) == ours, hex(x)
        checked += 1
    assert checked > 300


def test_survey_known_answers(coracle):
    assert coracle.siphash13_u64(0) == 0xBD60ACB658C79E45
    assert coracle.siphash13_u64(1) == 0x1E9F734161D62DD9
    assert coracle.siphash13_u64(0x1B) == 0xE38A965A565BD97F
    assert coracle.siphash13_u64(0xDEADBEEF) == 0x1E1D875FB6B69775
    s = b"ACGTNACGTAC"
    f, r = coracle.kmer_fwd_rc(s, 5)
    assert f.tolist() == [108, 432, 705, 774, 27, 108, 433]      # row 0: N reads as A forward ...
    assert r.tolist() == [27, 774, 705, 432, 108, 795, 710]      # ... but as 0 (not T) on the reverse strand
    w = coracle.kmer_words(s, 5, True)
    assert w.tolist() == [27, 432, 705, 432, 27, 108, 433]
    assert coracle.indices(w, 1_000_000).tolist() == [795071, 720115, 210720, 720115, 795071, 176367, 670755]
    assert coracle.indices(w, 2_000_000).tolist() == [1795071, 1720115, 1210720, 1720115, 1795071, 1176367, 670755]
    assert coracle.kmer_words(b"A" * 31, 31).tolist() == [0]
    assert coracle.kmer_fwd_rc(b"A" * 31, 31)[1].tolist() == [4**31 - 1]
    f, r = coracle.kmer_fwd_rc(b"T" * 32, 32)
    assert f.tolist() == [2**64 - 1] and r.tolist() == [0]
    w = coracle.kmer_words(b"ACGT" * 8, 31)
    assert w.tolist() == [488296166657017542] * 2 and coracle.neuron_index(int(w[0]), 2_000_000) == 1501330
    # pack_kmer skips non-ACGT and applies no mask (utils.rs:26-39)
    assert coracle.pack_kmer(b"ACGTN") == 27 and coracle.pack_kmer(b"NNNN") == 0 and coracle.pack_kmer(b"acgt") == 27
    assert coracle.kmer_words(b"ACNGT", 3, False).tolist() == [1, 6, 11]


@pytest.mark.parametrize("k", [1, 5, 15, 21, 31, 32])
def test_rolling_equals_closed_form(coracle, k):
    rng = np.random.default_rng(k)
    for n in (k, k + 1, 3 * k + 7, 400):
        s = random_dna(rng, n, 0.1, 0.2, 0.05)
        assert coracle.kmer_words(s, k, True).tolist() == op.kmer_words(s, k, True)
        assert coracle.kmer_words(s, k, False).tolist() == op.kmer_words(s, k, False)
        f, r = coracle.kmer_fwd_rc(s, k)
        assert f.tolist() == [op.fwd_word(s[i:i + k]) for i in range(n - k + 1)]
        assert r.tolist() == [op.rc_word(s[i:i + k]) for i in range(n - k + 1)]
    assert coracle.kmer_words(b"ACGT"[: k - 1], k).size == 0


def test_golden_fixture(coracle):
    g = json.load(open(GOLDEN))
    for case in g["cases"]:
        k, pool, canon = case["k"], case["pool"], case["canonical"]
        cur = np.zeros(pool, np.uint64)
        for si, s in enumerate(case["seqs"]):
            s = s.encode("latin1")
            w = coracle.kmer_words(s, k, canon)
            assert w.tolist() == case["words"][si]
            if canon:
                f, r = coracle.kmer_fwd_rc(s, k)
                assert f.tolist() == case["fwd"][si] and r.tolist() == case["rc"][si]
            assert coracle.indices(w, pool).tolist() == case["idx"][si]
            b = np.frombuffer(s, np.uint8)
            coracle.accumulate(b, np.array([0, b.size], np.uint64), k, pool, canon, cur)
        nz = np.nonzero(cur)[0]
        assert [[int(i), int(cur[i])] for i in nz] == [list(x) for x in case["currents"]]
    for x, h in g["siphash13"].items():
        assert coracle.siphash13_u64(int(x)) == h
    for blk in g["lif"]:
        counts = np.array([r[0] for r in blk["rows"]], np.uint64)
        for simd in (False, True):
            fired, v, r, sp = coracle.lif(counts, blk["steps"], blk["threshold"], blk["leak"], blk["refractory"],
                                          simd_semantics=simd)
            assert sp.tolist() == [row[1] for row in blk["rows"]]
            assert v.view(np.uint32).tolist() == [row[2] for row in blk["rows"]]
            assert r.tolist() == [row[3] for row in blk["rows"]]
    for c1, c2, n1, n2, vbits, r2 in g["lif_carried"]:
        cur = np.array([c1], np.uint64)
        _, v, r, sp = coracle.lif(cur, 1000, 1.0, 0.95, 2, simd_semantics=True)
        assert sp[0] == n1
        _, v, r, sp = coracle.lif(np.array([c2], np.uint64), 1000, 1.0, 0.95, 2, v, r, sp, simd_semantics=True)
        assert (int(sp[0]) - n1, int(v.view(np.uint32)[0]), int(r[0])) == (n2, vbits, r2)
    t = g["topn"]
    oi, os_ = coracle.top_n(np.array(t["spikes"], np.uint64), t["n"])
    assert [[int(a), int(b)] for a, b in zip(oi, os_)] == t["expect"]


def test_lif_transfer_table_and_fma_sensitivity(coracle):
    """SURVEY §A.3 table at the CLI's parameters; count=50 never reaches 1.0; the oracle must be
    built with -ffp-contract=off (separate multiply and add, as rustc emits)."""
    table = {0: 0, 50: 0, 51: 12, 52: 15, 53: 17, 54: 18, 55: 20, 56: 21, 57: 23, 58: 24, 59: 25, 75: 41, 100: 62,
             150: 100, 200: 125, 300: 167, 500: 200, 700: 250, 999: 250, 1000: 334, 10**6: 334}
    counts = np.array(list(table), np.uint64)
    fired, v, r, sp = coracle.lif(counts, 1000, 1.0, 0.95, 2)
    assert sp.tolist() == list(table.values()) and fired == sum(table.values())
    assert np.float32(v[1]) == np.float32(0.99999946) and r[1] == 0          # count = 50
    i300 = list(table).index(300)
    assert v[i300] == 0.0 and r[i300] == 2
    # separate rounding: v*leak+I evaluated with one rounding (FMA) differs in the last ulp for count=50
    f32 = np.float32
    vv, vf = f32(0), np.float64(0)
    I = f32(np.float64(50) / np.float64(1000))
    for _ in range(1000):
        vv = f32(f32(vv * f32(0.95)) + I)
    assert vv == v[1]
    # monotone non-decreasing on 0..1200
    cs = np.arange(0, 1201, dtype=np.uint64)
    _, _, _, sp = coracle.lif(cs, 1000, 1.0, 0.95, 2)
    assert (np.diff(sp.astype(np.int64)) >= 0).all() and sp[1000:].tolist() == [334] * 201


def test_lif_drivers_and_state_carry(coracle):
    """In-memory driver skips zero-current neurons (spiking_hash.rs:189-191); the SIMD driver steps
    every neuron (:561-647), so a resting neuron's leftover voltage decays; steps == 0 is a no-op."""
    cur = np.array([40, 0, 300], np.uint64)
    _, v, r, sp = coracle.lif(cur, 1000, 1.0, 0.95, 2)
    v0 = v.copy()
    zero = np.zeros(3, np.uint64)
    _, v1, r1, sp1 = coracle.lif(zero, 1000, 1.0, 0.95, 2, v.copy(), r.copy(), sp.copy(), simd_semantics=False)
    assert (v1 == v0).all() and (sp1 == sp).all()
    _, v2, r2, sp2 = coracle.lif(zero, 1000, 1.0, 0.95, 2, v.copy(), r.copy(), sp.copy(), simd_semantics=True)
    assert v2[0] < v0[0] and r2[2] == 0 and (sp2 == sp).all()
    f, v3, r3, sp3 = coracle.lif(cur, 0, 1.0, 0.95, 2, simd_semantics=True)
    assert f == 0 and sp3.sum() == 0
    # python twin == C for both drivers on a range of counts and carried state
    for cnt in (0, 1, 49, 50, 51, 333, 1000, 4000):
        for simd in (False, True):
            n, vv, rr = op.lif_neuron(cnt, 1000, 1.0, 0.95, 2, skip_zero=not simd)
            _, cv, cr, cs = coracle.lif(np.array([cnt], np.uint64), 1000, 1.0, 0.95, 2, simd_semantics=simd)
            assert (n, np.float32(vv), rr) == (int(cs[0]), cv[0], int(cr[0]))


def test_accumulate_invariants_and_threads(coracle):
    from neurokmer_b200.counter import flatten
    rng = np.random.default_rng(9)
    seqs = [random_dna(rng, int(n), 0.01, 0.05) for n in (0, 5, 30, 31, 32, 1000, 2_100_000, 77, 1_048_576 + 40)]
    bases, offsets = flatten(seqs)
    for k, pool in ((31, 2_000_000), (15, 65536)):
        c1, t1 = coracle.accumulate(bases, offsets, k, pool, True, threads=1)
        c8, t8 = coracle.accumulate(bases, offsets, k, pool, True, threads=5)
        assert t1 == t8 == sum(max(0, len(s) - k + 1) for s in seqs) == int(c1.sum())
        assert (c1 == c8).all()


def test_top_n_and_energy(coracle):
    sp = np.array([3, 7, 7, 0, 7, 1], np.uint64)
    oi, os_ = coracle.top_n(sp, 4)
    assert oi.tolist() == [1, 2, 4, 0] and os_.tolist() == [7, 7, 7, 3]
    oi, _ = coracle.top_n(sp, 100)
    assert oi.tolist() == [1, 2, 4, 0, 5, 3]
    assert op.top_n(sp.tolist(), 4) == [(1, 7), (2, 7), (4, 7), (0, 3)]
    assert coracle.energy(76082638, 1.0) == 76082638.0 and op.energy(10, 0.0015) == 0.01


def test_process_sequence_semantics(coracle):
    """process_sequence (spiking_hash.rs:203-273): raw count as the one-tick current, currents zeroed."""
    pool = 16
    v = np.zeros(pool, np.float32); r = np.zeros(pool, np.uint32); sp = np.zeros(pool, np.uint64)
    scratch = np.zeros(pool, np.uint64)
    fired = coracle.process_sequence(b"ACGTACGTACGT", 5, pool, True, 1.0, 0.95, 2, scratch, v, r, sp)
    assert fired == int(sp.sum()) > 0 and scratch.sum() == 0
    assert set(r[sp > 0].tolist()) == {2} and (v[sp > 0] == 0).all()
    assert coracle.process_sequence(b"ACG", 5, pool, True, 1.0, 0.95, 2, scratch, v, r, sp) == 0


def test_synth_twin_properties():
    from oracle.synth import synth_bases
    a = synth_bases(2, 0, 1 << 21, 3)
    b = synth_bases(2, 12345, 5000, 3)
    assert (a[12345:12345 + 5000] == b).all()               # position-addressable
    assert set(np.unique(a).tolist()) <= set(b"ACGTacgtN")
    n = (a == ord("N")).mean()
    assert 0.0005 < n < 0.02
    plain = synth_bases(1, 0, 100000, 0)
    assert set(np.unique(plain).tolist()) == set(b"ACGT")


def test_exact_map_variant_of_the_cpu_baseline(coracle):
    """nko_accumulate_exact_mt (hot path + the reference's per-task HashMap, merged `counts`, kmer_per_neuron —
    spiking_hash.rs:96,157-172): same currents as the plain accumulate, |counts| and the per-neuron distinct
    counts equal numpy's, for any thread count."""
    rng = np.random.default_rng(8)
    seqs = [rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=n).tobytes() for n in (200_000, 20, 0, 90_000, 31)]
    bases = np.frombuffer(b"".join(seqs), np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
    for k, pool, canon in ((21, 5000, True), (31, 100_003, True), (9, 4096, False)):
        cur, tot = coracle.accumulate(bases, offsets, k, pool, canon, threads=2)
        words = np.concatenate([coracle.kmer_words(s, k, canon) for s in seqs] + [np.zeros(0, np.uint64)])
        keys = np.unique(words)
        idx = np.array([coracle.neuron_index(int(w), pool) for w in keys], np.int64)
        for threads in (1, 3, 8):
            cur2, tot2, nd, uni = coracle.accumulate_exact(bases, offsets, k, pool, canon, threads=threads)
            assert tot2 == tot and nd == keys.size
            np.testing.assert_array_equal(cur2, cur)
            np.testing.assert_array_equal(uni, np.bincount(idx, minlength=pool).astype(np.uint32))


def test_shard_plan_partitions_the_windows():
    """nk_create_multi's shard plan (host logic, nk_debug_shard): the pieces of all members partition the windows
    of the batch — sequences cut by a shard boundary are read with a k-1 overlap and every window start has
    exactly one owner (the reference's fold over whole sequences, src/spiking_hash.rs:94-154, cut finer)."""
    from neurokmer_b200.counter import debug_shard, flatten
    from oracle.oracle_py import COracle
    c = COracle()
    rng = np.random.default_rng(12)
    for k, world in [(31, 2), (21, 3), (5, 8), (32, 4), (1, 5)]:
        lens = [0, 1, k - 1, k, 9000, 150, 20, 0, 30000, 4096, 4097, 12288 - k + 1, 7]
        seqs = [rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n).tobytes() for n in lens]
        bases, offsets = flatten(seqs)
        pool = 10007
        want, tot = c.accumulate(bases, offsets, k, pool, True)
        got = np.zeros(pool, np.uint64)
        windows = 0
        for r in range(world):
            start, po = debug_shard(offsets, k, world, r)
            assert start % 4096 == 0 or start == bases.size
            assert np.all(np.diff(po.astype(np.int64)) > 0) or po.size == 1
            piece = bases[start:start + int(po[-1])]
            cur, t = c.accumulate(piece, po, k, pool, True)
            got += cur
            windows += t
        assert windows == tot
        np.testing.assert_array_equal(got, want)
