"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU
oracle on the same seeded inputs.  Bar: bit-exact for words, indices, currents, spike
counts, refractory ticks, voltages (f32 bit patterns), totals and top-N.

Each test cites the reference lines whose behaviour it pins."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import random_dna

pytestmark = pytest.mark.gpu

REF = dict(threshold=1.0, leak=0.95, refractory=2, spike_cost=1.0)  # reference src/main.rs:37


def make(k, pool, canonical=True, **kw):
    from neurokmer_b200 import SpikingKmerCounter
    p = dict(REF); p.update(kw)
    return SpikingKmerCounter(k, p["threshold"], p["leak"], p["refractory"], p["spike_cost"], pool, canonical)


def oracle_counter(k, pool, canonical=True, steps=1000, **kw):
    from oracle.oracle_py import OracleCounter
    p = dict(REF); p.update(kw)
    return OracleCounter(k, p["threshold"], p["leak"], p["refractory"], p["spike_cost"], pool, canonical, steps, threads=4)


def assert_state_equal(c, o):
    np.testing.assert_array_equal(c.currents(), o.currents)
    np.testing.assert_array_equal(c.spike_counts(), o.spikes)
    np.testing.assert_array_equal(c.refractory_ticks(), o.r)
    np.testing.assert_array_equal(c.voltages().view(np.uint32), o.v.view(np.uint32))
    assert c.energy.total_spikes() == o.total_spikes
    assert c.energy_used() == o.energy_used()


def assert_topn_equal(c, o, n):
    got = c.top_abundant_neurons(n)
    oi, os_ = o.top_abundant_neurons(n)
    assert [g[0] for g in got] == [int(x) for x in oi]
    assert [g[1] for g in got] == [int(x) for x in os_]


# --- step 1+2: windowing, canonical min, SipHash, modulo ------------------------------------
@pytest.mark.parametrize("k", [1, 2, 5, 15, 16, 17, 21, 31, 32])
@pytest.mark.parametrize("canonical", [True, False])
def test_kmer_words_and_indices(coracle, k, canonical):
    """RollingKmerHash init/slide/canonical (models.rs:206-286), pack_kmer (utils.rs:26-39),
    map_kmer_to_neuron (spiking_hash.rs:78-82)."""
    rng = np.random.default_rng(100 + k)
    pool = 1_000_003 if k % 2 else 65536
    c = make(k, pool, canonical)
    for n, pn, pl, pi in [(k, 0, 0, 0), (k + 1, 0.2, 0, 0), (700, 0.05, 0.2, 0.02), (20000, 0.01, 0.01, 0.01),
                          (16384 + 40, 0.0, 0.0, 0.0), (33000, 0.3, 0.3, 0.1)]:
        s = random_dna(rng, n, pn, pl, pi)
        fwd, rc, words, idx = c.debug_kmers(s)
        ow = coracle.kmer_words(s, k, canonical)
        assert words.size == ow.size == n - k + 1
        np.testing.assert_array_equal(words, ow)
        if canonical:
            of, orc = coracle.kmer_fwd_rc(s, k)
            np.testing.assert_array_equal(fwd, of)
            np.testing.assert_array_equal(rc, orc)
        step = max(1, ow.size // 3000)
        np.testing.assert_array_equal(idx[::step], coracle.indices(ow[::step], pool))
    assert c.debug_kmers(b"ACGT"[: k - 1] if k > 1 else b"")[2].size == 0  # shorter than k: no windows


def test_hash_kat_and_random(coracle):
    """SipHasher13::new_with_keys(0,0) on 8 LE bytes (siphasher 1.0.2; SURVEY §A.4 vectors)."""
    c = make(31, 2_000_000)
    kat = {0: 0xBD60ACB658C79E45, 1: 0x1E9F734161D62DD9, 0x1B: 0xE38A965A565BD97F, 0xDEADBEEF: 0x1E1D875FB6B69775}
    words = np.array(list(kat.keys()), np.uint64)
    hs, ix = c.debug_hash(words)
    assert [int(h) for h in hs] == list(kat.values())
    assert [int(i) for i in ix] == [v % 2_000_000 for v in kat.values()]
    rng = np.random.default_rng(7)
    w = rng.integers(0, 2**64, size=5000, dtype=np.uint64)
    w[:8] = [0, 1, 2**64 - 1, 2**63, 2**62 - 1, 0xFFFFFFFF, 0x100000000, 4**31 - 1]
    for pool in [1, 2, 3, 1000, 65536, 1_000_000, 2_000_000, 16_000_000, 2**31, 2**31 + 1, 2**32 - 1]:
        cc = make(31, pool) if pool <= 16_000_000 else None
        if cc is None:
            continue
        hs, ix = cc.debug_hash(w)
        exp_h = np.array([coracle.siphash13_u64(int(x)) for x in w[:600]], np.uint64)
        np.testing.assert_array_equal(hs[:600], exp_h)
        np.testing.assert_array_equal(ix[:600], exp_h % np.uint64(pool))
        cc.close()


# --- step 3: currents ------------------------------------------------------------------------
def ragged_batch(rng, k):
    lens = [0, 1, k - 1, k, k + 1, 3 * k, 511, 512, 513, 16383, 16384, 16385, 40000, 0, 7, 150, 150, 20, 150]
    seqs = [random_dna(rng, max(0, n), 0.01, 0.05, 0.005) for n in lens]
    return seqs


@pytest.mark.parametrize("k,pool,canonical", [(21, 1_000_000, True), (31, 2_000_000, True), (15, 65536, True),
                                              (32, 999_983, True), (31, 1_000_000, False), (7, 1000, False), (1, 3, True)])
def test_currents_ragged(coracle, k, pool, canonical):
    """currents[idx] += 1 per window, sequences shorter than k contribute nothing
    (spiking_hash.rs:102-139); Σ currents = Σ max(0, L-k+1)."""
    from neurokmer_b200 import flatten
    rng = np.random.default_rng(k * 1000 + pool % 997)
    seqs = ragged_batch(rng, k)
    bases, offsets = flatten(seqs)
    c = make(k, pool, canonical)
    c.process_batch(bases, offsets)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, canonical, threads=4)
    got = c.currents()
    np.testing.assert_array_equal(got, exp)
    assert int(got.sum()) == tot == sum(max(0, len(s) - k + 1) for s in seqs)
    assert c.timings()["kmers"] == tot


def test_empty_and_degenerate_inputs():
    c = make(31, 1000)
    c.process_parallel([])
    assert c.currents().sum() == 0 and c.energy.total_spikes() == 0
    c.process_parallel([b"", b"ACGT", b"N" * 30])
    assert c.currents().sum() == 0
    c.process_parallel([b"A" * 31])  # fwd 0, rc 4^31-1, canon 0  
    cur = c.currents()
    assert cur.sum() == 1
    from oracle.oracle_py import neuron_index
    assert cur[neuron_index(0, 1000)] == 1


# --- full pipeline: LIF + totals + top-N -------------------------------------------------------
@pytest.mark.parametrize("k,pool,n", [(21, 100_000, 3_000_000), (15, 4096, 2_000_000), (31, 50_000, 2_500_000)])
def test_process_parallel_full(coracle, k, pool, n):
    """process_parallel (spiking_hash.rs:84-201): currents overwrite, in-memory LIF driver
    (zero-current neurons skipped), EnergyTracker, top_abundant_neurons (:661-673)."""
    rng = np.random.default_rng(n + k)
    seqs = [random_dna(rng, n // 2, 0.001, 0.01), random_dna(rng, n // 3, 0.0, 0.0), random_dna(rng, n // 6, 0.01, 0.0)]
    from neurokmer_b200 import flatten
    bases, offsets = flatten(seqs)
    c = make(k, pool)
    o = oracle_counter(k, pool)
    c.process_batch(bases, offsets)
    o.process_parallel(bases, offsets)
    assert c.timings()["lif_path"] in (2, 3)  # fresh state: per-count table (3 = fused with top-N)
    assert_state_equal(c, o)
    assert o.total_spikes > 0
    for topn in (1, 20, 500, 3000):
        assert_topn_equal(c, o, topn)
    # second call on the same counter: neuron state carries over, currents are overwritten,
    # and the direct simulation path runs (state is no longer uniform)
    seqs2 = [random_dna(rng, n // 4, 0.0, 0.0)]
    b2, o2 = flatten(seqs2)
    c.process_batch(b2, o2)
    o.process_parallel(b2, o2)
    assert c.timings()["lif_path"] == 1
    assert_state_equal(c, o)
    assert_topn_equal(c, o, 20)


def test_lif_table_equals_direct(coracle):
    """The per-count table (fresh state) and the direct simulation give identical
    (spike_count, voltage, refractory) for every neuron — incl. the count=50 edge that ends
    at v=0.99999946 without ever spiking."""
    rng = np.random.default_rng(5)
    pool = 20000
    seqs = [random_dna(rng, 1_500_000)]
    a, b = make(21, pool), make(21, pool)
    b.debug_set_lif_path(1)
    a.process_parallel(seqs); b.process_parallel(seqs)
    assert a.timings()["lif_path"] in (2, 3) and b.timings()["lif_path"] == 1
    np.testing.assert_array_equal(a.currents(), b.currents())
    np.testing.assert_array_equal(a.spike_counts(), b.spike_counts())
    np.testing.assert_array_equal(a.refractory_ticks(), b.refractory_ticks())
    np.testing.assert_array_equal(a.voltages().view(np.uint32), b.voltages().view(np.uint32))
    assert a.energy.total_spikes() == b.energy.total_spikes() > 0
    cur = a.currents()
    assert cur.min() < 50 < cur.max()
    # transfer function at the reference's parameters
    table = {0: 0, 50: 0, 51: 12, 52: 15, 55: 20, 59: 25, 75: 41, 100: 62, 150: 100, 200: 125, 300: 167, 500: 200}
    sp = a.spike_counts()
    for cnt, spikes in table.items():
        hit = np.nonzero(cur == cnt)[0]
        if hit.size:
            assert set(sp[hit].tolist()) == {spikes}, (cnt, spikes)


@pytest.mark.parametrize("params", [dict(threshold=0.5, leak=0.9, refractory=0), dict(threshold=2.5, leak=1.0, refractory=5),
                                    dict(threshold=1.0, leak=0.0, refractory=1), dict(threshold=0.05, leak=0.5, refractory=3)])
def test_lif_other_parameters(coracle, params):
    """LifNeuron::update (models.rs:34-51) for parameters other than the CLI's, steps != 1000."""
    rng = np.random.default_rng(11)
    seqs = [random_dna(rng, 400_000)]
    from neurokmer_b200 import flatten
    bases, offsets = flatten(seqs)
    for steps in (1, 37, 1000):
        c = make(15, 5000, **params); c.set_steps(steps)
        o = oracle_counter(15, 5000, steps=steps, **params)
        assert c.get_steps() == steps
        c.process_batch(bases, offsets); o.process_parallel(bases, offsets)
        assert_state_equal(c, o)
        c.process_batch(bases, offsets); o.process_parallel(bases, offsets)  # direct path, carried state
        assert_state_equal(c, o)
        assert_topn_equal(c, o, 20)
    c = make(15, 5000); c.set_steps(0)  # steps == 0: no simulation (spiking_hash.rs:548-551)
    c.process_batch(bases, offsets)
    assert c.energy.total_spikes() == 0 and c.currents().sum() == 400_000 - 14


def test_streaming_equals_in_memory(coracle):
    """process_file_streaming minus parsing (spiking_hash.rs:277-486): batches accumulate,
    totals overwrite currents, SIMD-semantics LIF (every neuron stepped); README's
    'in-memory == streaming' claim."""
    rng = np.random.default_rng(21)
    from neurokmer_b200 import flatten
    batches = [flatten([random_dna(rng, 600_000, 0.001), random_dna(rng, 100, 0.0)]),
               flatten([random_dna(rng, 40, 0.0), random_dna(rng, 900_000, 0.0, 0.05)]),
               flatten([]), flatten([random_dna(rng, 10)])]
    k, pool = 31, 30000
    c = make(k, pool); o = oracle_counter(k, pool)
    c.stream_begin()
    for b, off in batches:
        c.stream_push(b, off)
    c.stream_end()
    o.process_streaming(batches)
    assert_state_equal(c, o)
    assert_topn_equal(c, o, 20)
    # in-memory on the concatenation gives the same totals
    allb, alloff = flatten([bytes(b[int(off[i]):int(off[i + 1])]) for b, off in batches for i in range(off.size - 1)])
    m = make(k, pool); m.process_batch(allb, alloff)
    np.testing.assert_array_equal(m.currents(), c.currents())
    assert m.energy.total_spikes() == c.energy.total_spikes()
    # a second streaming run on the same counter (carried state, zero-current neurons decay)
    c.stream_begin(); c.stream_push(*batches[1]); c.stream_end()
    o.process_streaming([batches[1]])
    assert_state_equal(c, o)


def test_chunked_h2d_pipeline_boundaries(coracle):
    """Host batches larger than the 32 MiB copy granule are cut into chunks with a k-1 halo;
    sequence ends that straddle a chunk boundary must neither lose nor double-count windows."""
    rng = np.random.default_rng(33)
    G = 32 << 20
    lens = [G - 17, 40, 3, G + 5, 31, 30, 1000]  # ends at G-17, G+23 (straddles), ...
    seqs = [random_dna(rng, n, 0.0005) for n in lens]
    from neurokmer_b200 import flatten
    bases, offsets = flatten(seqs)
    k, pool = 31, 2_000_000
    c = make(k, pool)
    c.process_batch(bases, offsets)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, True, threads=8)
    np.testing.assert_array_equal(c.currents(), exp)
    assert c.timings()["kmers"] == tot


def test_process_sequence(coracle):
    """process_sequence (spiking_hash.rs:203-273): one tick with the raw count, currents zeroed."""
    rng = np.random.default_rng(44)
    k, pool = 5, 64
    c = make(k, pool, True)
    v = np.zeros(pool, np.float32); r = np.zeros(pool, np.uint32); sp = np.zeros(pool, np.uint64)
    scratch = np.zeros(pool, np.uint64)
    total = 0
    for n in (3, 5, 40, 200, 7, 64, 300, 12):
        s = random_dna(rng, n, 0.05)
        c.process_sequence(s)
        total += coracle.process_sequence(s, k, pool, True, 1.0, 0.95, 2, scratch, v, r, sp)
        np.testing.assert_array_equal(c.spike_counts(), sp)
        np.testing.assert_array_equal(c.refractory_ticks(), r)
        np.testing.assert_array_equal(c.voltages().view(np.uint32), v.view(np.uint32))
        assert c.currents().sum() == 0
        assert c.energy.total_spikes() == total
    assert total > 0


def test_top_n_ties_and_limits():
    """Stable sort descending ⇒ ties by ascending index; min(top_n, pool) rows; zero-spike
    neurons are returned if they rank (spiking_hash.rs:661-673)."""
    rng = np.random.default_rng(55)
    c = make(15, 4096)
    c.process_parallel([random_dna(rng, 6_000_000)])  # ~1465 hits per neuron: all saturate at 334
    sp = c.spike_counts()
    assert set(sp.tolist()) == {334}
    top = c.top_abundant_neurons(20)
    assert [t[0] for t in top] == list(range(20)) and all(t[1] == 334 for t in top)
    assert all(t[2] is None for t in top)  # uniques not computed: never faked
    assert len(c.top_abundant_neurons(10_000)) == 4096
    full = c.top_abundant_neurons(4096)
    assert [t[0] for t in full] == list(range(4096))
    assert c.energy.total_spikes() == 334 * 4096
    z = make(31, 100)
    assert z.top_abundant_neurons(5) == [(i, 0, None) for i in range(5)]
    assert z.top_abundant_neurons(0) == []


def test_top_n_large_n(coracle):
    rng = np.random.default_rng(66)
    pool = 50_000
    c = make(21, pool); o = oracle_counter(21, pool)
    from neurokmer_b200 import flatten
    bases, offsets = flatten([random_dna(rng, 3_000_000)])
    c.process_batch(bases, offsets); o.process_parallel(bases, offsets)
    for n in (2048, 2049, 5000, 50_000):
        assert_topn_equal(c, o, n)


# --- files ----------------------------------------------------------------------------------------
def test_process_file_fasta_fastq(tmp_path, coracle):
    """stream_sequences record rules (utils.rs:9-24) + both CLI modes (main.rs:39-46)."""
    from neurokmer_b200 import NkError, flatten
    from neurokmer_b200.fastx import write_fasta, write_fastq
    rng = np.random.default_rng(77)
    seqs = [random_dna(rng, n, 0.002, 0.02) for n in (250_000, 61, 60, 59, 0, 31, 30, 100_000)]
    fa, fq = str(tmp_path / "a.fasta"), str(tmp_path / "a.fastq")
    write_fasta(fa, seqs); write_fastq(fq, seqs)
    bases, offsets = flatten(seqs)
    k, pool = 31, 10_000
    for path in (fa, fq):
        for streaming in (True, False):
            c = make(k, pool); o = oracle_counter(k, pool)
            (c.process_file_streaming if streaming else c.process_file_in_memory)(path)
            (o.process_streaming([(bases, offsets)]) if streaming else o.process_parallel(bases, offsets))
            assert_state_equal(c, o)
    # CRLF line ends are stripped
    crlf = str(tmp_path / "crlf.fasta")
    with open(fa, "rb") as f, open(crlf, "wb") as g:
        g.write(f.read().replace(b"\n", b"\r\n"))
    c = make(k, pool); c.process_file_streaming(crlf)
    exp, _ = coracle.accumulate(bases, offsets, k, pool, True)
    np.testing.assert_array_equal(c.currents(), exp)
    # gzip input is decompressed transparently (needletail sniffs the magic bytes)
    import gzip
    for src in (fa, fq):
        gzp = src + ".gz"
        with open(src, "rb") as f, gzip.open(gzp, "wb", compresslevel=1) as g:
            g.write(f.read())
        c = make(k, pool); c.process_file_streaming(gzp)
        np.testing.assert_array_equal(c.currents(), exp)
    # ... and so are bzip2, xz and zstd (needletail's "compression" feature)
    import bz2
    import lzma
    import pyarrow as pa
    raw = open(fa, "rb").read()
    for ext, blob in ((".bz2", bz2.compress(raw, 1)), (".xz", lzma.compress(raw, preset=0)),
                      (".zst", pa.Codec("zstd").compress(raw, asbytes=True))):
        with open(fa + ext, "wb") as g:
            g.write(blob)
        c = make(k, pool); c.process_file_streaming(fa + ext)
        np.testing.assert_array_equal(c.currents(), exp)
    # a malformed FASTQ record ends the stream: records before it still count (utils.rs:17-20)
    bad = str(tmp_path / "bad.fastq")
    with open(bad, "wb") as g:
        g.write(b"@r0\n" + seqs[0] + b"\n+\n" + b"I" * len(seqs[0]) + b"\n")
        g.write(b"@r1\n" + seqs[7] + b"\n+\n" + b"I" * 5 + b"\n")          # quality length mismatch
        g.write(b"@r2\n" + seqs[1] + b"\n+\n" + b"I" * len(seqs[1]) + b"\n")
    c = make(k, pool); c.process_file_streaming(bad)
    b0, o0 = flatten([seqs[0]])
    exp, _ = coracle.accumulate(b0, o0, k, pool, True)
    np.testing.assert_array_equal(c.currents(), exp)
    # open failure propagates (utils.rs:10 `?`)
    with pytest.raises(NkError) as ei:
        make(k, pool).process_file_streaming(str(tmp_path / "missing.fa"))
    assert ei.value.code == 2
    empty = str(tmp_path / "empty.fa"); open(empty, "wb").close()
    with pytest.raises(NkError):
        make(k, pool).process_file_streaming(empty)


def test_file_sequence_longer_than_batch(tmp_path, coracle):
    """A FASTA record longer than the 32 MiB pinned batch is cut into pieces overlapping by k-1."""
    from neurokmer_b200 import flatten
    from neurokmer_b200.fastx import write_fasta
    rng = np.random.default_rng(88)
    seqs = [random_dna(rng, (32 << 20) + 12345, 0.0003), random_dna(rng, 5000)]
    fa = str(tmp_path / "long.fasta"); write_fasta(fa, seqs)
    bases, offsets = flatten(seqs)
    k, pool = 31, 2_000_000
    c = make(k, pool); c.process_file_streaming(fa)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, True, threads=8)
    np.testing.assert_array_equal(c.currents(), exp)


# --- error behaviour ----------------------------------------------------------------------------
def test_bad_arguments():
    from neurokmer_b200 import NkError
    for k in (0, 33):
        with pytest.raises(NkError) as ei:
            make(k, 1000)
        assert ei.value.code == 1
    with pytest.raises(NkError):
        make(31, 0)
    c = make(31, 1000)
    with pytest.raises(NkError) as ei:
        c.stream_push(np.zeros(4, np.uint8), np.array([0, 4], np.uint64))
    assert ei.value.code == 6
    with pytest.raises(NkError):
        c.process_batch(np.zeros(4, np.uint8), np.array([1, 4], np.uint64))
    with pytest.raises(NkError):
        c.process_batch(np.zeros(4, np.uint8), np.array([0, 4, 2], np.uint64))
    with pytest.raises(NkError) as ei:
        c.get_count(5)
    assert ei.value.code == 7


# --- device-resident path + synthetic generator (what bench.py times) ---------------------------
def test_staged_synthetic_matches_host_path(coracle):
    k, pool, n = 31, 200_000, 5_000_000
    c = make(k, pool)
    lens = [2_000_000, 1_500_000, 1_000_000, 500_000]
    db, do = c.stage_reserve(n, len(lens))
    c.synth_fill(db, 2, 0, n, 3)
    c.synchronize()
    # copy offsets into the library's buffer and read the generated bases back
    from neurokmer_b200.devmem import copy_h2d, device_to_numpy
    copy_h2d(do, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64))
    bases = device_to_numpy(db, n)
    assert set(np.unique(bases).tolist()) <= set(b"ACGTacgtN")
    assert (bases == ord("N")).sum() > 0 and np.isin(bases, np.frombuffer(b"acgt", np.uint8)).sum() > 0
    c.process_staged(n, len(lens), 0)
    o = oracle_counter(k, pool)
    o.process_parallel(bases, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64))
    assert_state_equal(c, o)
    t = c.timings()
    assert t["kmers"] == sum(l - k + 1 for l in lens) and t["count_ms"] > 0 and t["launches"] >= 2


# --- committed golden fixtures + generator twin ---------------------------------------------------
def test_golden_fixture_gpu():
    """tests/golden/golden_small.json (made by tests/golden/make_golden.py from the pure-Python
    restatement) against the CUDA path: words, fwd/rc, indices, currents, LIF end states, top-N."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.json")))
    for case in g["cases"]:
        k, pool, canon = case["k"], case["pool"], case["canonical"]
        c = make(k, pool, canon)
        seqs = [s.encode("latin1") for s in case["seqs"]]
        for si, s in enumerate(seqs):
            fwd, rc, words, idx = c.debug_kmers(s)
            assert words.tolist() == case["words"][si] and idx.tolist() == case["idx"][si]
            if canon:
                assert fwd.tolist() == case["fwd"][si] and rc.tolist() == case["rc"][si]
        c.process_parallel(seqs)
        cur = c.currents()
        nz = np.nonzero(cur)[0]
        assert [[int(i), int(cur[i])] for i in nz] == [list(x) for x in case["currents"]]
        c.close()
    c = make(31, 2_000_000)
    hs, _ = c.debug_hash(np.array([int(x) for x in g["siphash13"]], np.uint64))
    assert hs.tolist() == list(g["siphash13"].values())


def test_lif_golden_rows_gpu():
    """LIF end states (spikes, voltage bits, refractory) per count from the golden file, through
    both GPU paths (per-count table and direct simulation), and the carried-state rows."""
    import json
    from neurokmer_b200.devmem import copy_h2d
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.json")))
    for blk in g["lif"]:
        counts = np.array([r[0] for r in blk["rows"]], np.uint64)
        for force_direct in (0, 1, 2):
            c = make(31, counts.size, threshold=blk["threshold"], leak=blk["leak"], refractory=blk["refractory"])
            c.set_steps(blk["steps"])
            c.debug_set_lif_path(force_direct)
            c.stream_begin()
            ptr = c.stream_accumulated()       # device currents: inject the golden counts
            c.synchronize()                    # the hand-off is stream-ordered, cudaMemcpy is not
            copy_h2d(ptr, counts)
            c.stream_finish()
            memo_able = blk["leak"] >= 0
            assert c.timings()["lif_path"] in {0: (2, 3), 1: (1,), 2: ((5,) if memo_able else (1,))}[force_direct]
            assert c.spike_counts().tolist() == [r[1] for r in blk["rows"]]
            assert c.voltages().view(np.uint32).tolist() == [r[2] for r in blk["rows"]]
            assert c.refractory_ticks().tolist() == [r[3] for r in blk["rows"]]
            c.close()
    rows = g["lif_carried"]
    c = make(31, len(rows))
    for col in (0, 1):
        c.stream_begin()
        ptr = c.stream_accumulated()
        c.synchronize()
        copy_h2d(ptr, np.array([r[col] for r in rows], np.uint64))
        c.stream_finish()
        if col == 0:
            assert c.spike_counts().tolist() == [r[2] for r in rows]
    assert c.spike_counts().tolist() == [r[2] + r[3] for r in rows]
    assert c.voltages().view(np.uint32).tolist() == [r[4] for r in rows]
    assert c.refractory_ticks().tolist() == [r[5] for r in rows]
    t = g["topn"]
    c = make(31, len(t["spikes"]))


def test_synth_generator_matches_numpy_twin():
    from neurokmer_b200.devmem import device_to_numpy
    from oracle.synth import synth_bases
    c = make(31, 1000)
    n = 3_000_000
    db, _ = c.stage_reserve(n, 1)
    for seed, start, flags in [(2, 0, 3), (5, 10**10 + 7, 3), (1, 123, 0), (4, 1 << 33, 1)]:
        c.synth_fill(db, seed, start, n, flags)
        c.synchronize()
        np.testing.assert_array_equal(device_to_numpy(db, n), synth_bases(seed, start, n, flags))


# --- full-size property checks (BASELINE configs at their own sizes, no oracle pass needed) ------
def test_config1_full_size_properties(coracle):
    """configs[0]: 10 Mbp random ACGT, k=21, pool 1M, canonical, in-memory: checksum of currents
    = number of windows; result identical through the host path, the staged path and 3 pushes;
    the oracle (fast enough at this size) agrees bit for bit."""
    from neurokmer_b200.devmem import copy_h2d
    from oracle.synth import synth_bases
    n, k, pool = 10_000_000, 21, 1_000_000
    bases = synth_bases(1, 0, n, 0)
    offsets = np.array([0, n], np.uint64)
    a = make(k, pool); a.process_batch(bases, offsets)
    cur = a.currents()
    assert int(cur.sum()) == n - k + 1 == a.timings()["kmers"]
    b = make(k, pool)
    db, do = b.stage_reserve(n, 1)
    b.synth_fill(db, 1, 0, n, 0); copy_h2d(do, offsets); b.synchronize()
    b.process_staged(n, 1, 0)
    np.testing.assert_array_equal(b.currents(), cur)
    np.testing.assert_array_equal(b.spike_counts(), a.spike_counts())
    # streaming in 3 pieces that overlap by k-1 bases == one batch
    s = make(k, pool); s.stream_begin()
    cuts = [0, 3_333_333, 7_000_001, n]
    for i in range(3):
        lo, hi = cuts[i], min(n, cuts[i + 1] + k - 1)
        s.stream_push(bases[lo:hi], np.array([0, hi - lo], np.uint64))
    s.stream_end()
    np.testing.assert_array_equal(s.currents(), cur)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, True, threads=8)
    np.testing.assert_array_equal(cur, exp)
    o = oracle_counter(k, pool); o.process_parallel(bases, offsets)
    assert_state_equal(a, o); assert_topn_equal(a, o, 20)


def test_config4_contention_properties():
    """configs[3] shape (k=15, pool 65,536 = power of two, heavy collisions) at 200 Mbp on device:
    Σ currents = windows; every neuron saturates at ceil(1000/3) = 334 spikes; the tie rule makes
    the top-20 neurons 0..19."""
    from neurokmer_b200.devmem import copy_h2d
    n, k, pool = 200_000_000, 15, 65536
    c = make(k, pool)
    db, do = c.stage_reserve(n, 1)
    c.synth_fill(db, 4, 0, n, 0); copy_h2d(do, np.array([0, n], np.uint64)); c.synchronize()
    c.process_staged(n, 1, 0)
    cur = c.currents()
    assert int(cur.sum()) == n - k + 1 and cur.min() >= 1000
    assert set(c.spike_counts().tolist()) == {334}
    assert c.energy.total_spikes() == 334 * pool
    assert [t[0] for t in c.top_abundant_neurons(20)] == list(range(20))


def test_config3_reads_properties(coracle):
    """configs[2] shape: 150 bp reads (here 2 M reads = 300 Mbp), 1 % with an N, 0.1 % shorter than k:
    Σ currents = Σ max(0, L-k+1); a 100k-read sample agrees with the oracle bit for bit."""
    rng = np.random.default_rng(3)
    nreads, k, pool = 2_000_000, 31, 2_000_000
    lens = np.full(nreads, 150, np.int64)
    lens[rng.random(nreads) < 0.001] = 20
    offsets = np.zeros(nreads + 1, np.uint64); offsets[1:] = np.cumsum(lens)
    from oracle.synth import synth_bases
    bases = synth_bases(3, 0, int(offsets[-1]), 0)
    hit = rng.random(nreads) < 0.01
    bases[(offsets[:-1][hit] + rng.integers(0, 20, size=int(hit.sum())).astype(np.uint64)).astype(np.int64)] = ord("N")
    c = make(k, pool); c.process_batch(bases, offsets)
    want = int(np.maximum(lens - k + 1, 0).sum())
    assert int(c.currents().sum()) == want == c.timings()["kmers"]
    m = 100_000
    s = make(k, pool); s.process_batch(bases[: int(offsets[m])], offsets[: m + 1])
    exp, _ = coracle.accumulate(bases[: int(offsets[m])], offsets[: m + 1], k, pool, True, threads=8)
    np.testing.assert_array_equal(s.currents(), exp)


# --- exact side tables and the CLI result block ------------------------------------
def _oracle_tables(coracle, seqs, k, pool, canonical):
    words = np.concatenate([coracle.kmer_words(s, k, canonical) for s in seqs] + [np.zeros(0, np.uint64)])
    keys, counts = np.unique(words, return_counts=True)
    idx = np.array([coracle.neuron_index(int(w), pool) for w in keys], np.int64)
    return keys, counts.astype(np.uint32), np.bincount(idx, minlength=pool).astype(np.uint32)


@pytest.mark.parametrize("k,pool,canonical", [(11, 5000, True), (31, 100_000, True), (9, 4096, False)])
def test_exact_side_tables(coracle, k, pool, canonical):
    """counts / get_count / kmer_per_neuron (spiking_hash.rs:26-27,157-172,675-678) and the
    `uniques` column of top_abundant_neurons (:667)."""
    rng = np.random.default_rng(k)
    seqs = [random_dna(rng, n, 0.01, 0.05) for n in (60_000, 0, 5, 40_000, 1234)]
    c = make(k, pool, canonical)
    assert [t[2] for t in c.top_abundant_neurons(3)] == [None] * 3
    c.enable_exact_counts(True)
    c.process_parallel(seqs)
    keys, counts, uni = _oracle_tables(coracle, seqs, k, pool, canonical)
    gk, gc = c.exact_table()
    np.testing.assert_array_equal(gk, keys); np.testing.assert_array_equal(gc, counts)
    np.testing.assert_array_equal(c.kmer_per_neuron(), uni)
    for i in (0, len(keys) // 2, len(keys) - 1):
        assert c.get_count(int(keys[i])) == int(counts[i])
    missing = next(x for x in range(1, 1000) if x not in set(keys[:2000].tolist()))
    assert c.get_count(missing) is None
    top = c.top_abundant_neurons(20)
    assert [t[2] for t in top] == [int(uni[t[0]]) for t in top]
    # currents/spikes are unaffected by the side-table mode
    plain = make(k, pool, canonical); plain.process_parallel(seqs)
    np.testing.assert_array_equal(plain.currents(), c.currents())
    np.testing.assert_array_equal(plain.spike_counts(), c.spike_counts())
    # a second batch call REPLACES both tables (counts.clear(), :157)
    seqs2 = [random_dna(rng, 30_000)]
    c.process_parallel(seqs2)
    keys2, counts2, uni2 = _oracle_tables(coracle, seqs2, k, pool, canonical)
    gk, gc = c.exact_table()
    np.testing.assert_array_equal(gk, keys2); np.testing.assert_array_equal(gc, counts2)
    np.testing.assert_array_equal(c.kmer_per_neuron(), uni2)
    # streaming fills them the same way
    s = make(k, pool, canonical); s.enable_exact_counts(True)
    from neurokmer_b200 import flatten
    s.stream_begin(); s.stream_push(*flatten(seqs[:2])); s.stream_push(*flatten(seqs[2:])); s.stream_end()
    gk, gc = s.exact_table()
    np.testing.assert_array_equal(gk, keys); np.testing.assert_array_equal(gc, counts)


def test_exact_tables_empty_inputs():
    c = make(31, 1000); c.enable_exact_counts(True)
    c.process_parallel([])                         # nothing counted at all
    assert c.exact_table()[0].size == 0 and c.kmer_per_neuron().sum() == 0 and c.get_count(5) is None
    c.process_parallel([b"ACGT", b""])             # only sequences shorter than k
    assert c.exact_table()[0].size == 0
    assert c.top_abundant_neurons(3) == [(0, 0, 0), (1, 0, 0), (2, 0, 0)]
    c.process_parallel([b"A" * 40])
    keys, counts = c.exact_table()
    assert keys.tolist() == [0] and counts.tolist() == [10] and c.get_count(0) == 10


def test_exact_tables_process_sequence(coracle):
    """process_sequence ADDS to counts (:218-221) and counts, per neuron, the sequences that touched it (:262-264)."""
    rng = np.random.default_rng(8)
    k, pool = 7, 512
    c = make(k, pool, True); c.enable_exact_counts(True)
    seqs = [random_dna(rng, n, 0.02) for n in (300, 3, 150, 700)]
    total = {}
    touched = np.zeros(pool, np.uint32)
    for s in seqs:
        c.process_sequence(s)
        w = coracle.kmer_words(s, k, True)
        for x in w.tolist():
            total[x] = total.get(x, 0) + 1
        for i in set(coracle.neuron_index(int(x), pool) for x in w):
            touched[i] += 1
        gk, gc = c.exact_table()
        assert dict(zip(gk.tolist(), gc.tolist())) == total
        np.testing.assert_array_equal(c.kmer_per_neuron(), touched)


def test_python_surface(tmp_path, coracle):
    """src/python.rs:7-53: PySpikingCounter(k, pool_size).process_file / get_counts / energy_used, pack_kmer_py."""
    from neurokmer_b200 import PySpikingCounter, pack_kmer_py
    from neurokmer_b200.fastx import write_fasta
    rng = np.random.default_rng(12)
    seqs = [random_dna(rng, n, 0.01) for n in (500, 20, 800)]
    fa = str(tmp_path / "p.fa"); write_fasta(fa, seqs)
    py = PySpikingCounter(9, 256)
    py.process_file(fa)
    want = {}
    for s in seqs:
        for x in coracle.kmer_words(s, 9, False).tolist():   # python.rs builds a NON-canonical counter
            want[str(x)] = want.get(str(x), 0) + 1
    assert py.get_counts() == want
    v = np.zeros(256, np.float32); r = np.zeros(256, np.uint32); sp = np.zeros(256, np.uint64); sc = np.zeros(256, np.uint64)
    fired = sum(coracle.process_sequence(s, 9, 256, False, 1.0, 0.95, 2, sc, v, r, sp) for s in seqs)
    assert py.energy_used() == float(fired)
    assert pack_kmer_py(b"ACGTN") == 27


def test_cli_result_block(tmp_path, coracle):
    """The `neurokmer` binary prints the reference's result block (src/main.rs:49-74)."""
    import subprocess
    from neurokmer_b200.fastx import write_fasta
    from neurokmer_b200 import flatten
    from neurokmer_b200 import build as nkbuild
    nkbuild.build()  # no-op when libneurokmer.so and the CLI binary are up to date
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neurokmer_b200", "neurokmer")
    rng = np.random.default_rng(13)
    seqs = [random_dna(rng, n, 0.002, 0.01) for n in (400_000, 100, 250_000)]
    fa = str(tmp_path / "c.fa"); write_fasta(fa, seqs)
    k, pool = 21, 10_000
    # default: uniques of the printed rows by a second read of the file; --exact: the full exact table
    for streaming, extra in ((False, ["--exact"]), (True, ["--exact"]), (False, []), (True, [])):
        args = [exe, "-i", fa, "-k", str(k), "--pool-size", str(pool), "--canonical"] + extra + (["--streaming"] if streaming else [])
        out = subprocess.check_output(args, text=True)
        o = oracle_counter(k, pool)
        bases, offsets = flatten(seqs)
        (o.process_streaming([(bases, offsets)]) if streaming else o.process_parallel(bases, offsets))
        oi, os_ = o.top_abundant_neurons(20)
        _, _, uni = _oracle_tables(coracle, seqs, k, pool, True)
        want = ["", "=== Top 20 Abundant Neuron Groups (Highest Spike Rates) ==="]
        for rank, (i, s) in enumerate(zip(oi, os_)):
            want.append(f"{rank + 1:3}: Neuron {int(i):6} → {int(s):8} spikes ({int(uni[int(i)])} unique k-mers colliding)")
        want += ["", f"Total spikes fired: {o.total_spikes}", f"Simulated energy used: {o.total_spikes}",
                 f"Neuron pool size used: {pool}", f"Streaming mode: {'true' if streaming else 'false'}"]
        assert out.splitlines() == want
    # with --no-uniques the column is "n/a", never a number
    out = subprocess.check_output([exe, "-i", fa, "-k", "21", "--pool-size", "10000", "--canonical", "--no-uniques"], text=True)
    assert "(n/a unique k-mers colliding)" in out
    # errors: missing file -> exit code 1 and a message on stderr (the reference returns Err from main)
    p = subprocess.run([exe, "-i", str(tmp_path / "nope.fa")], capture_output=True, text=True)
    assert p.returncode == 1 and "cannot open" in p.stderr
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 2 and "--input" in p.stderr


# --- sharded flow (configs[4] shape) emulated on one GPU --------------------------------------------
def test_sharded_flow_two_ranks_one_gpu(coracle):
    """configs[4]: k=31, pool 16,000,000, input sharded by sequence chunk with k-1 overlap, full pool
    replica per rank, currents summed, LIF after the sum.  Two handles on one GPU stand in for two
    ranks (the all-reduce is an in-place device add through torch); the result must equal the
    unsharded run and the oracle bit for bit."""
    import torch
    from neurokmer_b200 import flatten
    from neurokmer_b200.shard import shard_batch
    from bench import CurrentsView
    rng = np.random.default_rng(5)
    k, pool = 31, 16_000_000
    seqs = [random_dna(rng, n, 0.001) for n in (1_500_000, 10, 900_000, 31, 2_000_000)]
    bases, offsets = flatten(seqs)
    ranks = [make(k, pool), make(k, pool)]
    views = []
    for r, c in enumerate(ranks):
        b, o = shard_batch(bases, offsets, k, 2, r)
        c.stream_begin(); c.stream_push(b, o)
        ptr = c.stream_accumulated(); c.synchronize()
        views.append(torch.as_tensor(CurrentsView(ptr, pool), device="cuda"))
    total = views[0] + views[1]
    for v in views:
        v.copy_(total)
    torch.cuda.synchronize()
    for c in ranks:
        c.stream_finish()
    ref = make(k, pool); ref.stream_begin(); ref.stream_push(bases, offsets); ref.stream_end()
    for c in ranks:
        np.testing.assert_array_equal(c.currents(), ref.currents())
        np.testing.assert_array_equal(c.spike_counts(), ref.spike_counts())
        assert c.energy.total_spikes() == ref.energy.total_spikes()
        assert c.top_abundant_neurons(20) == ref.top_abundant_neurons(20)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, True, threads=4)
    np.testing.assert_array_equal(ref.currents(), exp)
    assert int(exp.sum()) == tot


def test_sharded_pool_fused_reduce_two_ranks_one_gpu(coracle):
    """nk_dist_*: the reduce-scatter of the per-rank counts fused into the LIF kernel through peer
    pointers.  Two handles on one GPU stand in for two ranks (raw pointers instead of CUDA IPC; the
    all-gather of the result packs is two device copies).  Whole-pool answers (total spikes, top-N)
    must equal the single-handle run on every rank; per-neuron state must match inside each slice."""
    import torch
    from neurokmer_b200 import flatten
    from neurokmer_b200.devmem import copy_d2d
    from neurokmer_b200.shard import shard_batch
    rng = np.random.default_rng(6)
    for k, pool, world in [(31, 2_000_000, 2), (21, 100_003, 3)]:
        seqs = [random_dna(rng, n, 0.001) for n in (3_000_000, 10, 1_700_000, 31, 2_000_000)]
        bases, offsets = flatten(seqs)
        ranks = [make(k, pool) for _ in range(world)]
        raws = [c.dist_export()[1] for c in ranks]
        for r, c in enumerate(ranks):
            c.dist_setup(r, world, raw_ptrs=raws)
        ref = make(k, pool); ref.stream_begin(); ref.stream_push(bases, offsets); ref.stream_end()
        for job in range(2):  # two jobs back to back: accumulators must be clean in between
            for r, c in enumerate(ranks):
                c.reset(); c.stream_begin(); c.stream_push(*shard_batch(bases, offsets, k, world, r))
            for c in ranks:
                c.synchronize()                      # cross-rank barrier: every rank has counted
            posts = [c.dist_post() for c in ranks]
            for c in ranks:
                c.synchronize()
            n64, each = posts[0][1], posts[0][2]
            gathered = torch.zeros(world * n64, dtype=torch.int64, device="cuda")
            for r, (ptr, _, _) in enumerate(posts):
                copy_d2d(gathered.data_ptr() + 8 * n64 * r, ptr, 8 * n64)
            torch.cuda.synchronize()
            for c in ranks:
                c.dist_complete(gathered.data_ptr(), each)
            for c in ranks:
                assert c.energy.total_spikes() == ref.energy.total_spikes()
                assert c.energy_used() == ref.energy_used()
                assert c.top_abundant_neurons(20) == ref.top_abundant_neurons(20)
                assert c.top_abundant_neurons(7) == ref.top_abundant_neurons(7)
                assert c.timings()["kmers"] == ref.timings()["kmers"]
                lo, ln = c.dist_slice()
                np.testing.assert_array_equal(c.currents()[lo:lo + ln], ref.currents()[lo:lo + ln])
                np.testing.assert_array_equal(c.spike_counts()[lo:lo + ln], ref.spike_counts()[lo:lo + ln])
                np.testing.assert_array_equal(c.refractory_ticks()[lo:lo + ln], ref.refractory_ticks()[lo:lo + ln])
                np.testing.assert_array_equal(c.voltages()[lo:lo + ln].view(np.uint32), ref.voltages()[lo:lo + ln].view(np.uint32))
        assert sum(c.dist_slice()[1] for c in ranks) == pool


def test_config3_full_size_two_slices():
    """configs[2] at FULL size on device: 50 M reads x 150 bp = 7.49 Gbase > 2^32 window starts, so the
    staged batch is counted in two launches with a fold in between.  Properties: the k-mer total and the
    checksum of the currents equal Σ max(0, L-k+1); all neurons saturate (334 spikes)."""
    from neurokmer_b200.devmem import copy_h2d
    nreads, k, pool = 50_000_000, 31, 2_000_000
    lens = np.full(nreads, 150, np.int64); lens[::1000] = 20
    offs = np.zeros(nreads + 1, np.uint64); offs[1:] = np.cumsum(lens)
    n = int(offs[-1])
    assert n > 2**32
    c = make(k, pool)
    db, do = c.stage_reserve(n, nreads)
    c.synth_fill(db, 3, 0, n, 0); copy_h2d(do, offs); c.synchronize()
    c.process_staged(n, nreads, 0)
    want = int(np.maximum(lens - k + 1, 0).sum())
    assert c.timings()["kmers"] == want
    cur = c.currents()
    assert int(cur.sum()) == want
    assert set(c.spike_counts().tolist()) == {334} and c.energy.total_spikes() == 334 * pool


def test_exact_modulo_all_pool_sizes():
    """`finish() % pool_size` (spiking_hash.rs:81) for pool sizes up to 2^32-1 and adversarial values
    (exact multiples, multiples ± 1, extremes): the two-stage FP64-pipe routine, the integer routine and the
    one-stage routine the count kernel picks per pool size (split at bit 32 or 44) all equal Python's integer
    remainder."""
    from neurokmer_b200.counter import debug_mod
    rng = np.random.default_rng(99)
    pools = [1, 2, 3, 4, 5, 6, 7, 10, 1000, 4096, 65535, 65536, 65537, 999_983, 1_000_000, 2_000_000, 15_625, 16_000_000,
             2**21 - 2, 2**21 - 1, 2**21 + 1, 2_097_153, 2**32 // 3, 2**32 // 3 + 1, 1_431_655_766,
             2**24 - 1, 2**31 - 1, 2**31, 2**31 + 1, 4_294_967_291, 2**32 - 2, 2**32 - 1] + \
            [int(x) for x in rng.integers(1, 2**32, size=40, dtype=np.uint64)]
    for p in pools:
        v = rng.integers(0, 2**64, size=20000, dtype=np.uint64)
        q = rng.integers(0, (2**64 - 1) // p + 1, size=6000, dtype=np.uint64)
        mult = q * np.uint64(p)
        v[:6000] = mult                                   # exact multiples
        v[6000:12000] = mult + np.uint64(p - 1)           # one below the next multiple (may wrap: still valid input)
        v[12000:12010] = [0, 1, 2**64 - 1, 2**64 - 2, 2**63, 2**32, 2**32 - 1, p, p - 1 if p > 1 else 0, (p * 4096) % 2**64]
        v[12010:14000] |= np.uint64(0xFFFFFFFF00000000)   # large high words: the one-stage routine's x near 2^53
        v[14000:16000] |= np.uint64(0xFFFFF00000000000)
        want = np.array([int(x) % p for x in v], np.uint64)
        for which in (0, 1, 2):
            np.testing.assert_array_equal(debug_mod(v, p, which), want, err_msg=f"pool {p} routine {which}")


@pytest.mark.parametrize("k,pool,canonical", [(31, 2_000_000, False), (11, 65536, True), (16, 4099, False), (32, 1_000_003, True), (2, 7, True)])
def test_short_read_compaction_mode(coracle, k, pool, canonical):
    """Short-read batches (mean length < 2 KiB) take the warp-compaction instantiation of the count
    kernel: every template corner of it (non-canonical skip rule, k <= 16, power-of-two pool, k = 32)
    against the oracle, with N-rich reads, reads shorter than k and empty reads."""
    rng = np.random.default_rng(1000 + k)
    lens = rng.choice([0, 1, k - 1, k, k + 1, 36, 75, 100, 150, 151, 250, 511, 512, 513], size=6000)
    seqs = [random_dna(rng, int(n), 0.03, 0.1, 0.01) for n in lens]
    from neurokmer_b200 import flatten
    bases, offsets = flatten(seqs)
    c = make(k, pool, canonical)
    c.process_batch(bases, offsets)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, canonical, threads=4)
    np.testing.assert_array_equal(c.currents(), exp)
    assert c.timings()["kmers"] == tot == int(np.maximum(lens - k + 1, 0).sum())
    # the same reads as ONE long batch with exact tables on (word-append instantiation) agree too
    e = make(k, pool, canonical); e.enable_exact_counts(True); e.process_batch(bases, offsets)
    np.testing.assert_array_equal(e.currents(), exp)
    assert int(e.exact_table()[1].sum()) == tot


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_api_call_sequence_fuzz(coracle, seed):
    """Random sequences of the public entry points against the oracle's mirror of the reference's
    state rules: currents overwritten per batch call, neuron state / energy carried over, streaming vs
    in-memory LIF drivers, process_sequence's single tick, simulate_spikes_auto on stored currents,
    set_steps, reset, read-outs in between (they must not perturb anything)."""
    from neurokmer_b200 import flatten
    rng = np.random.default_rng(seed)
    k = int(rng.choice([5, 15, 21, 31]))
    pool = int(rng.choice([64, 1000, 5003, 65536]))
    canonical = bool(rng.integers(0, 2))
    c = make(k, pool, canonical)
    o = oracle_counter(k, pool, canonical)
    scratch = np.zeros(pool, np.uint64)

    def batch():
        n = int(rng.integers(0, 5))
        seqs = [random_dna(rng, int(rng.choice([0, 3, k, 40, 700, 5000, 60000])), 0.01, 0.05) for _ in range(n)]
        return flatten(seqs)

    for step in range(24):
        op = rng.choice(["batch", "stream", "sequence", "simulate", "steps", "reset", "readout"],
                        p=[0.28, 0.22, 0.14, 0.08, 0.08, 0.08, 0.12])
        if op == "batch":
            b, off = batch()
            c.process_batch(b, off); o.process_parallel(b, off)
        elif op == "stream":
            batches = [batch() for _ in range(int(rng.integers(0, 4)))]
            c.stream_begin()
            for b, off in batches:
                c.stream_push(b, off)
            c.stream_end()
            o.process_streaming(batches)
        elif op == "sequence":
            s = random_dna(rng, int(rng.choice([2, k, 50, 400])), 0.02)
            c.process_sequence(s)
            # process_sequence adds to the stored currents, ticks once, zeroes them (spiking_hash.rs:203-273)
            scratch[:] = o.currents
            fired = coracle.process_sequence(s, k, pool, canonical, 1.0, 0.95, 2, scratch, o.v, o.r, o.spikes)
            if len(s) >= k:
                o.currents = scratch.copy()
            o.total_spikes += fired
            o.energy_fixed += fired * 1000
        elif op == "simulate":
            c.simulate_spikes_auto(); o._simulate(simd=True)
        elif op == "steps":
            st = int(rng.choice([0, 1, 10, 1000]))
            c.set_steps(st); o.steps = st
        elif op == "reset":
            c.reset()
            o.v[:] = 0; o.r[:] = 0; o.spikes[:] = 0; o.currents[:] = 0; o.total_spikes = 0; o.energy_fixed = 0
        else:
            assert_topn_equal(c, o, int(rng.choice([1, 20, 64])))
            assert c.energy.total_spikes() == o.total_spikes
        if rng.random() < 0.5:
            assert_state_equal(c, o)
    assert_state_equal(c, o)
    assert_topn_equal(c, o, 20)


@pytest.mark.parametrize("seed,window,threads", [(1, 64, 4), (2, 97, 3), (3, 256, 8), (4, 1000, 2), (5, 4096, 16)])
def test_parallel_fasta_ingest_fuzz(tmp_path, coracle, seed, window, threads, monkeypatch):
    """The parallel FASTA ingest (windows stripped by several host threads, stitched with a k-1
    carry) against the serial Python reader + oracle, with tiny windows so that every boundary case
    occurs: lines and headers longer than a window, records spanning many windows, empty records,
    blank lines, CRLF, no trailing newline, '>' inside sequence lines."""
    from neurokmer_b200 import flatten
    from neurokmer_b200.fastx import read_fastx
    rng = np.random.default_rng(seed)
    parts = []
    for rec in range(int(rng.integers(3, 40))):
        hdr = b">r%d " % rec + bytes(rng.choice(np.frombuffer(b"abc >xyz", np.uint8), size=int(rng.choice([0, 5, 300])))).replace(b"\n", b"")
        parts.append(hdr + (b"\r\n" if seed % 2 else b"\n"))
        n = int(rng.choice([0, 1, 30, 31, 200, 5000]))
        seq = random_dna(rng, n, 0.02, 0.05)
        width = int(rng.choice([1, 7, 60, 61, 10**6]))
        for j in range(0, n, width):
            line = seq[j:j + width]
            if rng.random() < 0.02 and len(line) > 3:
                line = line[:2] + b">" + line[3:]          # '>' that is NOT at a line start is sequence data
            parts.append(line + (b"\r\n" if seed % 2 else b"\n"))
        if rng.random() < 0.2:
            parts.append(b"\n")                             # blank line
    blob = b"".join(parts)
    if seed % 3 == 0:
        blob = blob.rstrip(b"\r\n")                         # no trailing newline
    fa = str(tmp_path / "fuzz.fa")
    with open(fa, "wb") as f:
        f.write(blob)
    seqs = list(read_fastx(fa))
    bases, offsets = flatten(seqs)
    for k, pool in ((31, 5003), (5, 64), (1, 7)):
        exp, tot = coracle.accumulate(bases, offsets, k, pool, True)
        monkeypatch.setenv("NK_FASTA_THREADS", "1")        # serial reader
        s = make(k, pool); s.process_file_streaming(fa)
        np.testing.assert_array_equal(s.currents(), exp)
        monkeypatch.setenv("NK_FASTA_THREADS", str(threads))
        monkeypatch.setenv("NK_FASTA_WINDOW", str(window))
        p = make(k, pool); p.process_file_streaming(fa)
        np.testing.assert_array_equal(p.currents(), exp)
        assert p.timings()["kmers"] == tot
        monkeypatch.delenv("NK_FASTA_WINDOW")


@pytest.mark.parametrize("seed,window,threads,defect", [(1, 64, 4, None), (2, 200, 3, "qual"), (3, 333, 8, "plus"),
                                                        (4, 1000, 2, "trunc"), (5, 97, 5, None), (6, 150, 4, "blank")])
def test_parallel_fastq_ingest_fuzz(tmp_path, coracle, seed, window, threads, defect, monkeypatch):
    """Parallel FASTQ ingest (record starts from the line number modulo 4, windows parsed by several
    host threads) against the serial reader: reads of many lengths, CRLF, no trailing newline, quality
    lines that start with '@' or '+', and a malformed record somewhere in the middle (the iteration must
    end exactly there, utils.rs:17-20)."""
    from neurokmer_b200 import flatten
    from neurokmer_b200.fastx import read_fastx
    rng = np.random.default_rng(seed)
    nrec = int(rng.integers(60, 400))
    bad_at = int(rng.integers(5, nrec - 5)) if defect else -1
    eol = b"\r\n" if seed % 2 else b"\n"
    parts = []
    for i in range(nrec):
        n = int(rng.choice([0, 1, 20, 31, 75, 150, 151, 400]))
        seq = random_dna(rng, n, 0.02, 0.02)
        qual = bytes(rng.choice(np.frombuffer(b"@+I#5>", np.uint8), size=n))   # '@' and '+' may start a quality line
        rec = [b"@r%d" % i, seq, b"+", qual]
        if i == bad_at:
            if defect == "qual":
                rec[3] = qual + b"I"
            elif defect == "plus":
                rec[2] = b"-"
            elif defect == "blank":
                parts.append(eol)
        parts.append(eol.join(rec) + eol)
        if i == bad_at and defect == "trunc":
            parts[-1] = eol.join(rec[:2]) + eol
    blob = b"".join(parts)
    if seed % 3 == 0:
        blob = blob[: len(blob) - len(eol)]
    fq = str(tmp_path / "fuzz.fq")
    with open(fq, "wb") as f:
        f.write(blob)
    k, pool = 31, 5003
    monkeypatch.setenv("NK_FASTA_THREADS", "1")
    s = make(k, pool); s.process_file_streaming(fq)
    seqs = list(read_fastx(fq))
    if defect not in ("trunc",):
        assert len(seqs) == (bad_at if defect else nrec)
    bases, offsets = flatten(seqs)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, True)
    np.testing.assert_array_equal(s.currents(), exp)
    monkeypatch.setenv("NK_FASTA_THREADS", str(threads))
    monkeypatch.setenv("NK_FASTA_WINDOW", str(window))
    p = make(k, pool); p.process_file_streaming(fq)
    np.testing.assert_array_equal(p.currents(), exp)
    assert p.timings()["kmers"] == tot


# --- pre-packed input ("nk2" layout): same results as the ASCII entry points ---------------------
@pytest.mark.parametrize("k", [1, 2, 5, 15, 16, 17, 21, 31, 32])
@pytest.mark.parametrize("canonical", [True, False])
def test_packed_kmer_words_and_indices(coracle, k, canonical):
    """The pre-packed windowing stage (convert16_packed) yields the words of models.rs:206-286 /
    utils.rs:26-39 for the bytes the packed form was made from — `other` bases are code 0 on both
    strands (canonical) and skipped (pack_kmer)."""
    from neurokmer_b200 import pack_bases
    rng = np.random.default_rng(300 + k)
    pool = 1_000_003 if k % 2 else 65536
    c = make(k, pool, canonical)
    for n, pn, pl, pi in [(k, 0, 0, 0), (k + 1, 0.2, 0, 0), (700, 0.05, 0.2, 0.02), (20000, 0.01, 0.01, 0.01),
                          (16384 + 40, 0.0, 0.0, 0.0), (33000, 0.3, 0.3, 0.1), (4096, 0.0, 0.0, 0.0), (4097, 0.5, 0, 0)]:
        s = random_dna(rng, n, pn, pl, pi)
        codes, other, n_other = pack_bases(s)
        ow = coracle.kmer_words(s, k, canonical)
        for oth in ([other] if n_other else [other, None]):
            fwd, rc, words, idx = c.debug_kmers_packed(codes, oth, n)
            assert words.size == ow.size == n - k + 1
            np.testing.assert_array_equal(words, ow)
            if canonical:
                of, orc = coracle.kmer_fwd_rc(s, k)
                np.testing.assert_array_equal(fwd, of)
                np.testing.assert_array_equal(rc, orc)
            step = max(1, ow.size // 3000)
            np.testing.assert_array_equal(idx[::step], coracle.indices(ow[::step], pool))
    assert c.debug_kmers_packed(np.zeros(1, np.uint32), None, k - 1)[2].size == 0


def test_packed_ignores_codes_under_other_bits(coracle):
    """A set `other` bit wins over whatever code the producer left there (well-defined for any input)."""
    from neurokmer_b200 import pack_bases
    rng = np.random.default_rng(5)
    s = bytearray(random_dna(rng, 5000, 0.0, 0.0))
    codes, other, _ = pack_bases(bytes(s))
    for p in (0, 17, 31, 32, 4095, 4096, 4999):
        other[p >> 5] |= np.uint32(1 << (p & 31))  # codes keep the letter's code
        s[p] = ord("N")
    for canonical in (True, False):
        c = make(21, 1000, canonical)
        words = c.debug_kmers_packed(codes, other, len(s))[2]
        np.testing.assert_array_equal(words, coracle.kmer_words(bytes(s), 21, canonical))


@pytest.mark.parametrize("k,pool,canonical", [(21, 1_000_000, True), (31, 2_000_000, True), (15, 65536, True),
                                              (32, 999_983, True), (31, 1_000_000, False), (7, 1000, False), (1, 3, True)])
def test_packed_currents_ragged(coracle, k, pool, canonical):
    """nk_process_batch_packed on a ragged batch (sequences are not word-aligned in the packed form)."""
    from neurokmer_b200 import flatten, pack_bases
    rng = np.random.default_rng(k * 1000 + pool % 997 + 1)
    seqs = ragged_batch(rng, k)
    bases, offsets = flatten(seqs)
    codes, other, _ = pack_bases(bases)
    c = make(k, pool, canonical)
    c.process_batch_packed(codes, other, offsets)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, canonical, threads=4)
    got = c.currents()
    np.testing.assert_array_equal(got, exp)
    assert int(got.sum()) == tot and c.timings()["kmers"] == tot
    # the ASCII entry point on the same counter: identical state rules (currents overwritten)
    c2 = make(k, pool, canonical)
    c2.process_batch(bases, offsets)
    np.testing.assert_array_equal(c2.currents(), got)
    np.testing.assert_array_equal(c2.spike_counts(), c.spike_counts())


def test_packed_short_reads_and_no_other(coracle):
    """150 bp reads (compaction mode of the count kernel) in packed form; all-ACGT batch with other=NULL."""
    from neurokmer_b200 import flatten, pack_bases
    rng = np.random.default_rng(77)
    seqs = [random_dna(rng, int(n), 0.0, 0.02) for n in rng.choice([150, 150, 150, 20, 151, 31, 30], size=6000)]
    bases, offsets = flatten(seqs)
    codes, other, n_other = pack_bases(bases)
    assert n_other == 0
    k, pool = 31, 2_000_000
    exp, tot = coracle.accumulate(bases, offsets, k, pool, True, threads=4)
    for oth in (other, None):
        c = make(k, pool)
        c.process_batch_packed(codes, oth, offsets)
        np.testing.assert_array_equal(c.currents(), exp)
    seqs = [random_dna(rng, 150, 0.01, 0.02, 0.01) for _ in range(3000)]
    bases, offsets = flatten(seqs)
    codes, other, n_other = pack_bases(bases)
    assert n_other > 0
    exp, tot = coracle.accumulate(bases, offsets, k, pool, True, threads=4)
    c = make(k, pool)
    c.process_batch_packed(codes, other, offsets)
    np.testing.assert_array_equal(c.currents(), exp)


def test_packed_streaming_full_pipeline_and_chunk_straddle(coracle, monkeypatch):
    """nk_stream_push_packed across several pushes and several H2D chunks (1 Mbase granules here),
    then LIF + totals + top-N: identical to the oracle and to the ASCII stream."""
    from neurokmer_b200 import flatten, pack_bases
    monkeypatch.setenv("NK_PACKED_CHUNK_MBASES", "1")
    rng = np.random.default_rng(99)
    k, pool = 31, 20_000
    batches = []
    for lens in ([1_500_000, 700_000, 40], [(1 << 20) - 15, 31, 1 << 20, 100], [2_200_000]):
        seqs = [random_dna(rng, n, 0.002, 0.01, 0.001) for n in lens]
        batches.append(flatten(seqs))
    c = make(k, pool)
    c.stream_begin()
    for bases, offsets in batches:
        codes, other, _ = pack_bases(bases, threads=2)
        c.stream_push_packed(codes, other, offsets)
    c.stream_end()
    o = oracle_counter(k, pool)
    o.process_streaming(batches)
    assert_state_equal(c, o)
    assert_topn_equal(c, o, 20)
    c2 = make(k, pool)
    c2.stream_begin()
    for bases, offsets in batches:
        c2.stream_push(bases, offsets)
    c2.stream_end()
    np.testing.assert_array_equal(c2.spike_counts(), c.spike_counts())
    assert c2.top_abundant_neurons(20) == c.top_abundant_neurons(20)


def test_packed_exact_tables_and_state_errors(coracle):
    """Exact side tables fed from packed input; call-sequence errors mirror the ASCII entry points."""
    from neurokmer_b200 import NkError, flatten, pack_bases
    rng = np.random.default_rng(3)
    seqs = [random_dna(rng, 30_000, 0.01, 0.02), random_dna(rng, 500, 0.0, 0.0)]
    bases, offsets = flatten(seqs)
    codes, other, _ = pack_bases(bases)
    k, pool = 15, 4096
    c = make(k, pool)
    c.enable_exact_counts(True)
    c.process_batch_packed(codes, other, offsets)
    words = np.concatenate([coracle.kmer_words(s, k, True) for s in seqs])
    keys, counts = np.unique(words, return_counts=True)
    gk, gc = c.exact_table()
    np.testing.assert_array_equal(gk, keys)
    np.testing.assert_array_equal(gc, counts.astype(np.uint32))
    with pytest.raises(NkError) as ei:
        c.stream_push_packed(codes, other, offsets)
    assert ei.value.code == 6  # NK_ERR_STATE
    c.stream_begin()
    with pytest.raises(NkError):
        c.process_batch_packed(codes, other, offsets)
    c.stream_end()
    bad = offsets.copy(); bad[0] = 1
    with pytest.raises(NkError) as ei:
        c.process_batch_packed(codes, other, bad)
    assert ei.value.code == 1


def test_packed_staged_device_resident(coracle):
    """nk_stage_reserve_packed / nk_process_staged_packed: device-resident packed batch."""
    import torch
    from neurokmer_b200 import flatten, pack_bases
    from neurokmer_b200.devmem import copy_h2d
    rng = np.random.default_rng(11)
    seqs = [random_dna(rng, n, 0.003, 0.01) for n in (300_000, 77, 150_000)]
    bases, offsets = flatten(seqs)
    codes, other, _ = pack_bases(bases)
    k, pool = 31, 50_000
    c = make(k, pool)
    dc, dx, do = c.stage_reserve_packed(bases.size, len(seqs))
    copy_h2d(dc, codes); copy_h2d(dx, other); copy_h2d(do, offsets)
    c.process_staged_packed(bases.size, len(seqs), 0, True)
    o = oracle_counter(k, pool)
    o.process_parallel(bases, offsets)
    assert_state_equal(c, o)


def test_packed_zero_copy_from_pinned_memory(coracle, monkeypatch):
    """Packed arrays in pinned (device-mapped) host memory: the count kernel reads whole tiles
    straight from host memory, the ragged end is staged; same result as pageable input and oracle."""
    from neurokmer_b200 import PinnedBuffer, flatten, pack_bases
    rng = np.random.default_rng(21)
    k, pool = 31, 100_000
    seqs = [random_dna(rng, n, 0.004, 0.01, 0.001) for n in (700_001, 33, 90_000, 1_234_567)]
    bases, offsets = flatten(seqs)
    pc = PinnedBuffer(4 * ((bases.size + 15) // 16), np.uint32)
    po = PinnedBuffer(4 * ((bases.size + 31) // 32), np.uint32)
    pack_bases(bases, out_codes=pc.array, out_other=po.array)
    exp, tot = coracle.accumulate(bases, offsets, k, pool, True, threads=4)
    for zc in ("1", "0"):
        monkeypatch.setenv("NK_ZEROCOPY", zc)
        c = make(k, pool)
        c.process_batch_packed(pc.array, po.array, offsets)
        np.testing.assert_array_equal(c.currents(), exp)
        assert c.timings()["kmers"] == tot
        # staged H2D moves whole chunks; the zero-copy body is one launch + the staged ragged end
        assert c.timings()["h2d_bytes"] > 0
    o = oracle_counter(k, pool)
    o.process_parallel(bases, offsets)
    assert_state_equal(c, o)
    # a batch too small for a zero-copy body, and one that ends exactly on a tile boundary
    for n in (5000, 4 * 4096 + 128, 8 * 4096):
        s = random_dna(rng, n, 0.01, 0.0)
        b2, o2 = flatten([s])
        pack_bases(b2, out_codes=pc.array, out_other=po.array)
        monkeypatch.setenv("NK_ZEROCOPY", "1")
        c = make(k, pool)
        c.process_batch_packed(pc.array[: (n + 15) // 16], po.array[: (n + 31) // 32], o2)
        e2, _ = coracle.accumulate(b2, o2, k, pool, True, threads=2)
        np.testing.assert_array_equal(c.currents(), e2)


def test_sharded_pool_peer_signalled_run_one_gpu(coracle):
    """nk_dist_run: counting-finished flags, count reduce-scatter and result-pack exchange all through peer
    memory, no host barrier between the ranks.  Two / three handles on one GPU stand in for ranks (small pools:
    the waiting kernels of all handles must be co-resident on the one device)."""
    from neurokmer_b200 import flatten
    from neurokmer_b200.shard import shard_batch
    rng = np.random.default_rng(16)
    for k, pool, world in [(31, 40_000, 2), (21, 30_011, 3)]:
        seqs = [random_dna(rng, n, 0.001) for n in (1_500_000, 10, 700_000, 31, 900_000)]
        bases, offsets = flatten(seqs)
        ranks = [make(k, pool) for _ in range(world)]
        raws = [c.dist_export()[1] for c in ranks]
        for r, c in enumerate(ranks):
            c.dist_setup(r, world, raw_ptrs=raws)
        ref = make(k, pool); ref.stream_begin(); ref.stream_push(bases, offsets); ref.stream_end()
        for job in range(3):  # back to back: epochs advance, accumulators are clean in between
            for r, c in enumerate(ranks):
                c.reset(); c.stream_begin(); c.stream_push(*shard_batch(bases, offsets, k, world, r))
            for c in ranks:      # no barrier of any kind between the ranks
                c.dist_run()
            for c in ranks:
                assert c.energy.total_spikes() == ref.energy.total_spikes()
                assert c.top_abundant_neurons(20) == ref.top_abundant_neurons(20)
                assert c.timings()["kmers"] == ref.timings()["kmers"]
                lo, ln = c.dist_slice()
                np.testing.assert_array_equal(c.currents()[lo:lo + ln], ref.currents()[lo:lo + ln])
                np.testing.assert_array_equal(c.spike_counts()[lo:lo + ln], ref.spike_counts()[lo:lo + ln])
        for c in ranks:
            c.close()


def test_peer_signalled_run_times_out_on_a_missing_rank(monkeypatch):
    """A rank that never calls nk_dist_run must not hang the others: NK_ERR_STATE after the timeout."""
    from neurokmer_b200 import NkError, flatten
    monkeypatch.setenv("NK_DIST_TIMEOUT_MS", "300")
    rng = np.random.default_rng(17)
    bases, offsets = flatten([random_dna(rng, 200_000)])
    ranks = [make(31, 20_000) for _ in range(2)]
    raws = [c.dist_export()[1] for c in ranks]
    for r, c in enumerate(ranks):
        c.dist_setup(r, 2, raw_ptrs=raws)
    c = ranks[0]
    c.reset(); c.stream_begin(); c.stream_push(bases, offsets)
    c.dist_run()
    with pytest.raises(NkError) as ei:
        c.energy.total_spikes()
    assert ei.value.code == 6 and "timed out" in str(ei.value)
    c.reset(); c.stream_begin()
    with pytest.raises(NkError):
        c.dist_run()       # sticky until the ranks set up again


# --- uniques of the top rows by a second pass (no exact table) ---------------------------------------
@pytest.mark.parametrize("k,pool,canonical", [(21, 10_000, True), (31, 2_000_000, True), (15, 4096, False), (5, 7, True)])
def test_uniques_pass_matches_exact_tables(coracle, k, pool, canonical):
    """nk_uniques_begin/push/end: kmer_per_neuron (spiking_hash.rs:167-172) of the top rows from a second pass over the
    same batches equals the oracle's table and the exact-table mode; ASCII and packed re-supply, several pushes.
    pool 7: every window maps to a requested row, so the word array has to grow (> 2^20 matches)."""
    from neurokmer_b200 import NkError, flatten, pack_bases
    rng = np.random.default_rng(k * 7 + pool % 13)
    seqs = [random_dna(rng, n, 0.003, 0.01) for n in (900_000, 100, 450_000, 20)]
    b1, o1 = flatten(seqs[:2]); b2, o2 = flatten(seqs[2:])
    _, _, uni = _oracle_tables(coracle, seqs, k, pool, canonical)
    c = make(k, pool, canonical)
    c.stream_begin(); c.stream_push(b1, o1); c.stream_push(b2, o2); c.stream_end()
    n = min(20, pool)
    plain = c.top_abundant_neurons(n)
    assert all(r[2] is None for r in plain)
    rows = c.top_uniques(n, [(b1, o1), (b2, o2)])
    assert [r[:2] for r in rows] == [r[:2] for r in plain]
    assert [r[2] for r in rows] == [int(uni[r[0]]) for r in rows]
    assert [r[2] for r in c.top_abundant_neurons(min(5, n))] == [int(uni[r[0]]) for r in rows[:min(5, n)]]
    # packed re-supply, one push
    bases, offsets = flatten(seqs)
    codes, other, _ = pack_bases(bases)
    c.uniques_begin(n); c.uniques_push_packed(codes, other, offsets); c.uniques_end()
    assert c.top_abundant_neurons(n) == rows
    # the exact-table mode agrees
    cx = make(k, pool, canonical); cx.enable_exact_counts(True)
    cx.stream_begin(); cx.stream_push(bases, offsets); cx.stream_end()
    assert cx.top_abundant_neurons(n) == rows
    # a new job invalidates the rows; call-sequence errors
    c.process_batch(b1, o1)
    assert all(r[2] is None for r in c.top_abundant_neurons(n))
    with pytest.raises(NkError):
        c.uniques_push(b1, o1)
    with pytest.raises(NkError):
        c.uniques_end()
    c.uniques_begin(n)
    with pytest.raises(NkError):
        c.uniques_begin(n)
    c.uniques_end()
    with pytest.raises(NkError):
        c.get_count(0)  # the exact table is off: the second pass does not make get_count answerable


def test_process_file_fills_uniques_by_second_read(tmp_path, coracle):
    """nk_set_file_uniques: nk_process_file reads the file twice and the top rows carry `uniques`
    (FASTA and gzip FASTQ; the serial reader and the small-window parallel reader)."""
    import gzip
    from neurokmer_b200.fastx import write_fasta, write_fastq
    rng = np.random.default_rng(31)
    seqs = [random_dna(rng, n, 0.002, 0.02) for n in (300_000, 61, 0, 200_000)]
    fa, fq = str(tmp_path / "u.fa"), str(tmp_path / "u.fq")
    write_fasta(fa, seqs); write_fastq(fq, seqs)
    with open(fq, "rb") as f, gzip.open(fq + ".gz", "wb", compresslevel=1) as g:
        g.write(f.read())
    k, pool = 31, 50_000
    _, _, uni = _oracle_tables(coracle, seqs, k, pool, True)
    for path in (fa, fq + ".gz"):
        c = make(k, pool); c.set_file_uniques(20)
        c.process_file_streaming(path)
        rows = c.top_abundant_neurons(20)
        assert [r[2] for r in rows] == [int(uni[r[0]]) for r in rows]
        c2 = make(k, pool); c2.process_file_streaming(path)
        assert [r[:2] for r in c2.top_abundant_neurons(20)] == [r[:2] for r in rows]


def test_ascii_zero_copy_from_pinned_memory(coracle, monkeypatch):
    """ASCII batches in pinned (device-mapped) host memory are read in place by the count kernel (whole tiles;
    the ragged end is staged); pageable memory and NK_ZEROCOPY=0 take the staged pipeline.  Same results."""
    from neurokmer_b200 import PinnedBuffer, flatten
    rng = np.random.default_rng(23)
    k, pool = 31, 100_000
    seqs = [random_dna(rng, n, 0.004, 0.01, 0.001) for n in (700_001, 33, 90_000, 1_234_567, 150, 150, 20)]
    bases, offsets = flatten(seqs)
    pb = PinnedBuffer(bases.size)
    pb.array[:] = bases
    o = oracle_counter(k, pool)
    o.process_streaming([(bases, offsets)])
    for zc in ("1", "0"):
        monkeypatch.setenv("NK_ZEROCOPY", zc)
        for src in (pb.array, bases):
            c = make(k, pool)
            c.stream_begin(); c.stream_push(src, offsets); c.stream_end()
            assert_state_equal(c, o)
            assert c.timings()["kmers"] == sum(max(0, len(s) - k + 1) for s in seqs)
    # unaligned view of pinned memory (not 16-byte aligned): staged path, same answer
    monkeypatch.setenv("NK_ZEROCOPY", "1")
    pb2 = PinnedBuffer(bases.size + 16)
    pb2.array[3:3 + bases.size] = bases
    c = make(k, pool)
    c.stream_begin(); c.stream_push(pb2.array[3:3 + bases.size], offsets); c.stream_end()
    assert_state_equal(c, o)
    # short-read batch in pinned memory (compaction mode of the kernel over the zero-copy body)
    reads = [random_dna(rng, 150, 0.01, 0.0) for _ in range(4000)]
    rb, ro = flatten(reads)
    pr = PinnedBuffer(rb.size); pr.array[:] = rb
    c = make(k, pool); c.process_batch(pr.array, ro)
    exp, _ = coracle.accumulate(rb, ro, k, pool, True, threads=4)
    np.testing.assert_array_equal(c.currents(), exp)


def test_zero_copy_ragged_ends(coracle):
    """Zero-copy input whose end falls anywhere relative to a tile / a 16-byte line: the last tile's bulk copies
    are clamped to the end of the caller's pinned array (exact-size allocations), ASCII and packed."""
    from neurokmer_b200 import PinnedBuffer, flatten, pack_bases
    rng = np.random.default_rng(29)
    k, pool = 31, 65536
    T = 4096
    for n in (4 * T, 4 * T + 1, 4 * T + 15, 4 * T + 16, 4 * T + 17, 5 * T - 1, 5 * T + 31, 5 * T + 33, 6 * T + 64, 7 * T + 127, 7 * T + 129,
              16 * T + 4095):
        s = random_dna(rng, n, 0.01, 0.0)
        b, o = flatten([s[: n // 3], s[n // 3:]])
        exp, _ = coracle.accumulate(b, o, k, pool, True, threads=2)
        pa = PinnedBuffer(n); pa.array[:] = b
        c = make(k, pool); c.process_batch(pa.array, o)
        np.testing.assert_array_equal(c.currents(), exp, err_msg=f"ASCII n={n}")
        pc = PinnedBuffer(4 * ((n + 15) // 16), np.uint32); po = PinnedBuffer(4 * ((n + 31) // 32), np.uint32)
        pack_bases(b, out_codes=pc.array, out_other=po.array)
        c = make(k, pool); c.process_batch_packed(pc.array, po.array, o)
        np.testing.assert_array_equal(c.currents(), exp, err_msg=f"packed n={n}")
        c = make(k, pool, False); c.process_batch_packed(pc.array, po.array, o)
        exp_nc, _ = coracle.accumulate(b, o, k, pool, False, threads=2)
        np.testing.assert_array_equal(c.currents(), exp_nc, err_msg=f"packed non-canonical n={n}")


def _inject(c, counts, streaming=True):
    """one job whose per-neuron totals are `counts` (written straight into the device currents)"""
    from neurokmer_b200.devmem import copy_h2d
    c.stream_begin()
    ptr = c.stream_accumulated()
    c.synchronize()
    copy_h2d(ptr, np.ascontiguousarray(counts, np.uint64))
    c.stream_finish()


def test_lif_memoised_carried_state_equals_direct(coracle):
    """Carried state (second and later jobs on one counter): one simulation per distinct (state, count) key gives the
    state the per-neuron simulation gives, bit for bit, and both equal the oracle (models.rs:34-51 over
    spiking_hash.rs:544-659); a pool too diverse for the key table falls back to the direct kernel on the device."""
    rng = np.random.default_rng(41)
    pool = 600_000
    a, b = make(31, pool), make(31, pool)
    b.debug_set_lif_path(1)
    o = oracle_counter(31, pool)
    for job in range(4):
        counts = rng.poisson(56 if job % 2 == 0 else 900, size=pool).astype(np.uint64)
        counts[rng.integers(0, pool, 1000)] = 0
        counts[rng.integers(0, pool, 1000)] = rng.integers(1000, 5000, 1000)   # saturating counts
        _inject(a, counts); _inject(b, counts)
        o.currents = counts.copy(); o._simulate(True)
        assert a.timings()["lif_path"] == (3 if job == 0 else 5), job
        assert b.timings()["lif_path"] == 1
        for c in (a, b):
            np.testing.assert_array_equal(c.spike_counts(), o.spikes)
            np.testing.assert_array_equal(c.refractory_ticks(), o.r)
            np.testing.assert_array_equal(c.voltages().view(np.uint32), o.v.view(np.uint32))
            assert c.energy.total_spikes() == o.total_spikes
        assert a.top_abundant_neurons(50) == b.top_abundant_neurons(50)
    # nk_simulate on the stored currents goes the same way
    a.simulate_spikes_auto(); b.simulate_spikes_auto(); o._simulate(True)
    np.testing.assert_array_equal(a.spike_counts(), o.spikes)
    np.testing.assert_array_equal(a.voltages().view(np.uint32), b.voltages().view(np.uint32))
    # too diverse: > 2^19 distinct (state, count) keys -> nothing is applied by the memo kernels, the direct kernel runs
    pool = 1_200_000
    a, b = make(31, pool), make(31, pool)
    b.debug_set_lif_path(1)
    i = np.arange(pool, dtype=np.uint64)
    for counts in (i % 1000, (i // 1000) % 1000, (i * 7) % 1000):
        _inject(a, counts); _inject(b, counts)
    assert a.timings()["lif_path"] == 5
    np.testing.assert_array_equal(a.spike_counts(), b.spike_counts())
    np.testing.assert_array_equal(a.refractory_ticks(), b.refractory_ticks())
    np.testing.assert_array_equal(a.voltages().view(np.uint32), b.voltages().view(np.uint32))
    assert a.energy.total_spikes() == b.energy.total_spikes()
    # the in-memory driver's skip rule (zero-current neurons are not stepped) under memoisation
    pool = 300_000
    a, b = make(31, pool), make(31, pool)
    b.debug_set_lif_path(1)
    from neurokmer_b200 import flatten
    for n in (2_000_000, 300_000, 900_000):
        bases, offsets = flatten([random_dna(rng, n, 0.001)])
        a.process_batch(bases, offsets); b.process_batch(bases, offsets)
        np.testing.assert_array_equal(a.voltages().view(np.uint32), b.voltages().view(np.uint32))
        np.testing.assert_array_equal(a.spike_counts(), b.spike_counts())
        np.testing.assert_array_equal(a.refractory_ticks(), b.refractory_ticks())
    assert a.timings()["lif_path"] == 5 and b.timings()["lif_path"] == 1


@pytest.mark.parametrize("cost", [0.0015, 2.5, 0.0009, 0.0, -3.0, float("nan"), 1e30, 123456.789])
def test_energy_tracker_cost_truncation(cost):
    """EnergyTracker (models.rs:159-172, spiking_hash.rs:649-655): `(cost * 1000.0) as u64` truncates toward zero and
    saturates (NaN and negatives -> 0, >= 2^64 -> u64::MAX), the product with the spike count wraps mod 2^64,
    total_energy() = fixed / 1000.0.  Fresh job (fused kernel) and carried state (separate kernels)."""
    from neurokmer_b200 import flatten
    rng = np.random.default_rng(31)
    k, pool = 15, 4096
    bases, offsets = flatten([random_dna(rng, 600_000, 0.001), random_dna(rng, 150)])
    c = make(k, pool, spike_cost=cost)
    o = oracle_counter(k, pool, spike_cost=cost)
    for _ in range(2):
        c.process_batch(bases, offsets); o.process_parallel(bases, offsets)
        assert c.energy.total_spikes() == o.total_spikes > 0
        assert c.energy_used() == o.energy_used()
    from oracle.oracle_py import cost_fixed
    assert o.energy_used() == ((o.total_spikes * cost_fixed(cost)) & (2**64 - 1)) / 1000.0
    c.close()


def test_config2_full_size_against_oracle():
    """BASELINE configs[1] — the configuration the metric is quoted on — at FULL size (113 Mbase, 7 sequences,
    sparse N runs, 1 % lower case, k=31, pool 2 M, canonical, streaming), bit-exact against the oracle: currents,
    spike counts, voltages, refractory ticks, totals, top-20.  Three ways in: the device-resident batch bench.py's
    `value` times, the host batch its `e2e` times (pageable here, pinned below), and the pre-packed form."""
    import bench
    from neurokmer_b200 import PinnedBuffer, pack_bases
    from neurokmer_b200.devmem import copy_h2d
    from oracle.oracle_py import OracleCounter
    from oracle.synth import synth_bases
    k, pool = bench.K, bench.POOL
    n, nseq = bench.NBASES, len(bench.SEQ_LENS)
    offsets = np.concatenate([[0], np.cumsum(bench.SEQ_LENS)]).astype(np.uint64)
    c = make(k, pool)
    db, do = c.stage_reserve(n, nseq)
    c.synth_fill(db, bench.SEED, 0, n, bench.SYNTH_FLAGS); copy_h2d(do, offsets); c.synchronize()
    # the oracle's input: the generator's bytes read back from the device (the numpy twin takes ~25 s for 113 Mbase);
    # three windows of it are checked against the twin here, the generator itself in test_synth_generator_matches_numpy_twin
    from neurokmer_b200.devmem import device_to_numpy
    bases = device_to_numpy(db, n)
    for at in (0, 56_123_457, n - 1_000_000):
        np.testing.assert_array_equal(bases[at:at + 1_000_000], synth_bases(bench.SEED, at, 1_000_000, bench.SYNTH_FLAGS))
    o = OracleCounter(k, 1.0, 0.95, 2, 1.0, pool, True, 1000, threads=os.cpu_count() or 4)
    o.process_streaming([(bases, offsets)])
    assert int(o.currents.sum()) == bench.KMERS
    c.stream_begin(); c.process_staged(n, nseq, 1); c.stream_finish()
    assert c.timings()["kmers"] == bench.KMERS
    assert_topn_equal(c, o, 20); assert_state_equal(c, o)
    c.reset(); c.stream_begin(); c.stream_push(bases, offsets); c.stream_end()
    assert_topn_equal(c, o, 20); assert_state_equal(c, o)
    pin = PinnedBuffer(n); pin.array[:] = bases
    c.reset(); c.stream_begin(); c.stream_push(pin.array, offsets); c.stream_end()
    assert_topn_equal(c, o, 20); assert_state_equal(c, o)
    codes, other, _ = pack_bases(bases)
    c.reset(); c.stream_begin(); c.stream_push_packed(codes, other, offsets); c.stream_end()
    assert_topn_equal(c, o, 20); assert_state_equal(c, o)
    # one input over several members (all on this GPU): sequences cut at 113 M / 3
    from neurokmer_b200 import SpikingKmerCounter
    g = SpikingKmerCounter(k, 1.0, 0.95, 2, 1.0, pool, True, devices=[0, 0, 0])
    g.stream_begin(); g.stream_push(pin.array, offsets); g.stream_end()
    assert_topn_equal(g, o, 20); assert_state_equal(g, o)


@pytest.mark.parametrize("k,pool,canonical", [(31, 2_000_000, True), (21, 99_991, False), (5, 4096, True), (32, 65_537, True)])
def test_long_sequence_mode_without_bitmap_equals_bitmap_mode(coracle, k, pool, canonical, monkeypatch):
    """Long-sequence batches are counted without the invalid-start bitmap (every tile looks its sequence ends up
    in the offsets); NK_BITMAP=1 forces the bitmap kernels.  Both must equal the oracle, including tiles that
    hold more sequence ends than the per-tile list (a run of tiny sequences inside a long-sequence batch),
    sequences shorter than k, ends exactly on tile boundaries, and a batch that ends mid-tile."""
    from neurokmer_b200 import PinnedBuffer, flatten
    rng = np.random.default_rng(k + pool)
    lens = [300_000, 4096 * 3, k - 1, 1, 0, 4096 - 7, 7, 20_000] + [int(x) for x in rng.integers(0, 40, size=60)] + \
           [250_000, k, k + 1, 4096 * 5 + 1, 150_000]
    seqs = [random_dna(rng, n, 0.002, 0.01) for n in lens]
    bases, offsets = flatten(seqs)
    assert bases.size / len(seqs) >= 2048          # mean length: the long-sequence path
    exp, tot = coracle.accumulate(bases, offsets, k, pool, canonical, threads=4)
    pin = PinnedBuffer(bases.size); pin.array[:] = bases
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("NK_BITMAP", env)
        c = make(k, pool, canonical)
        for src in (bases, pin.array):
            c.process_batch(src, offsets)
            np.testing.assert_array_equal(c.currents(), exp)
            assert c.timings()["kmers"] == tot
        c.close()
    monkeypatch.delenv("NK_BITMAP", raising=False)


@pytest.mark.parametrize("k,pool", [(31, 2_000_000), (32, 99_991), (17, 4096), (5, 65_537), (1, 1000)])
def test_runs_of_n_are_counted_once_per_warp_not_once_per_window(coracle, k, pool):
    """A base that is not ACGT is code 0 on both strands (src/models.rs:237,249), so every window inside a run of N is
    the word 0 — and so is every window of a poly-A or poly-T run (canonical) — and the count kernel adds a whole
    lane's (or warp's) windows of such a run with ONE reduction.
    Runs shorter and longer than a lane's 48 bases and a warp's 512, runs that start or end on lane, warp and tile
    boundaries, runs that cross sequence ends (windows there are invalid), IUPAC letters and poly-A next to N
    (poly-A is word 0 too, by the normal path): currents and window totals equal the oracle's."""
    from neurokmer_b200 import PinnedBuffer, flatten, pack_bases
    rng = np.random.default_rng(1000 + k)
    acgt = np.frombuffer(b"ACGT", np.uint8)

    def dna(n):
        return acgt[rng.integers(0, 4, n, dtype=np.uint8)].copy()

    seq = dna(200_000)
    for start, length in [(100, 10), (1000, 47), (2000, 48), (3000, 49), (4096 - 20, 64), (8192, 512), (16384 - 3, 515),
                          (30_000, 5000), (40_960, 4096), (60_000 + 16, 1024), (90_001, 777), (199_000, 1000)]:
        seq[start:start + length] = ord("N")
    seq[120_000:120_100] = ord("A")                      # poly-A next to N: word 0 too
    seq[120_100:120_400] = ord("N")
    seq[140_000:142_000] = ord("T")                      # poly-T: the complement strand's word is 0
    seq[150_000:150_700] = ord("a"); seq[150_700:151_000] = ord("N"); seq[151_000:151_600] = ord("A")
    seq[160_000:160_900] = np.frombuffer(b"TtNn", np.uint8)[rng.integers(0, 4, 900)]      # T and N mixed: rc word 0
    seq[170_000:170_900] = np.frombuffer(b"AaNRY", np.uint8)[rng.integers(0, 5, 900)]     # A and non-ACGT mixed: fwd word 0
    seq[180_000:180_600] = ord("C"); seq[181_000:181_600] = ord("G")                       # not word 0: the normal path
    seq[130_000:130_200] = np.frombuffer(b"RYKMSWnn", np.uint8)[rng.integers(0, 8, 200)]
    tail_n = np.full(3000, ord("N"), np.uint8)           # a sequence that is one run, then one that starts inside N
    mixed = dna(50_000); mixed[:700] = ord("n"); mixed[-600:] = ord("N")
    seqs = [seq, tail_n, mixed, np.full(k, ord("N"), np.uint8), np.full(max(k - 1, 0), ord("N"), np.uint8), dna(70_000)]
    bases, offsets = flatten([s.tobytes() for s in seqs])
    assert bases.size / len(seqs) >= 2048                # the long-sequence (no bitmap) kernel
    for canonical in (True, False):
        exp, tot = coracle.accumulate(bases, offsets, k, pool, canonical, threads=4)
        c = make(k, pool, canonical)
        pin = PinnedBuffer(bases.size); pin.array[:] = bases
        for src in (bases, pin.array):
            c.process_batch(src, offsets)
            np.testing.assert_array_equal(c.currents(), exp)
            assert c.timings()["kmers"] == tot
        codes, other, _ = pack_bases(bases)
        c.process_batch_packed(codes, other, offsets)
        np.testing.assert_array_equal(c.currents(), exp)
        c.close()


@pytest.mark.parametrize("k,pool", [(31, 100_003), (15, 4096)])
def test_runs_of_n_in_exact_tables_and_the_uniques_pass(coracle, k, pool):
    """The count kernel sums the windows inside a run of N per warp (word 0, one neuron).  The exact tables then get
    ONE weighted record for all of them, the uniques pass one copy: counts[0] must still be the number of such windows
    (plus poly-A, which is word 0 by the normal path), kmer_per_neuron and the uniques column unchanged, and
    process_sequence (which ADDS to the table) must count the neuron as touched once per sequence."""
    rng = np.random.default_rng(77 + k)
    a = np.frombuffer(random_dna(rng, 150_000, 0.002), np.uint8).copy()
    a[10_000:16_000] = ord("N"); a[60_000:60_100] = ord("N"); a[99_990:104_096] = ord("n")
    a[120_000:120_200] = ord("A"); a[130_000:131_500] = ord("T"); a[140_000:141_000] = ord("A")
    seqs = [a.tobytes(), b"N" * 5000, random_dna(rng, 40_000), b"N" * (k - 1), b"N" * k]
    keys, counts, uni = _oracle_tables(coracle, seqs, k, pool, True)
    c = make(k, pool, True); c.enable_exact_counts(True)
    c.process_parallel(seqs)
    gk, gc = c.exact_table()
    np.testing.assert_array_equal(gk, keys); np.testing.assert_array_equal(gc, counts)
    assert keys[0] == 0 and c.get_count(0) == int(counts[0]) and int(counts[0]) > 9000
    np.testing.assert_array_equal(c.kmer_per_neuron(), uni)
    top = c.top_abundant_neurons(20)
    assert [t[2] for t in top] == [int(uni[t[0]]) for t in top]
    c.process_parallel([b"N" * 3000, b"n" * 10])           # a call whose every window is inside a run: one record
    gk, gc = c.exact_table()
    assert gk.tolist() == [0] and gc.tolist() == [3000 - k + 1] and c.get_count(0) == 3000 - k + 1
    assert int(c.kmer_per_neuron().sum()) == 1 and int(c.currents().sum()) == 3000 - k + 1
    # the same rows by the second, filtered pass (no table): the neuron of word 0 is among the top rows
    u = make(k, pool, True)
    u.process_parallel(seqs)
    from neurokmer_b200 import flatten
    rows = u.top_abundant_neurons(20)
    idx0 = coracle.neuron_index(0, pool)
    got = u.top_uniques(20, [flatten(seqs)])
    assert [(r[0], r[1]) for r in got] == [(r[0], r[1]) for r in rows]
    assert [r[2] for r in got] == [int(uni[r[0]]) for r in rows]
    assert idx0 in [r[0] for r in rows] or int(counts[0]) < 1000
    # process_sequence: merge mode with a sequence that is ONLY a run of N, then one that starts with a run
    s = make(k, pool, True); s.enable_exact_counts(True)
    total, touched = {}, np.zeros(pool, np.uint32)
    for q in (b"N" * 3000, b"N" * 700 + random_dna(rng, 2000), random_dna(rng, 1500)):
        s.process_sequence(q)
        w = coracle.kmer_words(q, k, True)
        for x in w.tolist():
            total[x] = total.get(x, 0) + 1
        for i in set(coracle.neuron_index(int(x), pool) for x in set(w.tolist())):
            touched[i] += 1
        gk, gc = s.exact_table()
        assert dict(zip(gk.tolist(), gc.tolist())) == total
        np.testing.assert_array_equal(s.kmer_per_neuron(), touched)


def test_pack_staging_of_host_batches_equals_oracle(coracle, monkeypatch):
    """Large host batches (>= 8 MiB) are packed to 2 bits per base + `other` bits by the staging pool's host threads on
    their way to the GPU and counted by the pre-packed kernel (NK_STAGE_PACK=1 forces it whatever the pool size;
    =0 forces plain copies / the in-place read).  Pageable and pinned sources, an odd source address, sequences cut by
    the 32 Mbase chunk boundary and by the 2 MiB pieces, runs of N, IUPAC, lower case, a last piece that is not a
    multiple of 64 bases: same currents, totals and state as the oracle, in all three modes."""
    from neurokmer_b200 import PinnedBuffer, flatten
    rng = np.random.default_rng(41)
    k, pool = 31, 1_000_003
    lens = [20_000_000, 13_554_433 - 20_000_000 % 7, 2_097_151, 150, 0, 30, 5_000_001]
    seqs = [np.frombuffer(random_dna(rng, n, 0.001, 0.01, 0.0005), np.uint8).copy() for n in lens]
    seqs[0][1_000_000:1_004_000] = ord("N"); seqs[0][2_097_100:2_097_300] = ord("n")     # across a 2 MiB piece
    seqs[1][13_000_000:13_000_700] = ord("T")
    bases, offsets = flatten([s.tobytes() for s in seqs])
    assert bases.size > (33 << 20)                       # more than one 32 Mbase chunk
    o = oracle_counter(k, pool)
    o.process_streaming([(bases, offsets)])
    pin = PinnedBuffer(bases.size + 64)
    pin.array[:bases.size] = bases
    odd = np.empty(bases.size + 5, np.uint8)[5:]         # pageable, address not a multiple of 4
    odd[:] = bases
    for mode in ("1", "0"):
        monkeypatch.setenv("NK_STAGE_PACK", mode)
        for src in (bases, odd, pin.array[:bases.size]):
            c = make(k, pool)
            c.stream_begin(); c.stream_push(src, offsets); c.stream_end()
            assert_state_equal(c, o)
            assert c.timings()["kmers"] == sum(max(0, n - k + 1) for n in lens)
            c.close()
    monkeypatch.setenv("NK_STAGE_PACK", "1")
    c = make(k, pool); c.enable_exact_counts(True)       # exact tables fed by the packed kernel instantiation
    c.process_batch(bases[:9_000_000], np.array([0, 4_000_000, 9_000_000], np.uint64))
    keys, counts, uni = _oracle_tables(coracle, [bases[:4_000_000].tobytes(), bases[4_000_000:9_000_000].tobytes()], k, pool, True)
    gk, gc = c.exact_table()
    np.testing.assert_array_equal(gk, keys); np.testing.assert_array_equal(gc, counts)
    c.close()
