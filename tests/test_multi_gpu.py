"""One input sharded over several GPUs inside one process (nk_create_multi) against the CPU oracle.

The reference spreads one input over the cores of one process (rayon fold/reduce, reference
src/spiking_hash.rs:94-154; worker threads, :292-403); a group handle does the same with GPUs and must
give bit-identical results: currents, spike counts, voltages, refractory ticks, totals, top-N.

Most tests put all members on device 0 (the partitioning, the cut sequences, the event-ordered exchange
and the state hand-overs are the same code); tests named *_two_devices need >= 2 GPUs and skip otherwise."""
import os
import subprocess

import numpy as np
import pytest

from conftest import random_dna
from test_parity_gpu import REF, assert_state_equal, assert_topn_equal, oracle_counter

pytestmark = pytest.mark.gpu


def make_group(k, pool, canonical=True, devices=(0, 0), **kw):
    from neurokmer_b200 import SpikingKmerCounter
    p = dict(REF); p.update(kw)
    return SpikingKmerCounter(k, p["threshold"], p["leak"], p["refractory"], p["spike_cost"], pool, canonical,
                              devices=list(devices))


def cut_batch(rng, k, total=3_000_000):
    """long sequences (cut by every shard boundary), reads, sequences shorter than k, empty ones"""
    lens = [total // 2, 0, 1, k - 1, k, total // 3, 150, 20, 150, total // 7, 7, 0, 40000]
    return [random_dna(rng, max(0, n), 0.002, 0.02, 0.002) for n in lens]


@pytest.mark.parametrize("k,pool,canonical,world", [(31, 2_000_000, True, 2), (21, 100_003, True, 3), (15, 65536, True, 4),
                                                    (31, 50_000, False, 3), (32, 999_983, True, 2), (1, 7, True, 2)])
def test_group_matches_oracle_and_carries_state(k, pool, canonical, world):
    """process_parallel on a fresh group (sliced pool), a second and third call on the same counter (state
    carried: gathered to group[0]), streaming with several pushes after a reset."""
    from neurokmer_b200 import flatten
    rng = np.random.default_rng(k * 31 + world)
    g = make_group(k, pool, canonical, devices=[0] * world)
    assert g.group_size() == world
    o = oracle_counter(k, pool, canonical)
    seqs = cut_batch(rng, k)
    bases, offsets = flatten(seqs)
    g.process_batch(bases, offsets)
    o.process_parallel(bases, offsets)
    assert g.timings()["kmers"] == sum(max(0, len(s) - k + 1) for s in seqs)
    assert_topn_equal(g, o, 20)
    assert_state_equal(g, o)          # slices gathered from the members
    assert_topn_equal(g, o, 50)       # more rows than the slices computed: state moves to group[0]
    assert_state_equal(g, o)
    # carried state: currents overwritten, v / r / spike_count / energy carried (spiking_hash.rs:174-176, :642-644)
    seqs2 = cut_batch(rng, k, 1_000_000)
    b2, o2 = flatten(seqs2)
    g.process_batch(b2, o2); o.process_parallel(b2, o2)
    assert_state_equal(g, o); assert_topn_equal(g, o, 20)
    g.stream_begin(); g.stream_push(bases, offsets); g.stream_push(b2, o2); g.stream_end()
    o.process_streaming([(bases, offsets), (b2, o2)])
    assert_state_equal(g, o); assert_topn_equal(g, o, 33)
    g.simulate_spikes_auto(); o._simulate(simd=True)
    assert_state_equal(g, o)
    # a reset group is a fresh counter: streaming semantics on the sliced path
    g.reset()
    o = oracle_counter(k, pool, canonical)
    g.stream_begin()
    for part in (seqs[:3], seqs[3:9], [], seqs[9:]):
        g.stream_push(*flatten(part))
    g.stream_end()
    o.process_streaming([flatten(seqs)])
    assert_topn_equal(g, o, 20)
    assert_state_equal(g, o)
    g.close()


def test_group_pinned_zero_copy_and_packed():
    """pinned batches are read in place by every member (its range of the caller's array); the pre-packed
    form is cut at the same tile-aligned positions"""
    from neurokmer_b200 import PinnedBuffer, flatten, pack_bases
    rng = np.random.default_rng(77)
    k, pool = 31, 300_007
    seqs = cut_batch(rng, k, 2_500_000)
    bases, offsets = flatten(seqs)
    o = oracle_counter(k, pool); o.process_streaming([(bases, offsets)])
    pin = PinnedBuffer(bases.size); pin.array[:] = bases
    g = make_group(k, pool, devices=[0, 0, 0])
    g.stream_begin(); g.stream_push(pin.array, offsets); g.stream_end()
    assert_topn_equal(g, o, 20); assert_state_equal(g, o)
    codes, other, n_other = pack_bases(bases)
    assert n_other > 0
    g.reset()
    g.stream_begin(); g.stream_push_packed(codes, other, offsets); g.stream_end()
    assert_topn_equal(g, o, 20); assert_state_equal(g, o)
    pc = PinnedBuffer(codes.nbytes, np.uint32); pc.array[:] = codes
    po = PinnedBuffer(other.nbytes, np.uint32); po.array[:] = other
    g.reset()
    g.stream_begin(); g.stream_push_packed(pc.array, po.array, offsets); g.stream_end()
    assert_topn_equal(g, o, 20); assert_state_equal(g, o)
    g.close()


def test_group_other_parameters_take_the_leader_path():
    """parameters the per-count table cannot serve (steps = 0 is a no-op; a threshold no count reaches within
    2^20) and non-default spike costs: the group falls back to summing on group[0]"""
    from neurokmer_b200 import flatten
    rng = np.random.default_rng(5)
    k, pool = 21, 40_009
    bases, offsets = flatten(cut_batch(rng, k, 1_200_000))
    for kw in (dict(spike_cost=0.0015), dict(spike_cost=2.5, leak=0.5, refractory=0), dict(threshold=5000.0), dict(threshold=0.25, leak=1.0)):
        g = make_group(k, pool, devices=[0, 0], **kw)
        o = oracle_counter(k, pool, **kw)
        g.process_batch(bases, offsets); o.process_parallel(bases, offsets)
        assert_state_equal(g, o); assert_topn_equal(g, o, 20)
        g.close()
    g = make_group(k, pool, devices=[0, 0]); g.set_steps(0)
    o = oracle_counter(k, pool, steps=0)
    g.process_batch(bases, offsets); o.process_parallel(bases, offsets)
    assert_state_equal(g, o)
    g.close()


def test_group_overflow_guard_spills(coracle):
    """A member that could exceed 2^32-1 window starts in one job folds its u32 counts into the u64 spill array
    its peers read (lowered here with nk_debug_set_fold_limit): sliced and leader paths, and the two-handle
    nk_dist_* flow of the one-process-per-GPU mode."""
    import torch
    from neurokmer_b200 import flatten
    from neurokmer_b200.devmem import copy_d2d
    from neurokmer_b200.shard import shard_batch
    from test_parity_gpu import make
    rng = np.random.default_rng(9)
    k, pool = 31, 200_003
    seqs = [random_dna(rng, n, 0.001) for n in (900_000, 31, 700_000, 5, 400_000)]
    bases, offsets = flatten(seqs)
    o = oracle_counter(k, pool)
    g = make_group(k, pool, devices=[0, 0, 0])
    g.debug_set_fold_limit(150_000)          # every member folds several times per push
    g.stream_begin()
    for _ in range(3):
        g.stream_push(bases, offsets)
    g.stream_end()
    o.process_streaming([(bases, offsets)] * 3)
    assert_topn_equal(g, o, 20); assert_state_equal(g, o)
    g.stream_begin(); g.stream_push(bases, offsets); g.stream_push(bases, offsets); g.stream_end()   # carried state: leader path
    o.process_streaming([(bases, offsets)] * 2)
    assert_state_equal(g, o)
    g.close()
    # one handle per rank (nk_dist_post with an external barrier): only rank 1 spills
    world = 2
    ranks = [make(k, pool) for _ in range(world)]
    raws = [c.dist_export()[1] for c in ranks]
    for r, c in enumerate(ranks):
        c.dist_setup(r, world, raw_ptrs=raws)
    ranks[1].debug_set_fold_limit(100_000)
    ref = oracle_counter(k, pool); ref.process_streaming([(bases, offsets)] * 2)
    for job in range(2):
        for r, c in enumerate(ranks):
            c.reset(); c.stream_begin()
            for _ in range(2):
                c.stream_push(*shard_batch(bases, offsets, k, world, r))
        for c in ranks:
            c.synchronize()
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
            assert c.energy.total_spikes() == ref.total_spikes
            oi, os_ = ref.top_abundant_neurons(20)
            assert [(t[0], t[1]) for t in c.top_abundant_neurons(20)] == [(int(a), int(b)) for a, b in zip(oi, os_)]
            lo, ln = c.dist_slice()
            np.testing.assert_array_equal(c.currents()[lo:lo + ln], ref.currents[lo:lo + ln])
    # ... and a stream that spilled may still end on the non-sharded path
    c = ranks[1]
    c.reset(); c.stream_begin(); c.stream_push(bases, offsets); c.stream_push(bases, offsets); c.stream_end()
    exp, _ = coracle.accumulate(bases, offsets, k, pool, True, threads=4)
    np.testing.assert_array_equal(c.currents(), 2 * exp)


def test_group_file_and_cli(tmp_path):
    """nk_process_file and the CLI's --devices on a group: FASTA with a sequence longer than a batch would be
    cut anyway; result block identical to the single-GPU run (uniques column by the second read on group[0])"""
    from neurokmer_b200 import SpikingKmerCounter
    rng = np.random.default_rng(3)
    seqs = [random_dna(rng, n, 0.001, 0.01) for n in (700_000, 150, 20, 450_000, 31)]
    fa = tmp_path / "in.fa"
    with open(fa, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">s%d\n" % i)
            for j in range(0, len(s), 60):
                f.write(s[j:j + 60] + b"\n")
    k, pool = 21, 10_000
    g = make_group(k, pool, devices=[0, 0, 0])
    one = SpikingKmerCounter(k, REF["threshold"], REF["leak"], REF["refractory"], REF["spike_cost"], pool, True)
    for streaming in (True, False):
        g.reset(); one.reset()
        (g.process_file_streaming if streaming else g.process_file_in_memory)(str(fa))
        (one.process_file_streaming if streaming else one.process_file_in_memory)(str(fa))
        assert g.top_abundant_neurons(20) == one.top_abundant_neurons(20)
        np.testing.assert_array_equal(g.currents(), one.currents())
        np.testing.assert_array_equal(g.spike_counts(), one.spike_counts())
        assert g.energy.total_spikes() == one.energy.total_spikes()
    g.set_file_uniques(20); one.set_file_uniques(20)
    g.reset(); one.reset()
    g.process_file_streaming(str(fa)); one.process_file_streaming(str(fa))
    rows = g.top_abundant_neurons(20)
    assert rows == one.top_abundant_neurons(20) and all(r[2] is not None for r in rows)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neurokmer_b200", "neurokmer")
    args = [exe, "-i", str(fa), "-k", str(k), "--pool-size", str(pool), "--canonical", "--streaming"]
    want = subprocess.check_output(args, text=True)
    got = subprocess.check_output(args + ["--devices", "0,0"], text=True)
    assert got == want and "unique k-mers colliding" in got
    # --exact on a group: the uniques column comes from the table merged across the members
    want_x = subprocess.check_output(args + ["--exact"], text=True)
    got_x = subprocess.check_output(args + ["--exact", "--devices", "0,0,0"], text=True)
    assert got_x == want_x == want
    g.enable_exact_counts(True); one.enable_exact_counts(True)
    g.reset(); one.reset()
    g.process_file_streaming(str(fa)); one.process_file_streaming(str(fa))
    assert g.top_abundant_neurons(20) == one.top_abundant_neurons(20)
    gk, gc = g.exact_table(); ok, oc = one.exact_table()
    np.testing.assert_array_equal(gk, ok); np.testing.assert_array_equal(gc, oc)
    np.testing.assert_array_equal(g.kmer_per_neuron(), one.kmer_per_neuron())


def test_group_rejects_what_it_cannot_do():
    from neurokmer_b200 import NkError
    from neurokmer_b200 import _lib
    g = make_group(31, 1000, devices=[0, 0])
    for call in (lambda: g.process_sequence(b"ACGT" * 20), lambda: g.get_count(5), lambda: g.stage_reserve(100, 1),
                 lambda: g.dist_export(), lambda: g.stream_accumulated()):
        with pytest.raises(NkError) as e:
            call()
        assert e.value.code == _lib.NK_ERR_UNSUPPORTED
    with pytest.raises(NkError) as e:
        g.stream_push(np.zeros(10, np.uint8), np.array([0, 10], np.uint64))
    assert e.value.code == _lib.NK_ERR_STATE
    g.close()
    with pytest.raises(NkError):
        make_group(31, 1000, devices=[0, 99])


def _exact_checks(g, coracle, seqs, k, pool, canonical):
    from test_parity_gpu import _oracle_tables
    keys, counts, uni = _oracle_tables(coracle, seqs, k, pool, canonical)
    gk, gc = g.exact_table()
    np.testing.assert_array_equal(gk, keys); np.testing.assert_array_equal(gc, counts)
    np.testing.assert_array_equal(g.kmer_per_neuron(), uni)
    for i in range(0, len(keys), max(1, len(keys) // 23)):
        assert g.get_count(int(keys[i])) == int(counts[i])
    present = set(keys.tolist())
    for x in range(1, 40):
        if x not in present:
            assert g.get_count(x) is None
    top = g.top_abundant_neurons(20)
    assert [t[2] for t in top] == [int(uni[t[0]]) for t in top]


@pytest.mark.parametrize("k,pool,canonical,world", [(11, 5000, True, 2), (31, 100_003, True, 3), (9, 4096, False, 4), (5, 7, True, 2),
                                                    (21, 1_000_000, True, 2)])
def test_group_exact_tables_merge_by_neuron_slice(coracle, k, pool, canonical, world):
    """nk_enable_exact_counts on a group (src/spiking_hash.rs:157-172, 442-447, 467-473: ONE map over the whole input):
    every member builds the table of its own windows, member d merges the records of its neuron slice from all of
    them.  A word that occurs in several members' shards (repeats, runs of N, poly-A) must come out ONCE with the sum;
    get_count asks the slice's owner; kmer_per_neuron and the uniques column come from the merged table; a second call
    replaces the table; the neuron state is what the plain group computes."""
    rng = np.random.default_rng(k * 7 + world)
    rep = random_dna(rng, 3000)
    seqs = [random_dna(rng, 120_000, 0.01, 0.05), rep * 5, b"", b"ACG", random_dna(rng, 50_000, 0.0, 0.0, 0.0) + b"N" * 3000 + rep,
            b"A" * 5000, rep + random_dna(rng, 70_000, 0.01), b"N" * 2500]
    g = make_group(k, pool, canonical, devices=[0] * world)
    g.enable_exact_counts(True)
    g.process_parallel(seqs)
    _exact_checks(g, coracle, seqs, k, pool, canonical)
    plain = make_group(k, pool, canonical, devices=[0] * world); plain.process_parallel(seqs)
    np.testing.assert_array_equal(plain.currents(), g.currents())
    np.testing.assert_array_equal(plain.spike_counts(), g.spike_counts())
    plain.close()
    seqs2 = [random_dna(rng, 40_000), rep]
    g.process_parallel(seqs2)                      # counts.clear() + refill (:157)
    _exact_checks(g, coracle, seqs2, k, pool, canonical)
    from neurokmer_b200 import flatten
    g.reset()
    g.stream_begin(); g.stream_push(*flatten(seqs[:3])); g.stream_push(*flatten(seqs[3:])); g.stream_end()
    _exact_checks(g, coracle, seqs, k, pool, canonical)
    g.process_parallel([])                         # nothing counted: empty tables
    assert g.exact_table()[0].size == 0 and int(g.kmer_per_neuron().sum()) == 0 and g.get_count(0) is None
    g.enable_exact_counts(False)
    g.process_parallel(seqs2)
    assert [t[2] for t in g.top_abundant_neurons(3)] == [None] * 3
    g.close()


def test_group_exact_tables_two_devices(coracle):
    """the slice merge over real peer copies"""
    _need_two()
    from neurokmer_b200 import device_count
    nd = min(device_count(), 8)
    rng = np.random.default_rng(5)
    k, pool = 31, 2_000_000
    rep = random_dna(rng, 20_000)
    seqs = [random_dna(rng, 900_000, 0.001, 0.01), rep * 3, random_dna(rng, 600_000) + b"N" * 5000 + rep, random_dna(rng, 400_000)]
    for devs in ([0, 1], list(range(nd))):
        g = make_group(k, pool, True, devices=devs)
        g.enable_exact_counts(True)
        g.process_parallel(seqs)
        _exact_checks(g, coracle, seqs, k, pool, True)
        g.close()


def _need_two():
    from neurokmer_b200 import device_count
    if device_count() < 2:
        pytest.skip("needs >= 2 GPUs")


def test_group_two_devices_cut_sequences():
    """the real thing: members on different GPUs, NVLink peer loads inside the slice kernels, result packs
    written into group[0]'s memory; 113-Mbase-shaped input scaled down, sequences cut between the GPUs"""
    _need_two()
    from neurokmer_b200 import PinnedBuffer, device_count, flatten
    nd = min(device_count(), 8)
    rng = np.random.default_rng(11)
    k, pool = 31, 2_000_000
    seqs = [random_dna(rng, n, 0.001, 0.01) for n in (3_000_000, 2_500_000, 2_000_000, 1_500_000, 1_000_000, 800_000, 500_000)]
    bases, offsets = flatten(seqs)
    o = oracle_counter(k, pool); o.process_streaming([(bases, offsets)])
    pin = PinnedBuffer(bases.size); pin.array[:] = bases
    for devs in ([0, 1], list(range(nd))):
        g = make_group(k, pool, devices=devs)
        for src in (bases, pin.array):
            g.reset()
            g.stream_begin(); g.stream_push(src, offsets); g.stream_end()
            assert_topn_equal(g, o, 20)
            assert_state_equal(g, o)
        o2 = oracle_counter(k, pool); o2.process_streaming([(bases, offsets)]); o2.process_parallel(bases, offsets)
        g.process_batch(bases, offsets)     # carried state across devices
        assert_state_equal(g, o2); assert_topn_equal(g, o2, 20)
        g.close()
