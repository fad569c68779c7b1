#!/usr/bin/env python
"""bench.py — canonical k-mers/s of the NeuroKmer counting hot path on B200.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
synthetic 113 Mbase multi-sequence stream (7 sequences {30,25,20,15,10,8,5} Mbase, sparse
N runs, 1 % soft-masked blocks, seed 2), k=31, pool 2,000,000, canonical, streaming
semantics (accumulate -> totals overwrite currents -> every neuron stepped 1000 ticks)
followed by the top-20 read-out.  One "step" = one whole job on a freshly reset counter.

  value : whole-job throughput with the bases already resident in HBM (CUDA events on the
          library's stream, L2 flushed before every step, max over ranks)
  e2e   : the same job through the public host API with HOST buffers: pinned bases ->
          nk_stream_push -> nk_stream_end -> nk_top_n (D2H of the result rows), wall clock around the
          call sequence.  Inside the push the library either lets the count kernel read the pinned batch
          in place across PCIe (`e2e_inplace` forces this), or — on a host with >= 10 staging workers —
          has its host threads pack the bases to 2 bits each on the way into the H2D copies
  e2e_prepacked: e2e with the input handed over in the pre-packed form (2 bits per base +
          `other` bits, nk_stream_push_packed): 3/8 of the bytes on PCIe; packing is untimed
  roofline     : the count kernel (windowing+SipHash+mod+RED) against the three limits the
                 north star names; denominators measured live by nk_calibrate + MEASURED_PEAKS.json
  cpu_baseline : the oracle port (oracle/nk_oracle.c, pthreads, all host cores) on a bounded
                 sample of the same workload — a reported baseline, not the target

  e2e_pageable : e2e with the bases in plain pageable memory (what the reference's `&[Vec<u8>]` is)
  e2e_file     : (N=1) the reference's real entry point — a FASTA file (115 MB, 60-column lines, in the page cache)
                 -> nk_process_file (raw bytes staged by a pool of host threads, records parsed on the device)
                 -> top-20, wall clock
  result.parity: after the timed legs rank 0 runs the CPU oracle over the WHOLE job (all ranks' shards) and
                 compares currents (sha256 of the u64 array), spike counts, total spikes and the top-20
  exact_tables : (N=1) the device-resident job with nk_enable_exact_counts (the reference's `counts` and
                 `kmer_per_neuron` maps), host wall clock per whole job
  config5      : (N>1) BASELINE configs[4] — 1.25 Gbase per GPU of ONE stream of 100 Mbase sequences (10 Gbp at N=8),
                 pool 16 M — device-resident, same timing rules as `value`; checked by "every window counted once"

N > 1 (torchrun, one process per GPU): weak scaling — ONE synthetic stream of N x 113 Mbase (7 sequences of
N x {30,25,20,15,10,8,5} Mbase) cut by window start into N equal ranges: rank r owns starts
[r*113M, (r+1)*113M) and reads k-1 bases past its range, so sequences are cut between GPUs.  Each rank
counts into a full accumulator replica.  `strong_scaling` (N > 1): the 113 Mbase job itself split N ways.  Default (--dist peer): each
rank owns a neuron slice; its LIF/top-N kernel sums that slice of every rank's counts through
NVLink peer memory, and the "finished counting" flags and result packs travel the same way — no
NCCL call per job.  --dist fused: the same kernels with an NCCL barrier + all-gather around them;
--dist allreduce: ONE NCCL all-reduce of the u64 currents, then LIF + top-N on every rank.

`--impl reference` times the reference's CPU algorithm (the oracle port: the Rust crate
cannot be built here, no cargo/rustc) on the host cores, same config/metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, POOL, STEPS_LIF, TOPN, SEED = 31, 2_000_000, 1000, 20, 2
SYNTH_FLAGS = 3
SEQ_LENS = [30_000_000, 25_000_000, 20_000_000, 15_000_000, 10_000_000, 8_000_000, 5_000_000]
NBASES = sum(SEQ_LENS)
KMERS = sum(l - K + 1 for l in SEQ_LENS)
WORKLOAD = "synthetic 113 Mbase FASTA-equivalent (7 seqs, sparse N runs, 1% lowercase), k=31, pool 2M, canonical, streaming"
LIF_REF = dict(threshold=1.0, leak=0.95, refractory=2, spike_cost=1.0)
SIPHASH_OPS = 131  # 32-bit integer ops per k-mer of SipHash-1-3 on one 8-byte block


def make_config(world: int, dist: str):
    """the `config` object of BOTH arms (the driver compares them)"""
    fused = world > 1 and dist in ("fused", "peer")
    return {"workload": WORKLOAD, "k": K, "pool_size": POOL, "lif_steps": STEPS_LIF, "top_n": TOPN,
            "kmers_per_step_per_gpu": KMERS, "l2": "flushed before every step (256 MiB fill)",
            "parallelism": f"dp{world}: ONE stream cut by window start (k-1 overlap), full accumulator replica per GPU, "
                           + ("neuron-sliced LIF/top-N with peer-memory reduce" if fused or world == 1 else "one NCCL all-reduce")
                           + (", peer-memory signalling (no NCCL per job)" if world > 1 and dist == "peer" else "")}


def stream_layout(world: int):
    """the whole job at `world` ranks: 7 sequences of world x SEQ_LENS bases"""
    lens = [l * world for l in SEQ_LENS]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    return lens, offs


def shard_of(offs: np.ndarray, lo: int, hi: int):
    """pieces (offsets relative to lo) the owner of window starts [lo, hi) counts: every sequence clipped to
    [lo, hi + K - 1) — the rule of neurokmer_b200/csrc/nk_multi.cu (shard_pieces)"""
    total = int(offs[-1])
    lim = min(hi + K - 1, total)
    out = [0]
    for s in range(len(offs) - 1):
        a, e = int(offs[s]), int(offs[s + 1])
        if a >= hi:
            break
        p0, p1 = max(a, lo), min(e, lim)
        if p1 > p0:
            out.append(p1 - lo)
    return np.array(out, np.uint64)


def write_fasta(path: str, bases: np.ndarray, offsets: np.ndarray, width: int = 60) -> int:
    """the workload as the file the reference would be given: one record per sequence, 60-column lines"""
    with open(path, "wb") as f:
        for i in range(len(offsets) - 1):
            s = bases[int(offsets[i]):int(offsets[i + 1])]
            f.write(b">seq%d synthetic workload of bench.py\n" % i)
            n = s.size // width * width
            body = np.empty((n // width, width + 1), np.uint8)
            body[:, :width] = s[:n].reshape(-1, width)
            body[:, width] = 10
            f.write(body.tobytes())
            if s.size > n:
                f.write(s[n:].tobytes() + b"\n")
    return os.path.getsize(path)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0:  # nvidia-smi takes a moment to emit its first row
                time.sleep(0.02)
            self.rows.clear()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)  # let the last 100 ms sample land
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 4 + i and r[4 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm (oracle port): used for cpu_baseline and for --impl reference
# ---------------------------------------------------------------------------------------------
def cpu_run(sample_bases: int, threads: int, bases_host: np.ndarray | None = None, exact_sample_bases: int = 0):
    """One bounded pass of the reference's CPU algorithm: accumulate over the first
    `sample_bases` bases of the workload (multi-threaded), LIF over the FULL 2M pool
    (multi-threaded), top-20.  Returns (whole-job k-mers/s extrapolated, detail)."""
    from oracle.oracle_py import COracle
    from oracle.synth import synth_bases
    c = COracle()
    sample_bases = min(sample_bases, NBASES)
    if bases_host is None:
        bases_host = synth_bases(SEED, 0, sample_bases, 3)
    bases = np.ascontiguousarray(bases_host[:sample_bases])
    offs = [0]
    for l in SEQ_LENS:
        if offs[-1] + l >= sample_bases:
            break
        offs.append(offs[-1] + l)
    offs.append(sample_bases)
    offsets = np.array(offs, np.uint64)
    sample_kmers = int(sum(max(0, int(offsets[i + 1] - offsets[i]) - K + 1) for i in range(len(offs) - 1)))
    t0 = time.perf_counter()
    cur, tot = c.accumulate(bases, offsets, K, POOL, True, threads=threads)
    t_acc = time.perf_counter() - t0
    assert tot == sample_kmers
    # scale the sample's currents to the full job's mean so the LIF pass does representative work
    scale = KMERS / max(sample_kmers, 1)
    cur_full = (cur.astype(np.float64) * scale).astype(np.uint64)
    t0 = time.perf_counter()
    fired, v, r, spikes = c.lif(cur_full, STEPS_LIF, 1.0, 0.95, 2, simd_semantics=True, threads=threads)
    t_lif = time.perf_counter() - t0
    t0 = time.perf_counter()
    c.top_n(spikes, TOPN)
    t_top = time.perf_counter() - t0
    t_job = t_acc * scale + t_lif + t_top
    detail = dict(t_acc_sample=t_acc, sample_kmers=sample_kmers, t_lif=t_lif, t_topn=t_top, t_job_extrapolated=t_job)
    if exact_sample_bases:
        # the reference's CPU path ALWAYS also builds its exact side tables (per-task HashMap<u64,u32>,
        # merged `counts`, kmer_per_neuron).  Timed on a smaller sample (hash-map inserts miss the caches) and
        # extrapolated linearly — optimistic for the map, whose miss rate grows with the number of distinct k-mers.
        nb = min(exact_sample_bases, sample_bases)
        offs_x = np.array([0, nb], np.uint64)
        t0 = time.perf_counter()
        _, tot_x, n_distinct, _ = c.accumulate_exact(bases[:nb], offs_x, K, POOL, True, threads=threads)
        t_x = time.perf_counter() - t0
        t_job_x = t_x * (KMERS / max(tot_x, 1)) + t_lif + t_top
        detail.update(exact_value=KMERS / t_job_x, exact_sample_bases=nb, exact_sample_s=t_x, exact_distinct=n_distinct,
                      exact_job_extrapolated=t_job_x)
    return KMERS / t_job, detail


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: the whole --steps K --warmup W run should end within a few minutes.  The LIF + top-N part
    # (~0.2 s on 16 cores) is always run in full; the accumulate part is sized from a conservative rate of
    # 15 Mbase/s per thread so that (K + W) steps take about two minutes (measured on the 16-core GPU box: the
    # fixed part of a step — LIF, top-N, array copies — is ~0.4 s), and never more than the whole workload.
    budget_s = max(0.05, 100.0 / max(1, args.steps + args.warmup) - 0.3)
    sample = int(os.environ.get("NK_REF_SAMPLE_BASES", str(min(NBASES, 8_000_000 * threads, int(15e6 * threads * budget_s)))))
    sample = max(sample, 1_000_000)
    from oracle.synth import synth_bases
    bases = synth_bases(SEED, 0, min(sample, NBASES), 3)
    vals = []
    for i in range(args.warmup + args.steps):
        v, d = cpu_run(sample, threads, bases)
        if i >= args.warmup:
            vals.append((v, d))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1]["t_acc_sample"] + x[1]["t_lif"] + x[1]["t_topn"] for x in vals]) * 1e3)
    samp = (f"accumulate over the first {min(sample, NBASES)} bases ({vals[-1][1]['sample_kmers']} k-mers) on {threads} "
            f"threads, LIF over the full 2M pool x1000 ticks, top-20; job time = acc*{KMERS}/sample + lif + topn")
    print(json.dumps({
        "impl": "reference", "metric": "canonical k-mers/sec (k=31, 2M pool)", "value": v, "unit": "kmers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": make_config(args.gpus, args.dist),
        "cpu_baseline": {"value": v, "unit": "kmers/s", "cores": threads, "kind": "port", "sample": samp},
        "e2e": {"value": v, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is a Rust crate (no cargo/rustc in this image): timed the C restatement oracle/nk_oracle.c",
    }))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class CurrentsView:
    """__cuda_array_interface__ over the library's u64 currents (viewed as int64 for NCCL sum)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def gpu_arm(args):
    import hashlib
    import torch
    import torch.distributed as dist
    from neurokmer_b200 import PinnedBuffer, SpikingKmerCounter
    from neurokmer_b200.devmem import copy_h2d, device_to_numpy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL_DEBUG is the environment's (the driver reads the communicator lines); NCCL's output and everything else
    # this process prints go to stderr — stdout carries the one JSON line
    local %= max(torch.cuda.device_count(), 1)  # launchers that expose one device per rank
    torch.cuda.set_device(local)
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = SpikingKmerCounter(K, LIF_REF["threshold"], LIF_REF["leak"], LIF_REF["refractory"], LIF_REF["spike_cost"],
                           POOL, True, device=local)
    stream = torch.cuda.ExternalStream(c.cuda_stream(), device=torch.device("cuda", local))

    # ---- the job: ONE stream of world x NBASES bases; this rank owns window starts [lo, hi) -----------------
    lens_all, offs_all = stream_layout(world)
    total_bases = int(offs_all[-1])
    total_kmers = int(sum(l - K + 1 for l in lens_all))
    lo, hi = rank * NBASES, (rank + 1) * NBASES
    offsets = shard_of(offs_all, lo, hi)                  # pieces of this rank (sequences cut at lo / hi)
    nseq, nb = len(offsets) - 1, int(offsets[-1])        # nb = NBASES + k-1 halo (except on the last rank)
    my_kmers = int(sum(max(0, int(offsets[i + 1] - offsets[i]) - K + 1) for i in range(nseq)))

    dev_bases, dev_offs = c.stage_reserve(nb, nseq)
    c.synth_fill(dev_bases, SEED, lo, nb, SYNTH_FLAGS)
    copy_h2d(dev_offs, offsets)
    c.synchronize()
    # host copies of the same bytes for the end-to-end legs: pinned, pageable, pre-packed
    pinned = PinnedBuffer(nb if not args.no_e2e else 16)
    pageable = None
    pk_codes = pk_other = None
    pack_ms = None
    if not args.no_e2e:
        pinned.array[:] = device_to_numpy(dev_bases, nb)
        pageable = np.array(pinned.array, copy=True)      # plain malloc memory: what the reference's &[Vec<u8>] is
        from neurokmer_b200 import pack_bases
        pk_codes = PinnedBuffer(4 * ((nb + 15) // 16), np.uint32)
        pk_other = PinnedBuffer(4 * ((nb + 31) // 32), np.uint32)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            _, _, n_other = pack_bases(pinned.array, threads=0, out_codes=pk_codes.array, out_other=pk_other.array)
            best = min(best, time.perf_counter() - t0)
        pack_ms = best * 1e3
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    cur_view = {}
    ar_events = []

    def allreduce_currents():
        ptr = c.stream_accumulated()
        if world > 1:
            if ptr not in cur_view:  # the library's currents buffer never moves: wrap it once
                cur_view[ptr] = torch.as_tensor(CurrentsView(ptr, POOL), device=torch.device("cuda", local))
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                a0.record(stream)
                dist.all_reduce(cur_view[ptr], op=dist.ReduceOp.SUM)
                a1.record(stream)
            ar_events.append((a0, a1))

    trace = []

    # sharded-pool mode (default for N > 1): the reduce-scatter of the counts is fused into the LIF
    # kernel over NVLink peer mappings; NCCL only carries a barrier and a few hundred bytes of results
    fused = world > 1 and args.dist in ("fused", "peer")
    peer = world > 1 and args.dist == "peer"
    if fused:
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        try:
            handle, _ = c.dist_export()
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            c.dist_setup(rank, world, handles=b"".join(handles))
        except Exception as e:  # no peer access between these devices (CUDA IPC refused): every rank must fall back
            print(f"[rank {rank}] sharded-pool setup failed ({e}); falling back to the NCCL all-reduce of the currents", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # also the barrier between setup and the first signal
        if int(ok.item()) == 0:
            fused = peer = False
            args.dist = "allreduce"
        barrier_t = torch.zeros(1, dtype=torch.int32, device="cuda")
        pack_views = {}

    def finish_distributed():
        """everything after counting, for N > 1"""
        if not fused:
            allreduce_currents()
            c.stream_finish()
            return
        if peer:
            # signal + wait + slice LIF/top-N + pack delivery + merge: kernels only, no NCCL, no host barrier
            c.dist_run()
            return
        a0, a1, a2, a3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        with torch.cuda.stream(stream):
            a0.record(stream)
            dist.all_reduce(barrier_t)                     # every rank has finished counting
            a1.record(stream)
        ptr, n64, each = c.dist_post()                     # slice LIF + top-N, counts summed over peers' memory
        if ptr not in pack_views:
            pack_views[ptr] = (torch.as_tensor(CurrentsView(ptr, n64), device=torch.device("cuda", local)),
                               torch.empty(world * n64, dtype=torch.int64, device="cuda"))
        mine, gathered = pack_views[ptr]
        with torch.cuda.stream(stream):
            a2.record(stream)
            dist.all_gather_into_tensor(gathered, mine)    # result packs; also: everyone is done reading my counts
            a3.record(stream)
        c.dist_complete(gathered.data_ptr(), each)
        ar_events.append((a0, a1, a2, a3))

    staged_now = {"nb": nb, "nseq": nseq}

    def job_resident():
        """everything of the job is ENQUEUED here, the result pack's D2H copy included; the caller stops the device
        clock behind it and only then reads the rows on the host (read_result)"""
        t = [time.perf_counter()]
        c.reset(); t.append(time.perf_counter())
        c.stream_begin(); t.append(time.perf_counter())
        c.process_staged(staged_now["nb"], staged_now["nseq"], 1); t.append(time.perf_counter())
        if world > 1:
            finish_distributed(); t.append(time.perf_counter()); t.append(time.perf_counter())
        else:
            t.append(time.perf_counter())
            c.stream_finish(); t.append(time.perf_counter())
        trace.append(np.diff(t) * 1e3)
        return None

    def host_job(push):
        def job():
            c.reset()
            c.stream_begin()
            push()
            if world > 1:
                finish_distributed()
            else:
                c.stream_finish()
            return c.top_abundant_neurons(TOPN)
        return job

    job_e2e = host_job(lambda: c.stream_push(pinned.array, offsets))
    job_e2e_pageable = host_job(lambda: c.stream_push(pageable, offsets))
    job_e2e_packed = host_job(lambda: c.stream_push_packed(pk_codes.array, pk_other.array, offsets))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(job, nsteps, sampler=None):
        per_step, phases = [], []
        if sampler:
            sampler.start()  # before the barrier: nvidia-smi needs ~0.3 s to emit its first row
        barrier()
        for _ in range(nsteps):
            with torch.cuda.stream(stream):
                flush.fill_(1)  # evict the input and the pool from L2 (untimed)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            if job is not job_resident:
                stream.synchronize()  # wall clock must not include the flush
            t0 = time.perf_counter()
            top = job()
            with torch.cuda.stream(stream):
                e1.record(stream)
            if top is None:                       # resident leg: the rows are read after the device clock stopped
                top = c.top_abundant_neurons(TOPN)
            e1.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            per_step.append((e0.elapsed_time(e1), wall))
            phases.append(c.timings())
        clocks = sampler.stop() if sampler else None
        barrier()
        return per_step, phases, top, clocks

    # warm-up (>= 3), then EXACTLY K timed steps of each leg
    W = max(args.warmup, 3)
    timed(job_resident, W)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()  # sampled through ALL timed legs (resident + end-to-end)
    ar_events.clear()
    steps_res, phases, top, _ = timed(job_resident, args.steps)
    if ar_events and len(ar_events[0]) == 4:
        ar_ms = float(np.mean([e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3]) for e in ar_events]))
    else:
        ar_ms = float(np.mean([a.elapsed_time(b) for a, b in ar_events])) if ar_events else 0.0
    # the job's result + the state the oracle is compared with (taken NOW: later legs reuse the counter)
    total_spikes = c.energy.total_spikes()
    slice_lo, slice_len = c.dist_slice() if fused else (0, POOL)
    my_currents = c.currents()[slice_lo:slice_lo + slice_len].copy()
    my_spikes = c.spike_counts()[slice_lo:slice_lo + slice_len].copy()

    legs = {}
    file_path, file_bytes = None, 0
    if not args.no_e2e and world == 1:
        import tempfile
        tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
        file_path = os.path.join(tmpdir, f"nk_bench_{os.getpid()}.fa")
        file_bytes = write_fasta(file_path, pinned.array, offsets)

    def job_e2e_file():
        c.reset()
        c.process_file_streaming(file_path)
        return c.top_abundant_neurons(TOPN)

    if not args.no_e2e:
        # (e2e_inplace runs first: after a leg in which the CPU has read the pinned batch, the GPU's reads of it across
        #  PCIe are slower — 2.76 instead of 2.35 ms — presumably snoops of lines the host's caches still hold)
        for name, job in (("e2e_inplace", job_e2e), ("e2e", job_e2e), ("e2e_pageable", job_e2e_pageable),
                          ("e2e_prepacked", job_e2e_packed), ("e2e_file", job_e2e_file)):
            if name == "e2e_file" and file_path is None:
                continue
            if name == "e2e_inplace":   # the same pinned batch with the host-side packing switched off (A/B)
                os.environ["NK_STAGE_PACK"] = "0"
            timed(job, 2)
            st, ph, tp, _ = timed(job, args.steps)
            os.environ.pop("NK_STAGE_PACK", None)
            assert tp == top, f"{name} and resident legs disagree"
            legs[name] = (st, ph)

    stage_ceiling = None
    if file_path and "e2e_file" in legs:
        stage_ceiling = c.debug_stage_file(file_path)
    if file_path:
        try:
            os.unlink(file_path)
        except OSError:
            pass

    # ---- what the box's host links can do at all: every rank copies its pinned shard to its GPU at the same time ----
    h2d_ceiling = None
    if not args.no_e2e:
        src = torch.from_numpy(pinned.array)      # the library's pinned allocation (cudaMallocHost)
        dst = torch.empty(nb, dtype=torch.uint8, device="cuda")
        best = 1e9
        for rep in range(6):
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            dst.copy_(src, non_blocking=True)
            a1.record()
            a1.synchronize()
            tt = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if rep:
                best = min(best, float(tt[0]))
        h2d_ceiling = {"ms_slowest_rank": best, "gbs_per_rank": nb / (best * 1e-3) / 1e9, "gbs_aggregate": world * nb / (best * 1e-3) / 1e9,
                       "what": "all ranks cudaMemcpyAsync their pinned 113 MB shard to their GPU simultaneously (max over ranks)"}
        del dst

    # ---- exact side tables (N = 1): the same device-resident job with nk_enable_exact_counts -------
    exact_rec = None
    if world == 1 and not args.no_e2e:
        c.enable_exact_counts(True)
        ts = []
        for it in range(7):
            c.reset()
            t0 = time.perf_counter()
            c.stream_begin(); c.process_staged(staged_now["nb"], staged_now["nseq"], 1); c.stream_finish()
            spk = c.energy.total_spikes()   # observes the result: the job and its tables are complete
            ts.append((time.perf_counter() - t0) * 1e3)
        n_keys = c.exact_table_size()
        uni_sum = int(c.kmer_per_neuron().sum(dtype=np.uint64))
        exact_rec = {"ms_per_job": float(np.mean(ts[2:])), "ms_per_job_min": float(np.min(ts[2:])), "jobs": len(ts) - 2,
                     "timing": "host wall clock around one whole job (count + append, bucket partition, per-bucket dedup)",
                     "distinct_kmers": int(n_keys), "windows": int(c.timings()["kmers"]),
                     # every distinct word belongs to exactly one neuron: the per-neuron uniques add up to the table size
                     "check": {"uniques_sum_equals_distinct": bool(uni_sum == int(n_keys)), "total_spikes": int(spk)}}
        c.enable_exact_counts(False)
        c.reset()

    # ---- strong scaling (N > 1): the 113 Mbase job itself, cut N ways ------------------------------------------
    strong = None
    if world > 1:
        _, offs1 = stream_layout(1)
        per = -(-NBASES // world)
        slo, shi = min(NBASES, rank * per), min(NBASES, (rank + 1) * per)
        s_offsets = shard_of(offs1, slo, shi)
        s_nseq, s_nb = len(s_offsets) - 1, int(s_offsets[-1])
        db, do = c.stage_reserve(s_nb, s_nseq)           # the staged buffer only grows: same device memory
        c.synth_fill(db, SEED, slo, s_nb, SYNTH_FLAGS)
        copy_h2d(do, s_offsets)
        c.synchronize()
        staged_now.update(nb=s_nb, nseq=s_nseq)
        timed(job_resident, 2)
        st, ph, s_top, _ = timed(job_resident, args.steps)
        strong = (st, ph, s_top, c.energy.total_spikes())
        staged_now.update(nb=nb, nseq=nseq)
    # ---- configs[4] (N > 1): 1.25 Gbase per GPU of ONE stream of 100 Mbase sequences (10 Gbp at N=8), pool 16 M ----
    config5 = None
    if world > 1 and peer and args.workload == "config2" and not args.no_config5:
        K5, POOL5, SEED5, PER5, SEQ5, STEPS5 = 31, 16_000_000, 5, 1_250_000_000, 100_000_000, 5
        total5 = world * PER5
        offs5 = np.array(list(range(0, total5, SEQ5)) + [total5], np.uint64)
        kmers5 = int(sum(max(0, int(offs5[i + 1] - offs5[i]) - K5 + 1) for i in range(len(offs5) - 1)))
        lo5, hi5 = rank * PER5, (rank + 1) * PER5
        po5 = shard_of(offs5, lo5, hi5)
        c5 = SpikingKmerCounter(K5, LIF_REF["threshold"], LIF_REF["leak"], LIF_REF["refractory"], LIF_REF["spike_cost"],
                                POOL5, True, device=local)
        handle5, _ = c5.dist_export()
        handles5 = [None] * world
        dist.all_gather_object(handles5, handle5)
        c5.dist_setup(rank, world, handles=b"".join(handles5))
        db5, do5 = c5.stage_reserve(int(po5[-1]), len(po5) - 1)
        c5.synth_fill(db5, SEED5, lo5, int(po5[-1]), 1)
        copy_h2d(do5, po5)
        c5.synchronize()
        barrier()   # also separates the setup from the first signal
        stream5 = torch.cuda.ExternalStream(c5.cuda_stream(), device=torch.device("cuda", local))
        ms5, top5 = 0.0, None
        for it in range(2 + STEPS5):
            barrier()
            with torch.cuda.stream(stream5):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream5)
            c5.reset(); c5.stream_begin(); c5.process_staged(int(po5[-1]), len(po5) - 1, 1); c5.dist_run()
            with torch.cuda.stream(stream5):
                e1.record(stream5)
            top5 = c5.top_abundant_neurons(TOPN)
            e1.synchronize()
            if it >= 2:
                ms5 += e0.elapsed_time(e1)
        t5 = torch.tensor([ms5], dtype=torch.float64, device="cuda")
        dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        ph5 = c5.timings()
        config5 = {"workload": f"configs[4]: {total5 / 1e9:.2f} Gbp in 100 Mbase sequences cut over {world} GPUs, k=31, pool 16M, canonical, streaming",
                   "value": kmers5 * STEPS5 / (float(t5[0]) * 1e-3), "unit": "kmers/s", "ms_per_step": float(t5[0]) / STEPS5, "steps": STEPS5,
                   "kmers_per_step": kmers5, "count_ms": ph5["count_ms"], "post_ms": ph5["post_ms"],
                   "exchange_ms": {"wait": ph5["exch_wait_ms"], "reduce": ph5["exch_reduce_ms"], "merge": ph5["merge_ms"]},
                   # no oracle run at this size (10 Gbp): the property the domain offers — every window counted exactly once
                   "check": {"kmers_counted_equal_windows": bool(int(ph5["kmers"]) == kmers5), "total_spikes": c5.energy.total_spikes(),
                             "top1": list(top5[0][:2]) if top5 else None}}
        c5.close()
    clocks = sampler.stop() if sampler else None

    if os.environ.get("NK_TRACE"):
        print(f"[rank {rank}] per-step (event ms, wall ms):", [(round(a, 3), round(b, 3)) for a, b in steps_res[:12]], file=sys.stderr)
        print(f"[rank {rank}] host ms per call (reset, begin, staged, accumulate+allreduce, finish, topn):",
              np.round(np.mean(trace[-args.steps:], axis=0), 3), file=sys.stderr)

    def leg_ms(name):
        return float(sum(x[1] for x in legs[name][0])) if name in legs else float("nan")

    ph_mean = lambda key, src=None: float(np.mean([p[key] for p in (src or phases)]))
    file_ms = leg_ms("e2e_file")
    vals = [float(sum(x[0] for x in steps_res)), leg_ms("e2e"), leg_ms("e2e_pageable"), leg_ms("e2e_prepacked"),
            float(sum(x[0] for x in strong[0])) if strong else float("nan"),
            ph_mean("count_ms"), ph_mean("exch_wait_ms"), ph_mean("exch_reduce_ms"), ph_mean("merge_ms"), ph_mean("post_ms"),
            leg_ms("e2e_inplace")]
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, pg_ms, pk_ms, strong_ms, count_ms_max, wait_ms, reduce_ms, merge_ms, post_ms, ip_ms = (float(x) for x in t)
    job_kmers = total_kmers * args.steps
    value = job_kmers / (dev_ms * 1e-3)

    # ---- parity: the CPU oracle over the WHOLE job (all ranks' shards), outside every timed region ---------
    parity = {"checked": False}
    if not args.no_parity:
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object((slice_lo, my_currents, my_spikes), gathered, dst=0)
        else:
            gathered = [(slice_lo, my_currents, my_spikes)]
        if rank == 0:
            from oracle.oracle_py import COracle
            t0 = time.perf_counter()
            whole = torch.empty(total_bases, dtype=torch.uint8, device="cuda")
            c.synth_fill(whole.data_ptr(), SEED, 0, total_bases, SYNTH_FLAGS)   # the generator is position-addressable
            c.synchronize()
            host = whole.cpu().numpy()
            del whole
            orc = COracle()
            threads = os.cpu_count() or 1
            o_cur, o_tot = orc.accumulate(host, offs_all, K, POOL, True, threads=threads)
            o_fired, _, _, o_spk = orc.lif(o_cur, STEPS_LIF, LIF_REF["threshold"], LIF_REF["leak"], LIF_REF["refractory"],
                                           simd_semantics=True, threads=threads)
            o_idx, o_sp = orc.top_n(o_spk, TOPN)
            g_cur = np.zeros(POOL, np.uint64)
            g_spk = np.zeros(POOL, np.uint64)
            for lo_r, cur_r, spk_r in gathered:
                g_cur[lo_r:lo_r + cur_r.size] = cur_r
                g_spk[lo_r:lo_r + spk_r.size] = spk_r
            parity = {
                "checked": True,
                "oracle": "oracle/nk_oracle.c over the whole job (all ranks' shards as ONE stream), LIF, top-%d" % TOPN,
                "currents_equal": bool(np.array_equal(g_cur, o_cur)),
                "currents_sha256": hashlib.sha256(g_cur.tobytes()).hexdigest(),
                "oracle_currents_sha256": hashlib.sha256(o_cur.tobytes()).hexdigest(),
                "spike_counts_equal": bool(np.array_equal(g_spk, o_spk)),
                "total_spikes_equal": bool(int(o_fired) == int(total_spikes)),
                "kmers_equal": bool(int(o_tot) == total_kmers == int(g_cur.sum())),
                "topn_equal": bool([(int(a), int(b)) for a, b in zip(o_idx, o_sp)] == [(r[0], r[1]) for r in top]),
                "oracle_seconds": round(time.perf_counter() - t0, 2), "oracle_threads": threads,
            }
            if strong:  # the strong-scaling leg runs the N=1 job: its result is checked against the same oracle at N=1 size
                _, offs1 = stream_layout(1)
                s_cur, _ = orc.accumulate(host[:NBASES], offs1, K, POOL, True, threads=threads)  # same generator, same positions
                s_fired, _, _, s_spk = orc.lif(s_cur, STEPS_LIF, LIF_REF["threshold"], LIF_REF["leak"], LIF_REF["refractory"],
                                               simd_semantics=True, threads=threads)
                s_idx, s_sp = orc.top_n(s_spk, TOPN)
                parity["strong_total_spikes_equal"] = bool(int(s_fired) == int(strong[3]))
                parity["strong_topn_equal"] = bool([(int(a), int(b)) for a, b in zip(s_idx, s_sp)] == [(r[0], r[1]) for r in strong[2]])

    if rank == 0:
        pk, pk_kind = peaks()
        count_ms = ph_mean("count_ms")
        kps_kernel = my_kmers / (count_ms * 1e-3)
        # the three limits of the north star, denominators measured on this device
        c.reset()
        alu_ops = c.calibrate(0)
        sip_ops = c.calibrate(1)
        red_ps = c.calibrate(2)
        c.reset()
        hbm_bound = pk["hbm_gbs"] * 1e9 / (nb / my_kmers)       # 1 B of ASCII per k-mer
        int_bound = sip_ops / SIPHASH_OPS
        red_bound = red_ps
        binding = min((hbm_bound, "hbm"), (int_bound, "int32"), (red_bound, "l2_atomic"))
        traffic, traffic_src = ncu_traffic()
        roofline = {
            "bound": binding[1], "kernel": "count_kernel (windowing+SipHash-1-3+mod+RED.ADD)",
            "achieved": kps_kernel * SIPHASH_OPS / 1e9 if binding[1] == "int32" else kps_kernel / 1e9,
            "peak": sip_ops / 1e9 if binding[1] == "int32" else binding[0] / 1e9,
            "unit": "Gop/s (32-bit integer, SipHash mix)" if binding[1] == "int32" else "G/s",
            "frac": kps_kernel / binding[0],
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload from the committed
            # `ncu --set full` capture (profiles/, see tools/ncu_summary.py); null when no capture matches this kernel
            "traffic": traffic, "traffic_unit": "bytes per launch (ncu)", "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": nb,
            "bound_note": "north star: binding roofline = slowest of {HBM input bytes, SipHash int ops, L2 pool updates}; "
                          "the integer ALU pipe binds; the HBM view is under 'hbm'",
            "kernel_ms": count_ms, "kernel_kmers_per_s": kps_kernel,
            "limits_kmers_per_s": {"hbm": hbm_bound, "int32": int_bound, "l2_atomic": red_bound},
            "hbm": {"achieved": nb / (count_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": nb / (count_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_source": pk_kind + " (MEASURED_PEAKS.json)"},
            "peaks_measured_live": {"alu_lop3_shf_gops": alu_ops / 1e9, "sipround_mix_gops": sip_ops / 1e9,
                                    "red_add_u32_hashed_addresses_gps": red_ps / 1e9},
            "algorithmic_per_kmer": {"hbm_bytes": nb / my_kmers, "int32_ops": SIPHASH_OPS, "l2_reductions": 1},
        }
        ph = {k: ph_mean(k) for k in ("mark_ms", "count_ms", "fold_ms", "lif_ms", "topn_ms", "post_ms")}
        launches = int(phases[-1]["launches"] + phases[-1]["topn_launches"])
        # CPU baseline on a bounded sample (rank 0, N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu and not args.no_e2e:
            threads = os.cpu_count() or 1
            sample = min(NBASES, 8_000_000 * threads)
            v, d = cpu_run(sample, threads, pinned.array, exact_sample_bases=min(sample, 2_000_000 * threads))
            cpu = {"value": v, "unit": "kmers/s", "cores": threads, "kind": "port",
                   "sample": (f"accumulate over the first {sample} bases ({d['sample_kmers']} k-mers, {d['t_acc_sample']:.2f} s) + LIF over "
                              f"the full 2M pool ({d['t_lif']:.2f} s) + top-20; job time extrapolated = {d['t_job_extrapolated']:.2f} s"),
                   # the same with the exact side tables the reference's CPU path always builds (hot path + f1)
                   "with_exact_map": {"value": d["exact_value"], "unit": "kmers/s",
                                      "sample": (f"accumulate + per-thread hash maps + merged counts + kmer_per_neuron over the first "
                                                 f"{d['exact_sample_bases']} bases ({d['exact_distinct']} distinct k-mers, "
                                                 f"{d['exact_sample_s']:.2f} s), same LIF + top-20; job time extrapolated = "
                                                 f"{d['exact_job_extrapolated']:.2f} s")}}

        def host_path(name):
            if name not in legs:
                return None
            moved = int(legs[name][1][-1]["h2d_bytes"])
            return ("packed to 2 bits per base by the library's host threads on the way" if moved < nb // 2
                    else "ASCII bytes cross the link (in-place read of pinned memory, or copies of pageable memory)")

        def leg_record(name, ms, extra=None):
            if name not in legs:
                return None
            ph_l = legs[name][1]
            walls = sorted(x[1] for x in legs[name][0])   # this rank's wall clock per step
            rec = {"value": job_kmers / (ms * 1e-3), "unit": "kmers/s", "ms_per_step": ms / args.steps,
                   "ms_per_step_median": walls[len(walls) // 2], "ms_per_step_min": walls[0], "ms_per_step_max": walls[-1],
                   "h2d_bytes_per_step": int(ph_l[-1]["h2d_bytes"]),
                   "d2h_bytes_per_step": int(ph_l[-1]["d2h_bytes"] + 24 + 16 * TOPN)}
            if rec["h2d_bytes_per_step"]:
                rec["pcie_gbs_per_rank"] = rec["h2d_bytes_per_step"] / (ms / args.steps * 1e-3) / 1e9
            rec.update(extra or {})
            return rec

        if peer:
            collective_ms = wait_ms + reduce_ms + merge_ms
        else:
            collective_ms = ar_ms
        line = {
            "metric": f"canonical k-mers/sec (k={K}, {POOL // 1_000_000}M pool)", "value": value, "unit": "kmers/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": make_config(world, args.dist),
            # host path of the pinned batch: with >= 10 staging workers (a host with cores to spare) the library's threads pack
            # the bases to 2 bits each on their way into the H2D copies (h2d_bytes_per_step = 3/8 of the input); otherwise
            # the count kernel reads the pinned batch in place across PCIe.  `e2e_inplace` is the second path forced.
            "e2e": leg_record("e2e", e2e_ms, {"host_path": host_path("e2e")}),
            "e2e_inplace": leg_record("e2e_inplace", ip_ms, {"host_path": host_path("e2e_inplace")}),
            # the same job from PAGEABLE host memory (plain malloc: the reference's &[Vec<u8>]): through the staging pool
            # (packed on the way where the host has the cores, plain copies otherwise)
            "e2e_pageable": leg_record("e2e_pageable", pg_ms, {"host_path": host_path("e2e_pageable")}),
            # same job, same API, input handed over in the library's pre-packed form (2-bit codes + `other`
            # bits in pinned host memory: nk_stream_push_packed).  The packing itself (nk_pack_bases, host
            # SIMD, all cores) is NOT in this timed region — its cost is reported beside it.
            "e2e_prepacked": leg_record("e2e_prepacked", pk_ms, {"host_pack_ms_untimed": pack_ms, "host_pack_threads": os.cpu_count()}),
            # the reference's own entry point: a FASTA file (page cache / tmpfs) -> nk_process_file -> top-20, wall clock
            "e2e_file": leg_record("e2e_file", file_ms, {"file_bytes": file_bytes, "file": "FASTA, 7 records, 60-column lines",
                                                         "ratio_to_e2e": (file_ms / e2e_ms) if e2e_ms == e2e_ms and e2e_ms > 0 else None,
                                                         # what this host can do at all: the staging pool reading the file
                                                         # into pinned memory with no device work / with the H2D copies
                                                         "host_read_ceiling_ms": stage_ceiling[0] if stage_ceiling else None,
                                                         "host_read_plus_h2d_ceiling_ms": stage_ceiling[1] if stage_ceiling else None}),
            "gpu_launches": launches * args.steps,
            "phases_ms": ph, "collective_ms": collective_ms,
            # the exchange on its own (north star): measured by the kernels with %globaltimer, max over ranks.
            #   wait_ms    waiting for the peers' "finished counting" flags and result packs (rank skew)
            #   reduce_ms  the fused phase that reads this rank's neuron slice of every rank's counts over NVLink
            #              (reduce-scatter) + LIF look-up + top-N histogram; at N=1 the same phase on local memory
            #   merge_ms   merging the ranks' result packs
            "exchange": {"wait_ms": wait_ms, "reduce_ms": reduce_ms, "merge_ms": merge_ms, "post_kernel_ms": post_ms,
                         "h2d_ceiling": h2d_ceiling,
                         "count_ms_slowest_rank": count_ms_max,
                         "bytes_read_from_peers_per_rank": int(phases[-1]["exch_bytes"])},
            "collective": ("none" if world == 1 else (
                "none per job: counting-finished flags, count reduce-scatter and result-pack exchange are NVLink peer-memory "
                "loads/stores inside the kernels; collective_ms = flag/pack waits + the fused reduce phase + pack merge" if peer else (
                    "barrier + all-gather of result packs (NCCL); count reduce-scatter fused into the LIF "
                    "kernel over NVLink peer memory" if fused else "all-reduce of the u64 currents (NCCL)"))),
            "lif_path": int(phases[-1]["lif_path"]),
            "strong_scaling": None if not strong else {
                "value": KMERS * args.steps / (strong_ms * 1e-3), "unit": "kmers/s", "ms_per_step": strong_ms / args.steps,
                "kmers_per_step": KMERS, "note": "the N=1 job (113 Mbase, 7 sequences) cut by window start into N ranges"},
            "config5": config5, "exact_tables": exact_rec,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "result": {"total_spikes": total_spikes, "top1": list(top[0][:2]) if top else None, "parity": parity},
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic():
    """dram bytes (read + write) per launch of the count kernel from the committed ncu capture — only if the
    capture was taken from THIS kernel source (sha256 of nk_count.cu + nk_device.cuh recorded beside it)"""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "count_kernel_traffic.json")) as f:
            rec = json.load(f)
        h = hashlib.sha256()
        for name in ("nk_count.cu", "nk_device.cuh"):
            with open(os.path.join(ROOT, "neurokmer_b200", "csrc", name), "rb") as f:
                h.update(f.read())
        if rec.get("source_sha256") != h.hexdigest():
            return None, "profiles/count_kernel_traffic.json is from another build of the kernel"
        return float(rec["dram_bytes_per_launch"]), rec.get("capture", "profiles/count_kernel_traffic.json")
    except Exception:
        return None, "no ncu capture committed for this kernel build"


def select_workload(name: str):
    """configs[1] is the bench workload; configs[4] (10 Gbp, pool 16 M, sharded over the GPUs) can be
    timed with --workload config5 (per-GPU shard = 1.25 Gbp, i.e. the full 10 Gbp at --gpus 8)."""
    global K, POOL, SEQ_LENS, NBASES, KMERS, WORKLOAD, SEED, SYNTH_FLAGS
    if name == "config5":
        K, POOL, SEED, SYNTH_FLAGS = 31, 16_000_000, 5, 1
        SEQ_LENS = [100_000_000] * 12 + [50_000_000]
        WORKLOAD = ("synthetic multi-FASTA-equivalent, 1.25 Gbase per GPU (10 Gbp at 8 GPUs), sparse N runs, k=31, "
                    "pool 16M, canonical, streaming")
    NBASES = sum(SEQ_LENS)
    KMERS = sum(l - K + 1 for l in SEQ_LENS)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"])
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (large workloads)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the job's result (large workloads)")
    ap.add_argument("--no-config5", action="store_true", help="N > 1: skip the configs[4] sub-record (1.25 Gbase per GPU, pool 16M)")
    ap.add_argument("--dist", default="peer", choices=["peer", "fused", "allreduce"],
                    help="N > 1: peer = sharded pool, every exchange through NVLink peer memory inside the kernels (default); "
                         "fused = same kernels with an NCCL barrier + all-gather around them; allreduce = NCCL all-reduce of the currents")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
