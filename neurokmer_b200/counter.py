"""Python mirror of the reference's counter API on top of the C ABI.

`SpikingKmerCounter` has the reference type's names, argument order and error
behaviour (reference src/spiking_hash.rs:39-715); `PySpikingCounter` has the
surface of src/python.rs:7-53.  All arithmetic happens in libneurokmer.so (CUDA,
sm_100a): this module only flattens inputs and formats outputs.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import NkConfig, NkTimings, NkTopEntry, check

BytesLike = Union[bytes, bytearray, memoryview, np.ndarray]


def flatten(seqs: Iterable[BytesLike]) -> Tuple[np.ndarray, np.ndarray]:
    """`&[Vec<u8>]` -> (concatenated uint8 bases, uint64 offsets[nseq+1])."""
    arrs = [np.frombuffer(s, dtype=np.uint8) if not isinstance(s, np.ndarray) else np.ascontiguousarray(s, np.uint8)
            for s in seqs]
    offsets = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offsets[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
        bases = np.concatenate(arrs) if offsets[-1] else np.zeros(0, np.uint8)
    else:
        bases = np.zeros(0, np.uint8)
    return bases, offsets


def _ptr(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None or a.size == 0 else a.ctypes.data


def pack_bases(bases: BytesLike, threads: int = 0, body: int = 0, out_codes: Optional[np.ndarray] = None,
               out_other: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray, int]:
    """ASCII bases -> the pre-packed "nk2" form (include/neurokmer.h): (codes u32, other u32, n_other).
    Done by the library's SIMD host packer (nk_pack_bases); `body` 1..3 names one of its bodies."""
    a = np.frombuffer(bases, np.uint8) if not isinstance(bases, np.ndarray) else np.ascontiguousarray(bases, np.uint8)
    L = _lib.lib()
    nc, no = int(L.nk_packed_code_words(a.size)), int(L.nk_packed_other_words(a.size))
    codes = out_codes if out_codes is not None else np.zeros(max(nc, 1), np.uint32)
    other = out_other if out_other is not None else np.zeros(max(no, 1), np.uint32)
    assert codes.dtype == np.uint32 and other.dtype == np.uint32 and codes.size >= nc and other.size >= no
    n_other = C.c_uint64()
    if body:
        check(L.nk_debug_pack_body(_ptr(a), a.size, codes.ctypes.data, other.ctypes.data, body, C.byref(n_other)))
    else:
        check(L.nk_pack_bases(_ptr(a), a.size, codes.ctypes.data, other.ctypes.data, threads, C.byref(n_other)))
    return codes[:max(nc, 1)], other[:max(no, 1)], int(n_other.value)


class PinnedBuffer:
    """Page-locked host memory from the library (nk_host_alloc) viewed as a numpy array."""

    def __init__(self, nbytes: int, dtype=np.uint8):
        self._p = C.c_void_p()
        check(_lib.lib().nk_host_alloc(C.byref(self._p), nbytes))
        self.nbytes = nbytes
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=nbytes).view(dtype)

    def free(self) -> None:
        if self._p is not None and self._p.value:
            self.array = None
            _lib.lib().nk_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _Energy:
    """`counter.energy.total_spikes()` / `.total_energy()` — reference src/models.rs:145-173."""

    def __init__(self, owner: "SpikingKmerCounter"):
        self._o = owner

    def total_spikes(self) -> int:
        out = C.c_uint64()
        check(self._o._L.nk_total_spikes(self._o._h, C.byref(out)))
        return out.value

    def total_energy(self) -> float:
        return self._o.energy_used()


class SpikingKmerCounter:
    """reference src/spiking_hash.rs:16-37 (`new` at :40-77)."""

    def __init__(self, k: int, threshold: float, leak: float, refractory: int, spike_cost: float,
                 pool_size: int, use_canonical: bool, device: int = 0, devices: Optional[Sequence[int]] = None):
        """`devices`: shard every input over these GPUs inside this process (nk_create_multi); the object is
        used exactly like a single-GPU counter."""
        self._L = _lib.lib()
        cfg = NkConfig()
        check(self._L.nk_config_default(C.byref(cfg)))
        cfg.k, cfg.threshold, cfg.leak = k, threshold, leak
        cfg.refractory, cfg.spike_cost, cfg.pool_size = refractory, spike_cost, pool_size
        cfg.use_canonical, cfg.device = int(bool(use_canonical)), device
        self._h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            check(self._L.nk_create_multi(C.byref(cfg), arr, len(devices), C.byref(self._h)))
        else:
            check(self._L.nk_create(C.byref(cfg), C.byref(self._h)))
        self.k, self.pool_size, self.use_canonical = k, pool_size, bool(use_canonical)
        self.energy = _Energy(self)

    def group_size(self) -> int:
        n = C.c_int32()
        check(self._L.nk_group_size(self._h, C.byref(n)))
        return n.value

    # lifecycle ------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.nk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        check(self._L.nk_reset(self._h))

    # reference entry points -------------------------------------------------
    def process_parallel(self, seqs: Sequence[BytesLike]) -> None:
        """src/spiking_hash.rs:84-201"""
        bases, offsets = flatten(seqs)
        self.process_batch(bases, offsets)

    def process_batch(self, bases: np.ndarray, offsets: np.ndarray) -> None:
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        check(self._L.nk_process_batch(self._h, _ptr(bases), offsets.ctypes.data, offsets.size - 1))

    def process_sequence(self, seq: BytesLike) -> None:
        """src/spiking_hash.rs:203-273"""
        a = np.frombuffer(seq, np.uint8) if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq, np.uint8)
        check(self._L.nk_process_sequence(self._h, _ptr(a), a.size))

    def process_file_streaming(self, path: str) -> None:
        """src/spiking_hash.rs:277-486; raises NkError(NK_ERR_IO) where the reference returns Err."""
        check(self._L.nk_process_file(self._h, str(path).encode(), 1))

    def process_file_in_memory(self, path: str) -> None:
        """src/main.rs:44-45: stream_sequences().collect() + process_parallel."""
        check(self._L.nk_process_file(self._h, str(path).encode(), 0))

    # streaming by batches (process_file_streaming minus parsing)
    def stream_begin(self) -> None:
        check(self._L.nk_stream_begin(self._h))

    def stream_push(self, bases: np.ndarray, offsets: np.ndarray) -> None:
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        check(self._L.nk_stream_push(self._h, _ptr(bases), offsets.ctypes.data, offsets.size - 1))

    def stream_end(self) -> None:
        check(self._L.nk_stream_end(self._h))

    # pre-packed input (2 bits per base; include/neurokmer.h "nk2" layout) -------------------
    def process_batch_packed(self, codes: np.ndarray, other: Optional[np.ndarray], offsets: np.ndarray) -> None:
        codes = np.ascontiguousarray(codes, np.uint32)
        other = None if other is None else np.ascontiguousarray(other, np.uint32)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        check(self._L.nk_process_batch_packed(self._h, codes.ctypes.data, None if other is None else other.ctypes.data,
                                              offsets.ctypes.data, offsets.size - 1))

    def stream_push_packed(self, codes: np.ndarray, other: Optional[np.ndarray], offsets: np.ndarray) -> None:
        codes = np.ascontiguousarray(codes, np.uint32)
        other = None if other is None else np.ascontiguousarray(other, np.uint32)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        check(self._L.nk_stream_push_packed(self._h, codes.ctypes.data, None if other is None else other.ctypes.data,
                                            offsets.ctypes.data, offsets.size - 1))

    def simulate_spikes_auto(self) -> None:
        """src/spiking_hash.rs:697-714"""
        check(self._L.nk_simulate(self._h))

    def top_abundant_neurons(self, top_n: int) -> List[Tuple[int, int, Optional[int]]]:
        """src/spiking_hash.rs:661-673 -> [(idx, spike_count, uniques)]; `uniques` is None
        until the exact side table exists — never a made-up number."""
        n = min(int(top_n), self.pool_size)
        if n <= 0:
            return []
        out = (NkTopEntry * n)()
        got = C.c_uint64()
        check(self._L.nk_top_n(self._h, n, out, C.byref(got)))
        return [(int(e.idx), int(e.spikes), None if e.uniques == _lib.NK_UNIQUES_NOT_COMPUTED else int(e.uniques))
                for e in out[: got.value]]

    # `uniques` of the top rows by a second pass over the input (no O(windows) table) ----------------
    def uniques_begin(self, top_n: int) -> None:
        check(self._L.nk_uniques_begin(self._h, top_n))

    def uniques_push(self, bases: np.ndarray, offsets: np.ndarray) -> None:
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        check(self._L.nk_uniques_push(self._h, _ptr(bases), offsets.ctypes.data, offsets.size - 1))

    def uniques_push_packed(self, codes: np.ndarray, other: Optional[np.ndarray], offsets: np.ndarray) -> None:
        codes = np.ascontiguousarray(codes, np.uint32)
        other = None if other is None else np.ascontiguousarray(other, np.uint32)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        check(self._L.nk_uniques_push_packed(self._h, codes.ctypes.data, None if other is None else other.ctypes.data,
                                             offsets.ctypes.data, offsets.size - 1))

    def uniques_end(self) -> None:
        check(self._L.nk_uniques_end(self._h))

    def top_uniques(self, top_n: int, batches) -> List[Tuple[int, int, Optional[int]]]:
        """top_abundant_neurons(top_n) with the `uniques` column filled by re-supplying the (bases, offsets)
        batches that were counted."""
        self.uniques_begin(top_n)
        for bases, offsets in batches:
            self.uniques_push(bases, offsets)
        self.uniques_end()
        return self.top_abundant_neurons(top_n)

    def set_file_uniques(self, top_n: int) -> None:
        """process_file_* then fills `uniques` of the top_n rows by reading the file a second time"""
        check(self._L.nk_set_file_uniques(self._h, top_n))

    def enable_exact_counts(self, on: bool = True) -> None:
        """Build the reference's `counts` / `kmer_per_neuron` side tables (off by default)."""
        check(self._L.nk_enable_exact_counts(self._h, int(on)))

    def exact_table(self) -> Tuple[np.ndarray, np.ndarray]:
        """(k-mer words, counts) — the reference's `counts` map.  The library hands the table out grouped by neuron
        range (a hash map has no order); it is put in ascending word order here for the caller's convenience."""
        n = C.c_uint64()
        check(self._L.nk_exact_table_size(self._h, C.byref(n)))
        keys, counts = np.zeros(max(n.value, 1), np.uint64), np.zeros(max(n.value, 1), np.uint32)
        check(self._L.nk_copy_exact_table(self._h, keys.ctypes.data, counts.ctypes.data))
        keys, counts = keys[: n.value], counts[: n.value]
        order = np.argsort(keys, kind="stable")
        return keys[order], counts[order]

    def exact_table_size(self) -> int:
        n = C.c_uint64()
        check(self._L.nk_exact_table_size(self._h, C.byref(n)))
        return int(n.value)

    def kmer_per_neuron(self) -> np.ndarray:
        return self._copy(self._L.nk_copy_uniques, np.uint32)

    def get_count(self, kmer: int) -> Optional[int]:
        """src/spiking_hash.rs:675-678"""
        cnt, found = C.c_uint32(), C.c_int32()
        check(self._L.nk_get_count(self._h, kmer, C.byref(cnt), C.byref(found)))
        return int(cnt.value) if found.value else None

    def energy_used(self) -> float:
        out = C.c_double()
        check(self._L.nk_energy_used(self._h, C.byref(out)))
        return out.value

    def set_steps(self, steps: int) -> None:
        check(self._L.nk_set_steps(self._h, steps))

    def get_steps(self) -> int:
        out = C.c_uint64()
        check(self._L.nk_get_steps(self._h, C.byref(out)))
        return out.value

    # parity taps ------------------------------------------------------------
    def debug_kmers(self, seq: BytesLike):
        a = np.frombuffer(seq, np.uint8) if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq, np.uint8)
        n = max(0, a.size - self.k + 1)
        fwd, rc, words, idx = (np.zeros(max(n, 1), np.uint64) for _ in range(4))
        got = C.c_uint64()
        check(self._L.nk_debug_kmers(self._h, _ptr(a), a.size, fwd.ctypes.data, rc.ctypes.data, words.ctypes.data,
                                     idx.ctypes.data, C.byref(got)))
        m = got.value
        return fwd[:m], rc[:m], words[:m], idx[:m]

    def debug_kmers_packed(self, codes: np.ndarray, other: Optional[np.ndarray], length: int):
        codes = np.ascontiguousarray(codes, np.uint32)
        other = None if other is None else np.ascontiguousarray(other, np.uint32)
        n = max(0, length - self.k + 1)
        fwd, rc, words, idx = (np.zeros(max(n, 1), np.uint64) for _ in range(4))
        got = C.c_uint64()
        check(self._L.nk_debug_kmers_packed(self._h, codes.ctypes.data, None if other is None else other.ctypes.data,
                                            length, fwd.ctypes.data, rc.ctypes.data, words.ctypes.data,
                                            idx.ctypes.data, C.byref(got)))
        m = got.value
        return fwd[:m], rc[:m], words[:m], idx[:m]

    def debug_hash(self, words: np.ndarray):
        w = np.ascontiguousarray(words, np.uint64)
        hs, ix = np.zeros(max(w.size, 1), np.uint64), np.zeros(max(w.size, 1), np.uint64)
        check(self._L.nk_debug_hash(self._h, _ptr(w), w.size, hs.ctypes.data, ix.ctypes.data))
        return hs[: w.size], ix[: w.size]

    def _copy(self, fn, dtype) -> np.ndarray:
        out = np.zeros(self.pool_size, dtype)
        check(fn(self._h, out.ctypes.data))
        return out

    def currents(self) -> np.ndarray:
        return self._copy(self._L.nk_copy_currents, np.uint64)

    def spike_counts(self) -> np.ndarray:
        return self._copy(self._L.nk_copy_spike_counts, np.uint64)

    def voltages(self) -> np.ndarray:
        return self._copy(self._L.nk_copy_voltages, np.float32)

    def refractory_ticks(self) -> np.ndarray:
        return self._copy(self._L.nk_copy_refractory, np.uint32)

    def debug_set_lif_path(self, mode: int) -> None:
        check(self._L.nk_debug_set_lif_path(self._h, mode))

    def debug_parse_file(self, path: str) -> Tuple[int, int, int]:
        """(records, bases, FNV-1a digest) of the device-side record parser on a plain FASTA / FASTQ file"""
        nr, nb, hs = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self._L.nk_debug_parse_file(self._h, str(path).encode(), C.byref(nr), C.byref(nb), C.byref(hs)))
        return nr.value, nb.value, hs.value

    def debug_stage_file(self, path: str) -> Tuple[float, float]:
        """(host-only ms, ms with the H2D copies) of the staging pool on a file: the ceiling of the file path"""
        a, b = C.c_double(), C.c_double()
        check(self._L.nk_debug_stage_file(self._h, str(path).encode(), C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_set_fold_limit(self, limit: int) -> None:
        check(self._L.nk_debug_set_fold_limit(self._h, limit))

    def calibrate(self, which: int) -> float:
        out = C.c_double()
        check(self._L.nk_calibrate(self._h, which, C.byref(out)))
        return out.value

    def timings(self) -> dict:
        t = NkTimings()
        check(self._L.nk_last_timings(self._h, C.byref(t)))
        return t.as_dict()

    # device-resident input ----------------------------------------------------
    def stage_reserve(self, nbytes: int, nseq: int) -> Tuple[int, int]:
        b, o = C.c_void_p(), C.c_void_p()
        check(self._L.nk_stage_reserve(self._h, nbytes, nseq, C.byref(b), C.byref(o)))
        return b.value, o.value

    def process_staged(self, nbytes: int, nseq: int, mode: int = 0) -> None:
        check(self._L.nk_process_staged(self._h, nbytes, nseq, mode))

    def stage_reserve_packed(self, nbases: int, nseq: int) -> Tuple[int, int, int]:
        c, x, o = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._L.nk_stage_reserve_packed(self._h, nbases, nseq, C.byref(c), C.byref(x), C.byref(o)))
        return c.value, x.value, o.value

    def process_staged_packed(self, nbases: int, nseq: int, mode: int = 0, has_other: bool = True) -> None:
        check(self._L.nk_process_staged_packed(self._h, nbases, nseq, mode, int(has_other)))

    def synth_fill(self, dev_ptr: int, seed: int, start: int, n: int, flags: int = 0) -> None:
        check(self._L.nk_synth_fill(self._h, dev_ptr, seed, start, n, flags))

    def stream_accumulated(self) -> int:
        p = C.c_void_p()
        check(self._L.nk_stream_accumulated(self._h, C.byref(p)))
        return p.value

    def stream_finish(self) -> None:
        check(self._L.nk_stream_finish(self._h))

    # sharded-pool multi-GPU mode (reduce-scatter fused into the LIF kernel over NVLink peer memory)
    def dist_export(self) -> Tuple[bytes, int]:
        buf = C.create_string_buffer(64)
        raw = C.c_void_p()
        check(self._L.nk_dist_export(self._h, buf, C.byref(raw)))
        return buf.raw, raw.value

    def dist_setup(self, rank: int, world: int, handles: Optional[bytes] = None, raw_ptrs: Optional[Sequence[int]] = None) -> None:
        if raw_ptrs is not None:
            arr = (C.c_void_p * world)(*[C.c_void_p(p) for p in raw_ptrs])
            check(self._L.nk_dist_setup(self._h, rank, world, None, arr))
        else:
            check(self._L.nk_dist_setup(self._h, rank, world, C.c_char_p(handles), None))

    def dist_post(self) -> Tuple[int, int, int]:
        p, n64, each = C.c_void_p(), C.c_uint64(), C.c_uint64()
        check(self._L.nk_dist_post(self._h, C.byref(p), C.byref(n64), C.byref(each)))
        return p.value, n64.value, each.value

    def dist_complete(self, dev_gathered: int, n_each: int) -> None:
        check(self._L.nk_dist_complete(self._h, dev_gathered, n_each))

    def dist_run(self) -> None:
        """sharded-pool exchange with peer-memory signalling only (no NCCL / host barrier per job)"""
        check(self._L.nk_dist_run(self._h))

    def dist_slice(self) -> Tuple[int, int]:
        lo, ln = C.c_uint64(), C.c_uint64()
        check(self._L.nk_dist_slice(self._h, C.byref(lo), C.byref(ln)))
        return lo.value, ln.value

    def cuda_stream(self) -> int:
        p = C.c_void_p()
        check(self._L.nk_cuda_stream(self._h, C.byref(p)))
        return p.value or 0

    def synchronize(self) -> None:
        check(self._L.nk_synchronize(self._h))


def debug_shard(offsets: np.ndarray, k: int, world: int, rank: int) -> Tuple[int, np.ndarray]:
    """(start, piece offsets relative to start) of member `rank` in the shard plan of a batch (host only)."""
    offsets = np.ascontiguousarray(offsets, np.uint64)
    out = np.zeros(offsets.size, np.uint64)
    start, n = C.c_uint64(), C.c_uint64()
    check(_lib.lib().nk_debug_shard(offsets.ctypes.data, offsets.size - 1, k, world, rank, C.byref(start),
                                    out.ctypes.data, C.byref(n)))
    return int(start.value), out[: n.value + 1].copy()


def device_count() -> int:
    n = C.c_int32()
    check(_lib.lib().nk_device_count(C.byref(n)))
    return n.value


def debug_mod(values: np.ndarray, pool_size: int, which: int = 0) -> np.ndarray:
    """values % pool_size computed by the device's exact-modulo routine (parity tap)."""
    v = np.ascontiguousarray(values, np.uint64)
    out = np.zeros(max(v.size, 1), np.uint64)
    check(_lib.lib().nk_debug_mod(_ptr(v), v.size, pool_size, which, out.ctypes.data))
    return out[: v.size]


def pack_kmer(kmer: BytesLike) -> int:
    """reference src/utils.rs:26-39"""
    a = np.frombuffer(kmer, np.uint8) if not isinstance(kmer, np.ndarray) else np.ascontiguousarray(kmer, np.uint8)
    return int(_lib.lib().nk_pack_kmer(_ptr(a), a.size))


def pack_kmer_py(kmer: BytesLike) -> int:
    """src/python.rs:50-52"""
    return pack_kmer(kmer)


class PySpikingCounter:
    """Surface of the reference's PyO3 class (src/python.rs:7-48): `PySpikingCounter(k, pool_size)`
    with the LIF constants python.rs hard-codes (threshold 1.0, leak 0.95, refractory 2, cost 1.0)."""

    def __init__(self, k: int, pool_size: int):
        self._c = SpikingKmerCounter(k, 1.0, 0.95, 2, 1.0, pool_size, False)
        self._c.enable_exact_counts(True)  # get_counts() walks the exact table

    def process_file(self, path: str) -> None:
        # python.rs:21-28 feeds every record to process_sequence, in file order
        from .fastx import read_fastx
        for seq in read_fastx(path):
            self._c.process_sequence(seq)

    def get_counts(self) -> dict:
        # python.rs:31-40: {key.to_string(): count}
        keys, counts = self._c.exact_table()
        return {str(int(k)): int(c) for k, c in zip(keys, counts)}

    def energy_used(self) -> float:
        return self._c.energy_used()
