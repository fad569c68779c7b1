"""ctypes bindings of libneurokmer.so — one Python function per C-ABI entry point
declared in include/neurokmer.h.  No compute happens in Python: every call below
lands in the sm_100a kernels.  If the shared library is missing, import fails
loudly (there is no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NEUROKMER_LIB: alternative build of the same library (kernel-variant experiments, tools/variants.sh)
LIB_PATH = os.environ.get("NEUROKMER_LIB") or os.path.join(HERE, "libneurokmer.so")

NK_OK = 0
NK_ERR_BAD_ARG, NK_ERR_IO, NK_ERR_CUDA, NK_ERR_NO_DEVICE, NK_ERR_OOM, NK_ERR_STATE, NK_ERR_UNSUPPORTED = range(1, 8)
NK_UNIQUES_NOT_COMPUTED = 0xFFFFFFFF

# every symbol include/neurokmer.h declares (tests/test_abi.py checks this list against the header)
SYMBOLS = [
    "nk_config_default", "nk_create", "nk_destroy", "nk_reset", "nk_last_error", "nk_version",
    "nk_set_steps", "nk_get_steps", "nk_process_batch", "nk_stream_begin", "nk_stream_push",
    "nk_stream_end", "nk_process_file", "nk_process_sequence", "nk_simulate", "nk_top_n",
    "nk_total_spikes", "nk_energy_used", "nk_enable_exact_counts", "nk_get_count", "nk_exact_table_size",
    "nk_copy_exact_table", "nk_copy_uniques", "nk_debug_kmers", "nk_debug_hash", "nk_debug_mod",
    "nk_copy_currents", "nk_copy_spike_counts", "nk_copy_voltages", "nk_copy_refractory",
    "nk_last_timings", "nk_debug_set_lif_path", "nk_calibrate", "nk_stage_reserve", "nk_process_staged", "nk_stream_accumulated",
    "nk_stream_finish", "nk_dist_export", "nk_dist_setup", "nk_dist_post", "nk_dist_complete", "nk_dist_slice",
    "nk_cuda_stream", "nk_synchronize", "nk_synth_fill", "nk_host_alloc",
    "nk_host_free", "nk_pack_kmer",
    "nk_packed_code_words", "nk_packed_other_words", "nk_pack_bases", "nk_process_batch_packed",
    "nk_stream_push_packed", "nk_debug_kmers_packed", "nk_debug_pack_body", "nk_stage_reserve_packed",
    "nk_process_staged_packed", "nk_debug_fastx_digest", "nk_dist_run",
    "nk_debug_fasta_windows_digest", "nk_uniques_begin", "nk_uniques_push", "nk_uniques_push_packed", "nk_uniques_end", "nk_set_file_uniques",
    "nk_debug_set_fold_limit", "nk_device_count", "nk_create_multi", "nk_group_size", "nk_debug_shard", "nk_debug_parse_file", "nk_debug_stage_file",
]


class NkConfig(C.Structure):
    _fields_ = [
        ("k", C.c_uint32), ("refractory", C.c_uint32), ("pool_size", C.c_uint64), ("steps", C.c_uint64),
        ("spike_cost", C.c_double), ("threshold", C.c_float), ("leak", C.c_float),
        ("use_canonical", C.c_int32), ("device", C.c_int32),
    ]


class NkTopEntry(C.Structure):
    _fields_ = [("idx", C.c_uint64), ("spikes", C.c_uint64), ("uniques", C.c_uint32), ("_pad", C.c_uint32)]


class NkTimings(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("mark_ms", C.c_float), ("count_ms", C.c_float), ("fold_ms", C.c_float),
        ("lif_ms", C.c_float), ("topn_ms", C.c_float), ("total_ms", C.c_float),
        ("kmers", C.c_uint64), ("launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("lif_path", C.c_int32), ("_pad", C.c_int32), ("topn_launches", C.c_uint64),
        ("post_ms", C.c_float), ("exch_wait_ms", C.c_float), ("exch_reduce_ms", C.c_float), ("merge_ms", C.c_float),
        ("exch_bytes", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "_pad"}


class NkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libneurokmer error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m neurokmer_b200.build` "
            "(nvcc, sm_100a). neurokmer_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    P = C.POINTER
    sig = {
        "nk_config_default": (i32, [P(NkConfig)]),
        "nk_create": (i32, [P(NkConfig), P(vp)]),
        "nk_destroy": (i32, [vp]),
        "nk_device_count": (i32, [P(C.c_int32)]),
        "nk_create_multi": (i32, [P(NkConfig), P(C.c_int32), C.c_int32, P(vp)]),
        "nk_group_size": (i32, [vp, P(C.c_int32)]),
        "nk_debug_shard": (i32, [vp, u64, u32, C.c_int32, C.c_int32, P(u64), vp, P(u64)]),
        "nk_reset": (i32, [vp]),
        "nk_last_error": (C.c_char_p, []),
        "nk_version": (C.c_char_p, []),
        "nk_set_steps": (i32, [vp, u64]),
        "nk_get_steps": (i32, [vp, P(u64)]),
        "nk_process_batch": (i32, [vp, vp, vp, u64]),
        "nk_stream_begin": (i32, [vp]),
        "nk_stream_push": (i32, [vp, vp, vp, u64]),
        "nk_stream_end": (i32, [vp]),
        "nk_process_file": (i32, [vp, C.c_char_p, i32]),
        "nk_process_sequence": (i32, [vp, vp, u64]),
        "nk_simulate": (i32, [vp]),
        "nk_top_n": (i32, [vp, u64, P(NkTopEntry), P(u64)]),
        "nk_total_spikes": (i32, [vp, P(u64)]),
        "nk_energy_used": (i32, [vp, P(C.c_double)]),
        "nk_enable_exact_counts": (i32, [vp, i32]),
        "nk_get_count": (i32, [vp, u64, P(u32), P(C.c_int32)]),
        "nk_exact_table_size": (i32, [vp, P(u64)]),
        "nk_copy_exact_table": (i32, [vp, vp, vp]),
        "nk_copy_uniques": (i32, [vp, vp]),
        "nk_debug_kmers": (i32, [vp, vp, u64, vp, vp, vp, vp, P(u64)]),
        "nk_debug_hash": (i32, [vp, vp, u64, vp, vp]),
        "nk_debug_mod": (i32, [vp, u64, u64, i32, vp]),
        "nk_copy_currents": (i32, [vp, vp]),
        "nk_copy_spike_counts": (i32, [vp, vp]),
        "nk_copy_voltages": (i32, [vp, vp]),
        "nk_copy_refractory": (i32, [vp, vp]),
        "nk_last_timings": (i32, [vp, P(NkTimings)]),
        "nk_debug_set_lif_path": (i32, [vp, i32]),
        "nk_debug_set_fold_limit": (i32, [vp, u64]),
        "nk_calibrate": (i32, [vp, i32, P(C.c_double)]),
        "nk_stage_reserve": (i32, [vp, u64, u64, P(vp), P(vp)]),
        "nk_process_staged": (i32, [vp, u64, u64, i32]),
        "nk_stream_accumulated": (i32, [vp, P(vp)]),
        "nk_stream_finish": (i32, [vp]),
        "nk_dist_export": (i32, [vp, vp, P(vp)]),
        "nk_dist_setup": (i32, [vp, i32, i32, vp, P(vp)]),
        "nk_dist_post": (i32, [vp, P(vp), P(u64), P(u64)]),
        "nk_dist_complete": (i32, [vp, vp, u64]),
        "nk_dist_slice": (i32, [vp, P(u64), P(u64)]),
        "nk_dist_run": (i32, [vp]),
        "nk_uniques_begin": (i32, [vp, u64]),
        "nk_uniques_push": (i32, [vp, vp, vp, u64]),
        "nk_uniques_push_packed": (i32, [vp, vp, vp, vp, u64]),
        "nk_uniques_end": (i32, [vp]),
        "nk_set_file_uniques": (i32, [vp, u64]),
        "nk_cuda_stream": (i32, [vp, P(vp)]),
        "nk_synchronize": (i32, [vp]),
        "nk_synth_fill": (i32, [vp, vp, u64, u64, u64, u32]),
        "nk_host_alloc": (i32, [P(vp), u64]),
        "nk_host_free": (i32, [vp]),
        "nk_pack_kmer": (u64, [vp, u64]),
        "nk_packed_code_words": (u64, [u64]),
        "nk_packed_other_words": (u64, [u64]),
        "nk_pack_bases": (i32, [vp, u64, vp, vp, i32, P(u64)]),
        "nk_process_batch_packed": (i32, [vp, vp, vp, vp, u64]),
        "nk_stream_push_packed": (i32, [vp, vp, vp, vp, u64]),
        "nk_debug_kmers_packed": (i32, [vp, vp, vp, u64, vp, vp, vp, vp, P(u64)]),
        "nk_debug_pack_body": (i32, [vp, u64, vp, vp, i32, P(u64)]),
        "nk_stage_reserve_packed": (i32, [vp, u64, u64, P(vp), P(vp), P(vp)]),
        "nk_process_staged_packed": (i32, [vp, u64, u64, i32, i32]),
        "nk_debug_fastx_digest": (i32, [C.c_char_p, P(u64), P(u64), P(u64)]),
        "nk_debug_parse_file": (i32, [vp, C.c_char_p, P(u64), P(u64), P(u64)]),
        "nk_debug_stage_file": (i32, [vp, C.c_char_p, P(C.c_double), P(C.c_double)]),
        "nk_debug_fasta_windows_digest": (i32, [C.c_char_p, u64, P(u64), P(u64), P(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


def check(rc: int) -> None:
    if rc != NK_OK:
        raise NkError(rc, (lib().nk_last_error() or b"").decode("utf-8", "replace"))
