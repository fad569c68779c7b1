"""Tiny device-memory helpers over torch (plumbing for tests and bench: the library owns
its buffers and hands out raw device pointers)."""
from __future__ import annotations

import ctypes as C

import numpy as np


def _cudart():
    import torch  # noqa: F401  (loads libcudart into the process)
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    raise ImportError("libcudart not found")


_RT = None


def rt():
    global _RT
    if _RT is None:
        _RT = _cudart()
        _RT.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _RT.cudaMemcpy.restype = C.c_int
        _RT.cudaDeviceSynchronize.restype = C.c_int
    return _RT


def copy_d2d(dst: int, src: int, nbytes: int) -> None:
    rc = rt().cudaMemcpy(dst, src, nbytes, 3)
    if rc:
        raise RuntimeError(f"cudaMemcpy d2d failed: {rc}")


def copy_h2d(dst: int, arr: np.ndarray) -> None:
    arr = np.ascontiguousarray(arr)
    rc = rt().cudaMemcpy(dst, arr.ctypes.data, arr.nbytes, 1)
    if rc:
        raise RuntimeError(f"cudaMemcpy h2d failed: {rc}")


def device_to_numpy(src: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
    out = np.empty(nbytes, np.uint8)
    rc = rt().cudaMemcpy(out.ctypes.data, src, nbytes, 2)
    if rc:
        raise RuntimeError(f"cudaMemcpy d2h failed: {rc}")
    return out.view(dtype)
