"""neurokmer_b200 — B200-native (sm_100a) implementation of NeuroKmer's counting hot path.

The product is `libneurokmer.so` (C ABI, include/neurokmer.h).  This package is the
host-side mirror of the reference's public counter API over that ABI.  Importing the
counter classes requires the built shared library; there is no CPU fallback.
"""
from .counter import (PinnedBuffer, PySpikingCounter, SpikingKmerCounter, device_count, flatten,  # noqa: F401
                      pack_bases, pack_kmer, pack_kmer_py)
from ._lib import NkError  # noqa: F401

__all__ = ["SpikingKmerCounter", "PySpikingCounter", "PinnedBuffer", "device_count", "flatten", "pack_bases", "pack_kmer", "pack_kmer_py", "NkError"]
