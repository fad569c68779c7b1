import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")
    # the tests exercise the BUILT shared library; build it (nvcc, sm_100a) if the tree has none yet
    lib = os.path.join(ROOT, "neurokmer_b200", "libneurokmer.so")
    if not os.path.exists(lib):
        from neurokmer_b200 import build as nkbuild
        nkbuild.build()


@pytest.fixture(scope="session")
def coracle():
    """The CPU oracle (oracle/nk_oracle.c through ctypes) — the CHECKER, never the product."""
    from oracle.oracle_py import COracle
    return COracle()


def random_dna(rng: np.random.Generator, n: int, p_n: float = 0.0, p_lower: float = 0.0, p_iupac: float = 0.0) -> bytes:
    a = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n)
    if p_lower:
        m = rng.random(n) < p_lower
        a = np.where(m, a | 0x20, a)
    if p_n:
        m = rng.random(n) < p_n
        a = np.where(m, ord("N"), a)
    if p_iupac:
        m = rng.random(n) < p_iupac
        a = np.where(m, rng.choice(np.frombuffer(b"RYKMSWBDHVUnry-.*0\x00\xff\x01@[`{", np.uint8), size=n), a)
    return a.astype(np.uint8).tobytes()
