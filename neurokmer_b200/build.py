"""Builds libneurokmer.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

`python -m neurokmer_b200.build` or `neurokmer_b200.build.build()`.  nvcc
cross-compiles without a GPU; the resulting .so travels with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libneurokmer.so")
CLI = os.path.join(HERE, "neurokmer")
SOURCES = ["nk_count.cu", "nk_lif.cu", "nk_topn.cu", "nk_misc.cu", "nk_post.cu", "nk_exact.cu", "nk_api.cu", "nk_multi.cu", "nk_parse.cu", "nk_ingest.cu", "nk_fastx.cpp", "nk_pack.cpp", "nk_decomp.cpp"]
HEADERS = ["nk_device.cuh", "nk_kernels.cuh", "nk_host.h", "nk_internal.h", os.path.join("..", "..", "include", "neurokmer.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "--use_fast_math=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    jobs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [sp] + hdrs):
            jobs.append([nvcc] + flags + ["-c", sp, "-o", obj])
    if jobs:
        # the translation units are independent: compile them side by side (nk_count.cu alone takes about a minute)
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)

        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(run, jobs))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt", "-lz"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    cli_src = os.path.join(CSRC, "nk_cli.cpp")
    if os.path.exists(cli_src) and (force or _stale(CLI, [cli_src, LIB] + hdrs)):
        cmd = ["g++", "-O2", "-std=c++17", cli_src, "-o", CLI, "-L" + HERE, "-lneurokmer", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
