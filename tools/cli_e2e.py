#!/usr/bin/env python
"""File -> result block wall time of the CLI on the README-sized input (115 MB FASTA, 7 sequences,
k=31, pool 2M, canonical, streaming), for the record next to the reference's published 7.6 min.
Run under gpurun:  python tools/cli_e2e.py"""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.synth import synth_bases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = "/tmp/nk_cli_115mb.fasta"
lens = [30_000_000, 25_000_000, 20_000_000, 15_000_000, 10_000_000, 8_000_000, 5_000_000]
t0 = time.time()
with open(path, "wb") as f:
    start = 0
    for i, n in enumerate(lens):
        s = synth_bases(2, start, n, 3)
        start += n
        f.write(b">seq%d synthetic\n" % i)
        rows = s[: n // 60 * 60].reshape(-1, 60)
        out = np.empty((rows.shape[0], 61), np.uint8); out[:, :60] = rows; out[:, 60] = 10
        f.write(out.tobytes())
        if n % 60:
            f.write(s[n // 60 * 60:].tobytes() + b"\n")
print(f"wrote {os.path.getsize(path) / 1e6:.1f} MB in {time.time() - t0:.1f} s", flush=True)
exe = os.path.join(ROOT, "neurokmer_b200", "neurokmer")
for extra in (["--no-uniques"], [], ["--exact"]):
    for rep in range(2):
        t0 = time.time()
        p = subprocess.run([exe, "-i", path, "-k", "31", "--pool-size", "2000000", "--canonical", "--streaming", "--timing"] + extra,
                           text=True, capture_output=True, check=True)
        out, dt = p.stdout, time.time() - t0
    print(p.stderr.strip())
    print(f"CLI {' '.join(extra) or '(default)'}: {dt:.3f} s wall (process start to exit, second run)")
    print("\n".join(out.splitlines()[:4] + out.splitlines()[-4:]))
