#!/usr/bin/env python3
"""A two-GPU job inside ONE process (nk_create_multi), for an ncu capture of the exchange traffic of post_kernel.

    python tools/nvlink_capture.py                    # plain run: prints the job's timings and exchange bytes
    ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum -k regex:post_kernel \
        --clock-control none --csv --log-file gpurun_out/r02_nvlink.csv python tools/nvlink_capture.py

The in-process group orders its GPUs with CUDA events (no kernel waits on another kernel), so a profiler that
serialises the launches cannot deadlock it.  Each GPU's post_kernel reads its neuron slice of the OTHER GPU's
accumulators over NVLink: pool/world × 4 bytes × (world-1) received per GPU and job — 2 MB at pool 1M, world 2.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import SpikingKmerCounter, device_count  # noqa: E402


def main():
    n = min(2, device_count())
    if n < 2:
        print("needs two GPUs")
        return 1
    pool = int(os.environ.get("NK_POOL", 1_000_000))
    rng = np.random.default_rng(2)
    lens = [10_000_000] * 22 + [5_000_000]
    seqs = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, sum(lens), dtype=np.uint8)]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, pool, True, devices=list(range(n)))
    for it in range(3):
        c.reset()
        c.stream_begin()
        c.stream_push(seqs, offsets)
        c.stream_finish()
        top = c.top_abundant_neurons(10)
        t = c.timings()
        print(f"job {it}: kmers {t['kmers']} count {t['count_ms']:.3f} ms post {t['post_ms']:.3f} ms "
              f"exchange wait/reduce/merge {t['exch_wait_ms']:.3f}/{t['exch_reduce_ms']:.3f}/{t['merge_ms']:.3f} ms "
              f"exch_bytes {t['exch_bytes']} top1 {top[0][:2]}")
    c.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
