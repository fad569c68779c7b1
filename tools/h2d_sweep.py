"""End-to-end time of the bench job from PINNED host memory: in-place reads by the count kernel (zero-copy) against
DMA copies in chunks of NK_H2D_CHUNK_MB pipelined with one count launch per chunk.  One line per setting."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import bench
    from neurokmer_b200 import PinnedBuffer, SpikingKmerCounter
    from neurokmer_b200.devmem import device_to_numpy
    c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, 2_000_000, True)
    n = bench.NBASES
    offsets = np.concatenate([[0], np.cumsum(bench.SEQ_LENS)]).astype(np.uint64)
    db, _ = c.stage_reserve(n, 7)
    c.synth_fill(db, 2, 0, n, 3); c.synchronize()
    pin = PinnedBuffer(n); pin.array[:] = device_to_numpy(db, n)
    ts = []
    for it in range(25):
        c.reset()
        t0 = time.perf_counter()
        c.stream_begin(); c.stream_push(pin.array, offsets); c.stream_end(); top = c.top_abundant_neurons(20)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = sorted(ts[5:])
    print("zerocopy=%s chunk_mb=%s plan=%s: e2e median %.3f ms, min %.3f ms, spikes %d" % (
        os.environ.get("NK_ZEROCOPY", "1"), os.environ.get("NK_H2D_CHUNK_MB", "-"), os.environ.get("NK_H2D_PLAN", "-"),
        ts[len(ts) // 2], ts[0], c.energy.total_spikes()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, NK_ZEROCOPY="1"))
    for mb in (8, 16, 32):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, NK_ZEROCOPY="0", NK_H2D_CHUNK_MB=str(mb)))
    for plan in ("32,32,24,12,6,2", "32,32,32,8,3,1", "24,24,24,16,12,6,2", "16,16,16,16,16,16,8,3,1", "32,32,16,16,8,4,2,1",
                 "8,32,32,24,8,3,1", "4,16,32,32,16,8,4,1"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, NK_ZEROCOPY="0", NK_H2D_PLAN=plan))
