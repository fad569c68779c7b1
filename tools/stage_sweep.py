"""Where does the host -> device staging of PAGEABLE bytes spend its time?  Sweeps the staging pool's thread
count and piece size (fresh process per point: both are read once) and times nk_stream_push of a 113 MB
pageable batch and nk_process_file of the same bases as a FASTA file.  Prints one line per point."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import bench
    from neurokmer_b200 import SpikingKmerCounter
    from neurokmer_b200.devmem import device_to_numpy
    c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, 2_000_000, True)
    n = bench.NBASES
    offsets = np.concatenate([[0], np.cumsum(bench.SEQ_LENS)]).astype(np.uint64)
    db, _ = c.stage_reserve(n, 7)
    c.synth_fill(db, 2, 0, n, 3); c.synchronize()
    bases = device_to_numpy(db, n).copy()
    path = "/dev/shm/nk_sweep_%d.fa" % os.getpid()
    bench.write_fasta(path, bases, offsets)
    res = {}
    for name in ("push", "file"):
        ts = []
        for it in range(int(os.environ.get('NK_SWEEP_ITERS', 8))):
            c.reset()
            t0 = time.perf_counter()
            if name == "push":
                c.stream_begin(); c.stream_push(bases, offsets)
                t1 = time.perf_counter()
                c.stream_end(); c.top_abundant_neurons(20)
            else:
                c.process_file_streaming(path)
                t1 = time.perf_counter()
                c.top_abundant_neurons(20)
            t2 = time.perf_counter()
            ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
        ts = ts[3:]
        jobs = sorted(t[1] for t in ts)
        res[name] = (min(t[0] for t in ts), jobs[0], jobs[len(jobs) // 2], jobs[-1])
    os.unlink(path)
    print("threads=%s piece_kb=%s affinity=%s  push: stage %.2f ms, job min %.2f median %.2f max %.2f ms | file: %.2f ms, job min %.2f median %.2f max %.2f ms" % (
        os.environ.get("NK_STAGE_THREADS", "dflt"), os.environ.get("NK_STAGE_PIECE_KB", "2048"),
        "off" if os.environ.get("NK_STAGE_NO_AFFINITY") else "on", *res["push"], *res["file"]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    print("cpus:", os.cpu_count(), flush=True)
    grid = [(int(a), int(b)) for a, b in (x.split(":") for x in os.environ.get("NK_SWEEP", "").split(",") if x)] or \
           [(t, p) for t in (1, 2, 4, 8, 12) for p in (512, 2048, 8192)]
    for threads, piece in grid:
        if True:
            env = dict(os.environ, NK_STAGE_THREADS=str(threads), NK_STAGE_PIECE_KB=str(piece))
            subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
    env = dict(os.environ, NK_STAGE_THREADS="12", NK_STAGE_NO_AFFINITY="1")
    subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
