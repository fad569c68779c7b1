import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, bench
from neurokmer_b200 import SpikingKmerCounter
from neurokmer_b200.devmem import device_to_numpy
c = SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, 2_000_000, True)
n = bench.NBASES
offsets = np.concatenate([[0], np.cumsum(bench.SEQ_LENS)]).astype(np.uint64)
db, _ = c.stage_reserve(n, 7); c.synth_fill(db, 2, 0, n, 3); c.synchronize()
bases = device_to_numpy(db, n).copy()
path = "/dev/shm/nk_trace.fa"; bench.write_fasta(path, bases, offsets)
ts = []
for it in range(60):
    c.reset(); t0 = time.perf_counter(); c.process_file_streaming(path); c.top_abundant_neurons(20); ts.append((time.perf_counter() - t0) * 1e3)
    print("job %.2f ms" % ts[-1], file=sys.stderr)
os.unlink(path)
