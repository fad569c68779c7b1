"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel family once."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurokmer_b200 import SpikingKmerCounter, flatten

rng = np.random.default_rng(0)
seqs = [rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=n).tobytes() for n in (40_000, 150, 20, 0, 17_000, 31)]
b, o = flatten(seqs)
for canonical, pool, k in ((True, 10_000, 31), (False, 4096, 11)):
    c = SpikingKmerCounter(k, 1.0, 0.95, 2, 1.0, pool, canonical)
    c.enable_exact_counts(True)
    c.process_batch(b, o)                       # count(words) + fused post + exact finalize
    print(c.energy.total_spikes(), c.top_abundant_neurons(5)[:2], c.exact_table()[0].size)
    c.process_batch(b, o)                       # direct LIF path + separate top-N
    print(c.energy.total_spikes(), c.top_abundant_neurons(3000)[:1])
    c.process_sequence(seqs[0][:500])
    c.debug_kmers(seqs[0][:3000]); c.debug_hash(np.arange(100, dtype=np.uint64))
    c.reset(); c.stream_begin(); c.stream_push(b, o); c.stream_end(); print(c.top_abundant_neurons(20)[0])
    c.close()
print("sanitize case ok")
