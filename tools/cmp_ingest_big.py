"""serial vs parallel FASTA ingest on a ~2 GB file (run under gpurun)"""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.synth import synth_bases
path = "/tmp/nk_big.fasta"
with open(path, "wb") as f:
    for i in range(20):
        n = 100_000_000
        s = synth_bases(7, i * n, n, 1)
        f.write(b">chr%d\n" % i)
        rows = s[: n // 60 * 60].reshape(-1, 60)
        out = np.empty((rows.shape[0], 61), np.uint8); out[:, :60] = rows; out[:, 60] = 10
        f.write(out.tobytes()); f.write(s[n // 60 * 60:].tobytes() + b"\n")
print("file MB", os.path.getsize(path) / 1e6, flush=True)
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neurokmer_b200", "neurokmer")
for thr in (1, 8, 1, 8):
    p = subprocess.run([exe, "-i", path, "-k", "31", "--pool-size", "16000000", "--canonical", "--streaming", "--timing"],
                       env=dict(os.environ, NK_FASTA_THREADS=str(thr)), capture_output=True, text=True)
    t = {l.split()[1]: float(l.split(" at ")[1].split()[0]) for l in p.stderr.splitlines() if "[timing]" in l}
    print(f"threads={thr}: nk_process_file {t['nk_process_file'] - t['nk_create']:.3f} s", p.stdout.splitlines()[-4], flush=True)
os.remove(path)
