"""serial vs parallel FASTA ingest on a ~2 GB file (run under gpurun)"""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.synth import synth_bases
fmt = "fastq" if "--fastq" in sys.argv else "fasta"
path = "/tmp/nk_big." + fmt
if fmt == "fastq":
    # ~2 GB FASTQ: 6.4 M reads x 150 bp, 4-line records
    with open(path, "wb") as f:
        nreads, L = 800_000, 150
        for blk in range(8):
            s = synth_bases(9, blk * nreads * L, nreads * L, 0).reshape(nreads, L)
            hdr = np.frombuffer(b"@read/0000000000\n", np.uint8)
            rec = np.empty((nreads, hdr.size + L + 1 + 2 + L + 1), np.uint8)
            rec[:, :hdr.size] = hdr
            ids = np.arange(blk * nreads, (blk + 1) * nreads)
            for d in range(10):
                rec[:, 6 + 9 - d] = 48 + (ids // 10**d) % 10
            o = hdr.size
            rec[:, o:o + L] = s; rec[:, o + L] = 10; rec[:, o + L + 1] = ord("+"); rec[:, o + L + 2] = 10
            rec[:, o + L + 3:o + 2 * L + 3] = ord("I"); rec[:, o + 2 * L + 3] = 10
            f.write(rec.tobytes())
with open(path, "wb") if fmt == "fasta" else open(os.devnull, "wb") as f:
    for i in range(20 if fmt == "fasta" else 0):
        n = 100_000_000
        s = synth_bases(7, i * n, n, 1)
        f.write(b">chr%d\n" % i)
        rows = s[: n // 60 * 60].reshape(-1, 60)
        out = np.empty((rows.shape[0], 61), np.uint8); out[:, :60] = rows; out[:, 60] = 10
        f.write(out.tobytes()); f.write(s[n // 60 * 60:].tobytes() + b"\n")
print("file MB", os.path.getsize(path) / 1e6, flush=True)
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neurokmer_b200", "neurokmer")
for thr in (1, 8, 1, 8, 1, 8, 1, 8):
    p = subprocess.run([exe, "-i", path, "-k", "31", "--pool-size", "2000000", "--canonical", "--streaming", "--timing"],
                       env=dict(os.environ, NK_FASTA_THREADS=str(thr)), capture_output=True, text=True)
    t = {l.split()[1]: float(l.split(" at ")[1].split()[0]) for l in p.stderr.splitlines() if "[timing]" in l}
    print(f"threads={thr}: nk_process_file {t['nk_process_file'] - t['nk_create']:.3f} s", p.stdout.splitlines()[-4], flush=True)
os.remove(path)
