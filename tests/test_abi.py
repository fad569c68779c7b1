"""CPU tests of the drop-in boundary: the shared library loads, exports exactly the symbols
include/neurokmer.h declares, and the ctypes struct layouts match the C ones.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "neurokmer.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    from neurokmer_b200 import _lib
    return _lib


def header_symbols():
    txt = open(HDR).read()
    return sorted(set(re.findall(r"NK_API\s+[\w\s\*]+?\b(nk_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    declared = header_symbols()
    assert len(declared) >= 35
    assert sorted(built.SYMBOLS) == declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", built.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r"\sT\s+(nk_\w+)", out)))
    assert exported == declared
    lib = built.lib()
    assert lib.nk_version().startswith(b"neurokmer-b200")


def test_rust_shim_binds_only_declared_symbols():
    """integration/rust/src/ffi.rs (source only — no rustc here) must not drift from the header: every
    `pub fn nk_*` it declares is an entry point of include/neurokmer.h, and the #[repr(C)] structs list the
    header's fields in the header's order."""
    ffi = open(os.path.join(ROOT, "integration", "rust", "src", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (nk_\w+)\s*\(", ffi))
    assert len(bound) >= 30 and bound <= set(header_symbols()), bound - set(header_symbols())
    hdr = open(HDR).read()
    for c_name, rs_name in (("nk_config", "NkConfig"), ("nk_top_entry", "NkTopEntry")):
        c_body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (c_name, c_name), hdr, re.S).group(1)
        c_fields = re.findall(r"\b(\w+);", re.sub(r"/\*.*?\*/", "", c_body, flags=re.S))
        rs_body = re.search(r"pub struct %s \{(.*?)\}" % rs_name, ffi, re.S).group(1)
        rs_fields = re.findall(r"pub (\w+):", rs_body)
        assert rs_fields == c_fields, (c_name, c_fields, rs_fields)


def test_struct_layouts_match_c(built, tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "neurokmer.h"\nint main(){'
                   'printf("%zu %zu %zu\\n", sizeof(nk_config), sizeof(nk_top_entry), sizeof(nk_timings));'
                   'printf("%zu %zu %zu %zu\\n", offsetof(nk_config,pool_size), offsetof(nk_config,spike_cost), offsetof(nk_config,threshold), offsetof(nk_config,device));'
                   'printf("%zu %zu %zu %zu\\n", offsetof(nk_timings,kmers), offsetof(nk_timings,h2d_bytes), offsetof(nk_timings,lif_path), offsetof(nk_timings,topn_launches));'
                   'return 0;}')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    rows = [list(map(int, l.split())) for l in subprocess.check_output([str(exe)], text=True).splitlines()]
    L = built
    assert rows[0] == [C.sizeof(L.NkConfig), C.sizeof(L.NkTopEntry), C.sizeof(L.NkTimings)]
    assert rows[1] == [L.NkConfig.pool_size.offset, L.NkConfig.spike_cost.offset, L.NkConfig.threshold.offset, L.NkConfig.device.offset]
    assert rows[2] == [L.NkTimings.kmers.offset, L.NkTimings.h2d_bytes.offset, L.NkTimings.lif_path.offset, L.NkTimings.topn_launches.offset]


def test_host_only_entry_points(built):
    """Entry points that need no device: defaults, pack_kmer, argument validation, error strings."""
    lib = built.lib()
    cfg = built.NkConfig()
    assert lib.nk_config_default(C.byref(cfg)) == 0
    # reference CLI defaults (src/main.rs:10-37, src/spiking_hash.rs:70)
    assert (cfg.k, cfg.pool_size, cfg.use_canonical, cfg.refractory, cfg.steps) == (31, 1_000_000, 0, 2, 1000)
    assert (cfg.threshold, round(cfg.leak, 6), cfg.spike_cost) == (1.0, 0.95, 1.0)
    from neurokmer_b200 import pack_kmer, pack_kmer_py
    from oracle import oracle_py as op
    for s in (b"", b"A", b"ACGT", b"ACGTN", b"nnacgtNN", b"T" * 32, b"GATTACA-GATTACA"):
        assert pack_kmer(s) == op.pack_kmer(s) == pack_kmer_py(s)
    h = C.c_void_p()
    for bad_k in (0, 33):
        cfg.k = bad_k
        assert lib.nk_create(C.byref(cfg), C.byref(h)) == built.NK_ERR_BAD_ARG
        assert b"k must be in [1,32]" in lib.nk_last_error()
    cfg.k, cfg.pool_size = 31, 0
    assert lib.nk_create(C.byref(cfg), C.byref(h)) == built.NK_ERR_BAD_ARG
    cfg.pool_size = 1 << 32
    assert lib.nk_create(C.byref(cfg), C.byref(h)) == built.NK_ERR_UNSUPPORTED
    assert lib.nk_destroy(None) == 0 and lib.nk_reset(None) == built.NK_ERR_BAD_ARG


def test_host_packer_matches_numpy_twin(built):
    """nk_pack_bases (every SIMD body this CPU has, and the threaded driver) against the numpy
    restatement of the "nk2" layout, over all 256 byte values and ragged lengths."""
    from neurokmer_b200 import pack_bases
    from oracle import oracle_py as op
    lib = built.lib()
    assert lib.nk_packed_code_words(0) == 0 and lib.nk_packed_code_words(17) == 2 and lib.nk_packed_other_words(33) == 2
    rng = np.random.default_rng(7)
    cases = [np.zeros(0, np.uint8), np.frombuffer(b"ACGTNacgtn", np.uint8), np.arange(256, dtype=np.uint8)]
    for n in (1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 1000, 4097, 200_003):
        a = rng.choice(np.frombuffer(b"ACGTacgtNnRY-*\x00\xff", np.uint8), size=n,
                       p=[.2, .2, .2, .2, .03, .03, .03, .03, .02, .01, .01, .01, .01, .01, .005, .005])
        cases.append(a)
    cases.append(rng.integers(0, 256, size=70_001, dtype=np.uint8))
    bodies = [0, 1]
    for b in (2, 3):
        rc = lib.nk_debug_pack_body(None, 0, None, None, b, None)
        assert rc in (0, built.NK_ERR_UNSUPPORTED)
        if rc == 0:
            bodies.append(b)
    for a in cases:
        want_c, want_o, want_n = op.pack_nk2(a)
        for body in bodies:
            codes, other, n_other = pack_bases(a, body=body)
            assert n_other == want_n, (a.size, body)
            assert (codes[: want_c.size] == want_c).all(), (a.size, body)
            assert (other[: want_o.size] == want_o).all(), (a.size, body)
        codes, other, n_other = pack_bases(a, threads=3)
        assert n_other == want_n and (codes[: want_c.size] == want_c).all() and (other[: want_o.size] == want_o).all()
    # multi-range split (grain 65536) on a larger array, 4 threads
    a = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=1_000_003, p=[.24, .25, .25, .25, .01])
    want_c, want_o, want_n = op.pack_nk2(a)
    codes, other, n_other = pack_bases(a, threads=4)
    assert n_other == want_n and (codes == want_c).all() and (other == want_o).all()
    assert lib.nk_pack_bases(None, 5, None, None, 1, None) == built.NK_ERR_BAD_ARG


def test_host_packer_against_golden_layout_vectors(built):
    """tests/golden/golden_nk2.json pins the "nk2" layout (first base in bits 31:30, `other` bit p%32 of word p//32)."""
    import json
    from neurokmer_b200 import pack_bases
    from oracle import oracle_py as op
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_nk2.json")))
    for row in g["rows"]:
        b = row["bases"].encode("latin1")
        codes, other, n_other = pack_bases(b)
        nc, no = (len(b) + 15) // 16, (len(b) + 31) // 32
        assert codes[:nc].tolist() == row["codes"] and other[:no].tolist() == row["other"] and n_other == row["n_other"]
        tc, to, tn = op.pack_nk2(b)
        assert tc.tolist() == row["codes"] and to.tolist() == row["other"] and tn == row["n_other"]


def test_no_cpu_fallback_without_device(built):
    """Without a usable sm_100 device the product path must fail loudly, not compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    from neurokmer_b200 import NkError, SpikingKmerCounter
    with pytest.raises(NkError) as ei:
        SpikingKmerCounter(31, 1.0, 0.95, 2, 1.0, 1000, True)
    assert ei.value.code == built.NK_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu/reference legs may use oracle/."""
    pkg = os.path.join(ROOT, "neurokmer_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "libnk_oracle", "nko_", "oracle_py", "nk_oracle"):
                    assert needle not in txt, (os.path.join(dirpath, f), needle)


def test_fastx_python_reader_matches_rules(tmp_path):
    from neurokmer_b200.fastx import read_fastx, write_fasta, write_fastq
    seqs = [b"ACGT" * 40, b"", b"NNNN", b"acgtn"]
    fa, fq = str(tmp_path / "a.fa"), str(tmp_path / "a.fq")
    write_fasta(fa, seqs); write_fastq(fq, seqs)
    assert list(read_fastx(fa)) == seqs and list(read_fastx(fq)) == seqs
    with open(fq, "ab") as f:
        f.write(b"@bad\nACGT\n+\nII\n@after\nAC\n+\nII\n")
    assert list(read_fastx(fq)) == seqs  # stops at the malformed record (utils.rs:17-20)


def _digest(seqs):
    h, nb = 0xcbf29ce484222325, 0
    for s in seqs:
        for b in bytes(s) + b"\xff":
            h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
        nb += len(s)
    return len(seqs), nb, h


def _cxx_digest(lib, path):
    n, nb, h = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = lib.nk_debug_fastx_digest(str(path).encode(), C.byref(n), C.byref(nb), C.byref(h))
    return rc, (n.value, nb.value, h.value)


def test_cxx_reader_formats_and_compression(built, tmp_path):
    """The C++ record reader behind nk_process_file (host only): FASTA / FASTQ, CRLF, multi-line records,
    gzip / bzip2 / xz / zstd sniffed from the magic bytes (needletail's reader), multi-member gzip and
    multi-frame zstd, truncated streams, stop at the first malformed record (src/utils.rs:17-20)."""
    import bz2
    import gzip
    import lzma
    import pyarrow as pa
    from neurokmer_b200.fastx import read_fastx, write_fasta, write_fastq
    lib = built.lib()
    rng = np.random.default_rng(12)
    seqs = [rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=n).tobytes() for n in (3_000_000, 0, 61, 60, 59, 1, 250_000)]
    fa, fq = tmp_path / "a.fa", tmp_path / "a.fq"
    write_fasta(str(fa), seqs); write_fastq(str(fq), seqs)
    want = _digest(seqs)
    zstd = pa.Codec("zstd")
    for src in (fa, fq):
        raw = src.read_bytes()
        assert _cxx_digest(lib, src) == (0, want)
        variants = {
            ".gz": gzip.compress(raw, 1), ".bz2": bz2.compress(raw, 1), ".xz": lzma.compress(raw, preset=0),
            ".zst": zstd.compress(raw, asbytes=True),
            ".crlf": raw.replace(b"\n", b"\r\n"),
        }
        half = len(raw) // 2
        cut = raw.rfind(b"\n", 0, half) + 1 if src is fa else 0
        if cut:
            # concatenated members / frames: gzip (MultiGzDecoder) and zstd read all of them
            variants[".multi.gz"] = gzip.compress(raw[:cut], 1) + gzip.compress(raw[cut:], 1)
            variants[".multi.zst"] = zstd.compress(raw[:cut], asbytes=True) + zstd.compress(raw[cut:], asbytes=True)
        for ext, blob in variants.items():
            p = tmp_path / (src.name + ext)
            p.write_bytes(blob)
            rc, got = _cxx_digest(lib, p)
            assert rc == 0 and got == want, (src.name, ext)
            assert _digest(list(read_fastx(str(p)))) == want, (src.name, ext, "python reader")
    # a truncated compressed stream ends the iteration early instead of failing
    for ext, blob in ((".bz2", bz2.compress(fa.read_bytes(), 1)), (".xz", lzma.compress(fa.read_bytes(), preset=0)),
                      (".zst", zstd.compress(fa.read_bytes(), asbytes=True))):
        p = tmp_path / ("trunc.fa" + ext)
        p.write_bytes(blob[: len(blob) // 2])
        rc, got = _cxx_digest(lib, p)
        assert rc == 0 and 0 < got[1] < want[1], ext
    # malformed FASTQ record: everything before it is kept
    with open(fq, "ab") as f:
        f.write(b"@bad\nACGT\n+\nII\n@after\nAC\n+\nII\n")
    assert _cxx_digest(lib, fq) == (0, want)
    # open errors
    assert _cxx_digest(lib, tmp_path / "missing.fa")[0] == built.NK_ERR_IO
    (tmp_path / "empty.fa").write_bytes(b"")
    assert _cxx_digest(lib, tmp_path / "empty.fa")[0] == built.NK_ERR_IO
    (tmp_path / "junk.txt").write_bytes(b"hello\n")
    assert _cxx_digest(lib, tmp_path / "junk.txt")[0] == built.NK_ERR_IO


def test_fasta_readers_fuzz_against_python_twin(built, tmp_path):
    """Serial C++ reader and the parallel ingest's window planner/parser against the Python reader on randomised
    FASTA: ragged line widths, CRLF, empty lines, empty records, '>' inside sequence lines, no final newline,
    headers longer than a window.  Both newline-stripping bodies (AVX-512 VBMI2 / portable) are exercised: the
    portable one in a child process with NK_NO_SIMD_STRIP=1."""
    import subprocess
    import sys
    from neurokmer_b200.fastx import read_fastx
    lib = built.lib()
    rng = np.random.default_rng(2024)
    alphabet = np.frombuffer(b"ACGTNacgtn>", np.uint8)
    paths = []
    for case in range(12):
        out = bytearray()
        nrec = int(rng.integers(1, 9))
        for r in range(nrec):
            out += b">" + bytes(rng.choice(np.frombuffer(b"abc >|xyz", np.uint8), size=int(rng.integers(0, 300 if case % 3 else 5000)))) + (b"\r\n" if case % 4 == 1 else b"\n")
            total = int(rng.choice([0, 1, 59, 60, 61, 1000, 70_000]))
            width = int(rng.choice([1, 7, 60, 61, 64, 80, 100_000]))
            seq = rng.choice(alphabet, size=total, p=[.22, .22, .22, .22, .03, .02, .02, .02, .02, .005, .005]).tobytes()
            pos = 0
            while pos < len(seq):
                line = seq[pos:pos + width]
                if line.startswith(b">"):
                    line = b"A" + line[1:]          # a line may CONTAIN '>' but must not start with it
                out += line + (b"\r\n" if case % 4 == 1 else b"\n")
                pos += width
                if case % 5 == 2 and rng.random() < 0.05:
                    out += b"\n"                     # empty line inside a record
        if case % 6 == 3 and out.endswith(b"\n"):
            out = out[:-1]                           # no final newline
        p = tmp_path / f"fuzz{case}.fa"
        p.write_bytes(bytes(out))
        paths.append(str(p))
    for path in paths:
        want = _digest(list(read_fastx(path)))
        assert _cxx_digest(lib, path) == (0, want), path
        for window in (64, 100, 1000, 4096, 1 << 20):
            n, nb, h = C.c_uint64(), C.c_uint64(), C.c_uint64()
            assert lib.nk_debug_fasta_windows_digest(path.encode(), window, C.byref(n), C.byref(nb), C.byref(h)) == 0
            assert (n.value, nb.value, h.value) == want, (path, window)
    # the portable stripping body, in a child process
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from neurokmer_b200 import _lib; lib = _lib.lib()\n"
            "for p in sys.argv[1:]:\n"
            "    n, nb, h = C.c_uint64(), C.c_uint64(), C.c_uint64()\n"
            "    assert lib.nk_debug_fastx_digest(p.encode(), C.byref(n), C.byref(nb), C.byref(h)) == 0\n"
            "    a = (n.value, nb.value, h.value)\n"
            "    assert lib.nk_debug_fasta_windows_digest(p.encode(), 1000, C.byref(n), C.byref(nb), C.byref(h)) == 0\n"
            "    print(*a, n.value, nb.value, h.value)\n") % ROOT
    env = dict(os.environ, NK_NO_SIMD_STRIP="1")
    outp = subprocess.check_output([sys.executable, "-c", code] + paths, text=True, env=env)
    for path, line in zip(paths, outp.splitlines()):
        want = _digest(list(read_fastx(path)))
        vals = tuple(int(x) for x in line.split())
        assert vals[:3] == want and vals[3:] == want, path
