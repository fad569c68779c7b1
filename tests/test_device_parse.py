"""Record parsing on the device (neurokmer_b200/csrc/nk_parse.cu) against the host reader (nk_fastx.cpp, the
twin of needletail's record rules as the reference uses them, reference src/utils.rs:9-24): same records,
same bases, iteration ends at the first malformed FASTQ record.  Then the whole file path against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import random_dna
from test_parity_gpu import REF, assert_state_equal, assert_topn_equal, make, oracle_counter

pytestmark = pytest.mark.gpu


def host_digest(path):
    from neurokmer_b200 import _lib
    n, nb, h = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _lib.check(_lib.lib().nk_debug_fastx_digest(str(path).encode(), C.byref(n), C.byref(nb), C.byref(h)))
    return n.value, nb.value, h.value


def fasta_bytes(rng, recs, width, eol=b"\n", final_eol=True):
    out = bytearray()
    for i, s in enumerate(recs):
        out += b">rec%d some description > with a bracket" % i + eol
        if width <= 0:
            out += s + eol
        else:
            for j in range(0, len(s), width):
                out += s[j:j + width] + eol
    if not final_eol and out.endswith(eol):
        del out[-len(eol):]
    return bytes(out)


FASTA_CASES = {
    "60col": lambda rng: fasta_bytes(rng, [random_dna(rng, n, 0.01, 0.05, 0.01) for n in (100_000, 0, 61, 60, 59, 1, 250_000)], 60),
    "crlf": lambda rng: fasta_bytes(rng, [random_dna(rng, n, 0.01) for n in (50_000, 7, 33_333)], 70, b"\r\n"),
    "no_final_newline": lambda rng: fasta_bytes(rng, [random_dna(rng, n) for n in (5000, 12345)], 80, final_eol=False),
    "single_line_records": lambda rng: fasta_bytes(rng, [random_dna(rng, n, 0.001) for n in (300_000, 20_000, 16384, 16383, 4096)], 0),
    "gt_inside_lines": lambda rng: b">h1\nACGT>ACGT\nAC>\n>h2\n>h3\n\n\nGG\r\rTT\n>h4" ,
    "header_only": lambda rng: b">only a header",
    "long_headers": lambda rng: b">" + b"x" * 40_000 + b"\nACGT\n>" + b"y>" * 9000 + b"\nTTTT\nGGGG\n",
    "blank_lines": lambda rng: b">a\n\n\nAC\n\nGT\n\n>b\n\n",
    "many_small": lambda rng: fasta_bytes(rng, [random_dna(rng, int(n)) for n in rng.integers(0, 90, size=20_000)], 60),
}


@pytest.mark.parametrize("name", sorted(FASTA_CASES))
def test_fasta_device_parser_equals_host_reader(tmp_path, name):
    rng = np.random.default_rng(hash(name) % 2**32)
    p = tmp_path / (name + ".fa")
    p.write_bytes(FASTA_CASES[name](rng))
    c = make(31, 1000)
    assert c.debug_parse_file(p) == host_digest(p)
    c.close()


def fastq_bytes(rng, lens, eol=b"\n", final_eol=True, breakage=None):
    out = bytearray()
    for i, n in enumerate(lens):
        s = random_dna(rng, int(n), 0.01)
        q = bytes(rng.integers(33, 74, size=int(n), dtype=np.uint8))
        hdr, sep = b"@read%d/1 extra" % i, b"+"
        if breakage and breakage[0] == i:
            kind = breakage[1]
            if kind == "no_at": hdr = b"read%d" % i
            elif kind == "no_plus": sep = b"-"
            elif kind == "empty_plus": sep = b""
            elif kind == "short_qual": q = q[:-1] if n else b"I"
            elif kind == "long_qual": q = q + b"I"
            elif kind == "at_in_qual_ok": q = b"@" + q[1:] if n else q
        out += hdr + eol + s + eol + sep + eol + q + eol
    if not final_eol and out.endswith(eol):
        del out[-len(eol):]
    return bytes(out)


FASTQ_CASES = {
    "reads150": lambda rng: fastq_bytes(rng, [150] * 5000 + [20, 0, 151, 1]),
    "crlf": lambda rng: fastq_bytes(rng, rng.integers(0, 300, size=3000), b"\r\n"),
    "no_final_newline": lambda rng: fastq_bytes(rng, [100, 50, 75], final_eol=False),
    "crlf_no_final_newline": lambda rng: fastq_bytes(rng, [100, 50, 75], b"\r\n", final_eol=False),
    "long_reads": lambda rng: fastq_bytes(rng, [40_000, 16_384, 100_000, 5]),
    "bad_header": lambda rng: fastq_bytes(rng, [150] * 400, breakage=(137, "no_at")),
    "bad_separator": lambda rng: fastq_bytes(rng, [150] * 400, breakage=(201, "no_plus")),
    "empty_separator": lambda rng: fastq_bytes(rng, [150] * 400, breakage=(3, "empty_plus")),
    "short_quality": lambda rng: fastq_bytes(rng, [150] * 400, breakage=(399, "short_qual")),
    "long_quality": lambda rng: fastq_bytes(rng, [150] * 400, breakage=(0, "long_qual")),
    "at_sign_in_quality": lambda rng: fastq_bytes(rng, [150] * 400, breakage=(77, "at_in_qual_ok")),
    "truncated_after_sequence": lambda rng: fastq_bytes(rng, [150] * 10) + b"@last\nACGT\n",
    "truncated_after_plus": lambda rng: fastq_bytes(rng, [150] * 10) + b"@last\nACGT\n+\n",
    "trailing_blank_lines": lambda rng: fastq_bytes(rng, [150] * 10) + b"\n\n\n",
    "one_record": lambda rng: b"@r\nACGTN\n+\nIIIII",
}


@pytest.mark.parametrize("name", sorted(FASTQ_CASES))
def test_fastq_device_parser_equals_host_reader(tmp_path, name):
    rng = np.random.default_rng(hash(name) % 2**32)
    p = tmp_path / (name + ".fq")
    p.write_bytes(FASTQ_CASES[name](rng))
    c = make(31, 1000)
    assert c.debug_parse_file(p) == host_digest(p)
    c.close()


def test_device_parser_fuzz(tmp_path):
    """random line widths, terminators and record counts, files that end inside any segment of the scan"""
    rng = np.random.default_rng(2024)
    c = make(31, 1000)
    for trial in range(30):
        nrec = int(rng.integers(1, 40))
        if trial % 2 == 0:
            recs = [random_dna(rng, int(n), 0.02, 0.1, 0.02) for n in rng.integers(0, 60_000, size=nrec)]
            data = fasta_bytes(rng, recs, int(rng.choice([0, 1, 7, 60, 61, 80, 1000])), b"\n" if rng.random() < 0.6 else b"\r\n",
                               final_eol=bool(rng.integers(0, 2)))
            p = tmp_path / f"f{trial}.fa"
        else:
            br = None
            if rng.random() < 0.5:
                br = (int(rng.integers(0, nrec)), str(rng.choice(["no_at", "no_plus", "empty_plus", "short_qual", "long_qual"])))
            data = fastq_bytes(rng, rng.integers(0, 3000, size=nrec), b"\n" if rng.random() < 0.6 else b"\r\n",
                               final_eol=bool(rng.integers(0, 2)), breakage=br)
            p = tmp_path / f"f{trial}.fq"
        p.write_bytes(data)
        assert c.debug_parse_file(p) == host_digest(p), (trial, p)
    c.close()


@pytest.mark.parametrize("fmt", ["fa", "fq"])
def test_file_path_device_equals_host_and_oracle(tmp_path, fmt, monkeypatch):
    """nk_process_file through the device parser == through the host reader == the oracle on the records;
    the uniques pass re-uses the parsed file that is still on the device"""
    from neurokmer_b200 import flatten
    rng = np.random.default_rng(8)
    k, pool = 21, 50_021
    if fmt == "fa":
        seqs = [random_dna(rng, n, 0.002, 0.02) for n in (900_000, 20, 0, 400_000, 21, 150)]
        data = fasta_bytes(rng, seqs, 60)
    else:
        lens = [150] * 8000 + [20, 21, 0, 300]
        seqs, data = [], bytearray()
        for i, n in enumerate(lens):
            s = random_dna(rng, n, 0.01)
            seqs.append(s)
            data += b"@r%d\n" % i + s + b"\n+\n" + b"I" * n + b"\n"
        data = bytes(data)
    p = tmp_path / ("x." + fmt)
    p.write_bytes(data)
    o = oracle_counter(k, pool)
    o.process_streaming([flatten(seqs)])
    dev, host = make(k, pool), make(k, pool)
    dev.set_file_uniques(20); host.set_file_uniques(20)
    dev.process_file_streaming(str(p))
    monkeypatch.setenv("NK_GPU_PARSE", "0")
    host.process_file_streaming(str(p))
    monkeypatch.delenv("NK_GPU_PARSE")
    assert dev.timings()["kmers"] == host.timings()["kmers"] == sum(max(0, len(s) - k + 1) for s in seqs)
    assert_state_equal(dev, o); assert_topn_equal(dev, o, 20)
    assert dev.top_abundant_neurons(20) == host.top_abundant_neurons(20)      # including the uniques column
    assert all(r[2] is not None for r in dev.top_abundant_neurons(20))
    # in-memory driver (process_parallel semantics) and carried state through the device path
    o.process_parallel(*flatten(seqs))
    dev.process_file_in_memory(str(p))
    assert_state_equal(dev, o)


def test_pageable_batches_take_the_staging_pool():
    """a large batch in plain malloc memory is staged by the pool of host threads (2 MiB pieces, pinned slots):
    same result as the pinned / small-batch paths, chunk boundaries (32 MiB) included"""
    from neurokmer_b200 import flatten
    rng = np.random.default_rng(10)
    k, pool = 31, 400_009
    seqs = [random_dna(rng, n, 0.001) for n in (40_000_000, 150, 9_000_000, 31, 20_000_000)]
    bases, offsets = flatten(seqs)
    o = oracle_counter(k, pool); o.process_streaming([(bases, offsets)])
    c = make(k, pool)
    c.stream_begin(); c.stream_push(bases, offsets); c.stream_end()
    assert_state_equal(c, o); assert_topn_equal(c, o, 20)
    os.environ["NK_STAGE_POOL"] = "0"
    try:
        d = make(k, pool)
        d.stream_begin(); d.stream_push(bases, offsets); d.stream_end()
        assert_state_equal(d, o)
    finally:
        del os.environ["NK_STAGE_POOL"]


@pytest.mark.parametrize("case", ["records_60col", "tiny_records", "crlf_single_lines", "more_records_than_the_offsets_buffer"])
def test_overlapped_fasta_file_path(tmp_path, case, monkeypatch):
    """Large FASTA files are parsed and counted chunk by chunk while the rest of the file is still being read (the
    scan carries its state from chunk to chunk; the record open at the end of a chunk ends "infinitely far").  Forced
    here with 1 MiB chunks: record boundaries next to chunk boundaries, records shorter than k, empty records piling
    up on one position, CRLF, and a file with more records than the overlapped path has room for (it falls back to
    the two-phase path).  Same pool state as the host reader's path and as the oracle on the records."""
    from neurokmer_b200 import flatten
    rng = np.random.default_rng(abs(hash(case)) % 2**32)
    k, pool = 31, 70_001
    if case == "records_60col":
        lens = [1_500_000, 20, 0, 0, 0, 31, 30, 900_000, 1, 1_048_576 - 40, 64, 700_000, 5]
        seqs = [random_dna(rng, n, 0.002, 0.02) for n in lens]
        data = fasta_bytes(rng, seqs, 60)
    elif case == "tiny_records":
        seqs = [random_dna(rng, int(n), 0.01) for n in rng.integers(0, 120, size=60_000)]
        data = fasta_bytes(rng, seqs, 70)
    elif case == "crlf_single_lines":
        seqs = [random_dna(rng, n, 0.001) for n in (1_200_000, 800_000, 40, 1_500_000)]
        data = fasta_bytes(rng, seqs, 0, b"\r\n", final_eol=False)
    else:
        seqs = [b"ACGTACGTACGTACGTACGTACGTACGTACGTAC"] * 1_060_000   # > 2^20 records of 34 bases
        data = b"".join(b">r\n" + s + b"\n" for s in seqs)
    p = tmp_path / "big.fa"
    p.write_bytes(data)
    assert len(data) > 3 * (1 << 20)
    o = oracle_counter(k, pool)
    o.process_streaming([flatten(seqs)])
    monkeypatch.setenv("NK_FILE_CHUNK_MB", "1")
    dev = make(k, pool)
    dev.set_file_uniques(20)
    dev.process_file_streaming(str(p))
    assert dev.timings()["kmers"] == sum(max(0, len(s) - k + 1) for s in seqs)
    assert_state_equal(dev, o); assert_topn_equal(dev, o, 20)
    rows = dev.top_abundant_neurons(20)
    monkeypatch.setenv("NK_GPU_PARSE", "0")
    host = make(k, pool)
    host.set_file_uniques(20)
    host.process_file_streaming(str(p))
    assert rows == host.top_abundant_neurons(20)           # including the uniques column (resident parsed file re-used)
    monkeypatch.delenv("NK_GPU_PARSE")
    # a second job on the same counter (carried state) and the exact-table mode go through the same chunks
    o.process_streaming([flatten(seqs)])
    dev.process_file_streaming(str(p))
    assert_state_equal(dev, o)
    if case == "records_60col":
        e = make(k, pool); e.enable_exact_counts(True)
        e.process_file_streaming(str(p))
        np.testing.assert_array_equal(e.currents(), dev.currents())
        assert int(e.exact_table()[1].sum()) == dev.timings()["kmers"]
