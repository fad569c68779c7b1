"""Minimal FASTA/FASTQ record iterator and writers (host I/O helpers for tests, the
Python surface and bench input files).  Record rules follow the reference's use of
needletail (src/utils.rs:9-24): sequence = line(s) with terminators removed, iteration
stops at the first malformed record.  The production file path is the C++ reader behind
nk_process_file; this module is only used where the reference itself is per-record Python
(src/python.rs:21-28)."""
from __future__ import annotations

from typing import Iterator


def _open(path: str):
    with open(path, "rb") as f:
        magic = f.read(4)
    if magic[:2] == b"\x1f\x8b":  # compression sniffed from the magic bytes like needletail does
        import gzip
        return gzip.open(path, "rb")
    if magic[:3] == b"BZh":
        import bz2
        return bz2.open(path, "rb")
    if magic == b"\xfd7zX":
        import lzma
        return lzma.open(path, "rb")
    if magic == b"\x28\xb5\x2f\xfd":
        import io
        try:
            import pyarrow as pa  # the only zstd decoder in this image's Python (3.12 has no zstd module)
        except ImportError as e:
            raise OSError(f"{path}: zstd input needs pyarrow in this Python reader (the C++ reader decodes it itself)") from e
        with open(path, "rb") as f:
            return io.BytesIO(pa.input_stream(f, compression="zstd").read())
    return open(path, "rb")


def read_fastx(path: str) -> Iterator[bytes]:
    with _open(path) as f:
        first = f.read(1)
        if not first:
            raise OSError(f"{path}: empty file")
        if first not in (b">", b"@"):
            raise OSError(f"{path}: not FASTA/FASTQ")
        f.seek(0)
        if first == b">":
            seq, started = [], False
            for line in f:
                if line.startswith(b">"):
                    if started:
                        yield b"".join(seq)
                    seq, started = [], True
                else:
                    seq.append(line.rstrip(b"\r\n").replace(b"\r", b""))
            if started:
                yield b"".join(seq)
        else:
            while True:
                h = f.readline()
                if not h:
                    return
                if not h.startswith(b"@"):
                    return
                s = f.readline().rstrip(b"\r\n")
                p = f.readline()
                q = f.readline().rstrip(b"\r\n")
                if not p.startswith(b"+") or len(q) != len(s):
                    return
                yield s


def write_fasta(path: str, seqs, width: int = 60) -> None:
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            s = bytes(s)
            f.write(b">seq%d synthetic\n" % i)
            for j in range(0, len(s), width):
                f.write(s[j:j + width] + b"\n")


def write_fastq(path: str, seqs) -> None:
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            s = bytes(s)
            f.write(b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
