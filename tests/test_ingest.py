"""Host-side read ingest of libtagpu.so (tagpu_load_reads; row a14 of SURVEY.md §8: sequence = line 2 of every 4,
/root/reference/src/get_buffer.c:339-348).  Plain FASTQ goes through the parallel chunked path (exact newline counting,
16 MB chunks), gzip and FASTA through the serial one; all must produce the stream the oracle's loader produces."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

from turingassembler_b200.api import free_reads, load_reads


def _stream(files, threads):
    addr, n = load_reads(files, threads)
    out = bytes((C.c_uint8 * n).from_address(addr)) if n else b""
    free_reads(addr)
    return out


def _fastq(reads, eol=b"\n", last_newline=True):
    rec = b"".join(b"@r%d some text" % i + eol + r + eol + b"+" + eol + b"@" * len(r) + eol for i, r in enumerate(reads))
    return rec if last_newline else rec[: -len(eol)]


def test_small_shapes(oracle, tmp_path):
    rng = np.random.default_rng(1)
    reads = [bytes(rng.choice(list(b"ACGTN"), size=int(n)).astype(np.uint8)) for n in rng.integers(0, 300, size=500)]
    cases = {
        "plain": _fastq(reads), "crlf": _fastq(reads, b"\r\n"), "no_last_newline": _fastq(reads, last_newline=False),
        "crlf_no_last": _fastq(reads, b"\r\n", last_newline=False), "one": _fastq(reads[:1]), "empty": b"",
    }
    for name, data in cases.items():
        p = tmp_path / f"{name}.fq"
        p.write_bytes(data)
        want = oracle.load_reads([str(p)]).tobytes()
        for threads in (1, 3, 8):
            assert _stream([str(p)], threads) == want, (name, threads)
    gz = tmp_path / "x.fq.gz"
    gz.write_bytes(gzip.compress(cases["plain"]))
    assert _stream([str(gz), str(tmp_path / "crlf.fq")], 4) == oracle.load_reads([str(gz), str(tmp_path / "crlf.fq")]).tobytes()
    fa = tmp_path / "x.fa"
    fa.write_bytes(b">a\nACGT\nACGG\n>b\nTTTT\n")
    assert _stream([str(fa)], 2) == b"ACGTACGG\nTTTT\n"


def test_multi_chunk_file(oracle, tmp_path):
    """> 2 chunks of 16 MB, lines straddling the chunk borders, CRLF so that a border can fall between \\r and \\n"""
    rng = np.random.default_rng(2)
    n, L = 70_000, 251
    seqs = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(n, L))
    rec = np.empty((n, 12 + 2 + L + 2 + 3 + L + 2), np.uint8)
    ids = np.arange(n)
    rec[:, 0] = ord("@")
    for d in range(11):
        rec[:, 1 + d] = ord("0") + (ids // 10 ** (10 - d)) % 10
    o = 12
    rec[:, o:o + 2] = (13, 10); o += 2
    rec[:, o:o + L] = seqs; o += L
    rec[:, o:o + 2] = (13, 10); o += 2
    rec[:, o] = ord("+"); rec[:, o + 1:o + 3] = (13, 10); o += 3
    rec[:, o:o + L] = ord("I"); o += L
    rec[:, o:o + 2] = (13, 10)
    p = tmp_path / "big.fq"
    p.write_bytes(rec.tobytes())
    assert rec.nbytes > 2 * (16 << 20)
    want = (seqs.astype(np.uint8).tobytes(), n)
    got = _stream([str(p)], 8)
    assert len(got) == n * (L + 1)
    a = np.frombuffer(got, np.uint8).reshape(n, L + 1)
    assert np.all(a[:, L] == 10) and a[:, :L].tobytes() == want[0]
    assert got == _stream([str(p)], 1)
