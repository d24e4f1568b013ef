"""FASTQ records parsed on the device (csrc/tagpu_fastq.cuh, row a14 of SURVEY.md §8: the sequence is line 2 of every 4,
/root/reference/src/get_buffer.c:339-348) against the oracle's loader and the host parser, on the shapes of test_ingest.py,
and the files entry point that uses it against the host-parsed build."""
import ctypes as C
import os

import numpy as np
import pytest

from turingassembler_b200 import Tagpu
from turingassembler_b200.api import free_reads, load_reads

pytestmark = pytest.mark.gpu


def _fastq(reads, eol=b"\n", last_newline=True):
    rec = b"".join(b"@r%d some text" % i + eol + r + eol + b"+" + eol + b"@" * len(r) + eol for i, r in enumerate(reads))
    return rec if last_newline else rec[: -len(eol)]


def _host_stream(files):
    addr, n = load_reads(files, 4)
    out = bytes((C.c_uint8 * n).from_address(addr)) if n else b""
    free_reads(addr)
    return out


def test_device_parser_matches_oracle_and_host(oracle, tmp_path):
    rng = np.random.default_rng(1)
    reads = [bytes(rng.choice(list(b"ACGTN"), size=int(n)).astype(np.uint8)) for n in rng.integers(0, 300, size=3000)]
    cases = {
        "plain": _fastq(reads), "crlf": _fastq(reads, b"\r\n"), "no_last_newline": _fastq(reads, last_newline=False),
        "crlf_no_last": _fastq(reads, b"\r\n", last_newline=False), "one": _fastq(reads[:1]),
        "header_only": b"@r0 nothing else\n", "header_only_no_newline": b"@r0",
        "ends_in_sequence": b"@r0\nACGTACGT", "ends_in_sequence_cr": b"@r0\r\nACGTACGT\r",
        "empty_sequence_lines": b"@a\n\n+\n\n@b\nACGT\n+\nIIII\n@c\n\n+\n\n",
        "three_lines": b"@a\nACGTT\n+\n", "sequence_line_empty_at_end": b"@a\n\n", "cr_only_tail": b"@a\n\r",
    }
    t = Tagpu(0)
    for name, data in cases.items():
        p = tmp_path / f"{name}.fq"
        p.write_bytes(data)
        want = oracle.load_reads([str(p)]).tobytes()
        assert _host_stream([str(p)]) == want, name
        assert t.parse_fastq([data]) == want, name
    # several files in one call, sizes around the 4 KB block and the 1 MB ring slot
    blobs = [cases["plain"], cases["crlf_no_last"], cases["one"], _fastq(reads * 9)]
    paths = []
    for i, b in enumerate(blobs):
        p = tmp_path / f"multi{i}.fq"
        p.write_bytes(b)
        paths.append(str(p))
    assert len(blobs[3]) > (2 << 20)
    assert t.parse_fastq(blobs) == oracle.load_reads(paths).tobytes()
    for cut in (4095, 4096, 4097, 65536 + 1):
        assert t.parse_fastq([blobs[3][:cut]]) == _host_stream_of_bytes(tmp_path, blobs[3][:cut])
    t.close()


def _host_stream_of_bytes(tmp_path, data):
    p = tmp_path / "cut.fq"
    p.write_bytes(data)
    return _host_stream([str(p)])


def test_files_entry_point_parses_on_the_device(tmp_path):
    """build_graph_from_scratch on two FASTQ files: the device-parsed build (TAGPU_DEVICE_PARSE=1) and the host-parsed build
    give the same graph counters"""
    import subprocess
    import sys
    import json
    rng = np.random.default_rng(5)
    genome = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=60_000)
    def reads(seed):
        r = np.random.default_rng(seed)
        starts = r.integers(0, len(genome) - 151, size=20_000)
        return [genome[s:s + 151].tobytes() for s in starts]
    (tmp_path / "r1.fq").write_bytes(_fastq(reads(1)))
    (tmp_path / "r2.fq").write_bytes(_fastq(reads(2), b"\r\n", last_newline=False))
    code = (
        "import json, sys\n"
        "from turingassembler_b200.api import build_graph_from_scratch, free_asm_graph\n"
        "g = build_graph_from_scratch(31, 4, 8, [sys.argv[1]], [sys.argv[2]], sys.argv[3])\n"
        "print(json.dumps({'n_v': int(g.n_v), 'n_e': int(g.n_e), 'count': int(sum(g.edges[i].count for i in range(g.n_e))),"
        " 'len': int(sum(g.edges[i].seq_len for i in range(g.n_e)))}))\n"
    )
    outs = []
    for host in ("0", "1"):
        env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        env.pop("TAGPU_DEVICE_PARSE", None)
        if host == "0":
            env["TAGPU_DEVICE_PARSE"] = "1"
        r = subprocess.run([sys.executable, "-c", code, str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq"), str(tmp_path)],
                           capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append((json.loads(r.stdout.strip().split("\n")[-1]), r.stderr))
    assert outs[0][0] == outs[1][0]
    assert "records parsed on the GPU" in outs[0][1] and "records parsed on the GPU" not in outs[1][1]
    assert outs[0][0]["n_e"] > 0
