"""Coverage recount: kmer_count_on_edges + add_cnt_to_graph of the reference's `build_coverage`
sub-command (/root/reference/src/coverage/kmer_count.c:198-240,113-135; caller /root/reference/src/process.c:823-835).

CPU: the oracle restatement (oracle/cov_oracle.c) on the oracle's own level-0 graph must reproduce the golden vectors that
tests/golden/make_golden.py took from the UNMODIFIED reference (sorted `sequence<TAB>count` lines, so edge numbering does not
matter) — and, where the compiled reference is present, the per-edge counts of a live `TA_ref build_coverage` run on the
reference's own .bin.  GPU: the CUDA path against the oracle and the same golden vectors, through the native call on the
device-resident graph, on flat arrays, and through the reference-named entry points on FASTQ files."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import _oracle
import _reads
from _cases import CASES, case_stream, reads_for

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden_coverage.json")))


def _lines(flat, counts):
    from make_golden import coverage_lines
    return coverage_lines(flat, counts)


def _oracle_flat(oracle, stream, k, tmp_path):
    cnt = oracle.count(stream, k + 1)
    g = oracle.graph(k, cnt["hi"], cnt["lo"], cnt["count"])
    binp = str(tmp_path / f"cov_{k}.bin")
    oracle.save_bin(g, binp)
    oracle.free_graph(g)
    return _oracle.load_bin_flat(binp)


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_oracle_matches_reference_golden(oracle, key, tmp_path):
    gold = GOLDEN[key]
    stream = case_stream(gold["case"])
    flat = _oracle_flat(oracle, stream, gold["k"], tmp_path)
    counts = oracle.coverage_recount(stream, flat["e_len"], flat["e_off"], flat["e_seq"], flat["e_rc"])
    assert flat["n_e"] == gold["n_e"] and int(counts.sum(dtype=np.uint64)) == gold["count_sum"]
    assert hashlib.md5(_lines(flat, counts)).hexdigest() == gold["lines_md5"]


@pytest.mark.skipif(not os.path.exists(_oracle.TA_REF), reason="oracle/_ref/TA_ref not built (needs /root/reference)")
def test_oracle_vs_live_reference(oracle, tmp_path):
    """per-edge equality on the reference's own numbering: its level-0 .bin in, its coverage_built .bin as the answer"""
    kind, kw, _ = CASES["M1"]
    r1, r2 = reads_for(kind, kw)
    f1, f2 = str(tmp_path / "R1.fq"), str(tmp_path / "R2.fq")
    _reads.write_fastq(f1, r1, 1)
    _reads.write_fastq(f2, r2, 2)
    o0, oc = tmp_path / "l0", tmp_path / "cov"
    o0.mkdir()
    oc.mkdir()
    for cmd in ([_oracle.TA_REF, "build_0", "-1", f1, "-2", f2, "-l", "ust", "-k0", "31", "-t", "4", "-o", str(o0)],
                [_oracle.TA_REF, "build_coverage", "-i", str(o0 / "graph_k_31_level_0.bin"), "-1", f1, "-2", f2, "-l", "ust", "-t", "4", "-o", str(oc)]):
        p = subprocess.run(cmd, capture_output=True, text=True)
        assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    g0 = _oracle.load_bin_flat(str(o0 / "graph_k_31_level_0.bin"))
    want = _oracle.load_bin_flat(str(oc / "graph_k_31_coverage_built.bin"))["e_count"]
    got = oracle.coverage_recount(_reads.stream_of(r1, r2), g0["e_len"], g0["e_off"], g0["e_seq"], g0["e_rc"])
    assert np.array_equal(got, want)


def test_oracle_quirks(oracle):
    """the two visible properties of the reference's arithmetic: an N does not break the window (it reads as A and turns the
    base before it A->C / G->T), and `rev` is the reversed string with C and G swapped, not the reverse complement"""
    rng = np.random.default_rng(3)
    edge = "".join("ACGT"[i] for i in rng.integers(0, 4, 40))
    codes = np.array(["ACGT".index(c) for c in edge], np.uint32)
    words = np.zeros(3, np.uint32)
    for i, c in enumerate(codes):
        words[i >> 4] |= np.uint32(int(c) << ((i & 15) << 1))
    one = dict(e_len=np.array([40], np.uint32), e_off=np.array([0], np.uint64), e_seq=words, e_rc=np.array([0], np.int64))
    run = lambda reads: int(oracle.coverage_recount(("\n".join(reads) + "\n").encode(), **one)[0])
    assert run([edge]) == 10                                        # 10 windows of the read, each found once
    swap = {"A": "A", "C": "G", "G": "C", "T": "T"}
    assert run(["".join(swap[c] for c in reversed(edge))]) == 10    # found through `rev`
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    assert run(["".join(comp[c] for c in reversed(edge))]) < 10     # the true reverse complement is (mostly) NOT found
    assert run([edge[:31]]) == 0                                    # a read of exactly 31 bases is skipped (len > 31 required)
    # N at position p reads as A and ORs 1 into base p - 1: craft the read so that the corrupted string equals the edge
    p = next(i for i in range(5, 35) if edge[i] == "A" and edge[i - 1] in "CT")
    prev = {"C": "A", "T": "G"}[edge[p - 1]]
    crafted = edge[:p - 1] + prev + "N" + edge[p + 1:]
    assert run([crafted]) == 10 and run([edge[:p - 1] + prev + "A" + edge[p + 1:]]) < 10


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_gpu_matches_golden_and_oracle(tagpu, oracle, key):
    gold = GOLDEN[key]
    stream = case_stream(gold["case"])
    tagpu.set_cutoff(2)
    st = tagpu.build_host(stream, gold["k"])
    g = tagpu.graph()
    flat = dict(n_e=g["n_e"], e_len=g["e_len"], e_off=g["e_off"], e_seq=g["e_seq"], e_rc=g["e_rc"].astype(np.int64))
    want = oracle.coverage_recount(stream, flat["e_len"], flat["e_off"], flat["e_seq"], flat["e_rc"])
    on_device = tagpu.coverage_recount(stream)                                  # edges still on the device
    assert np.array_equal(on_device, want)
    from_flat = tagpu.coverage_recount(stream, dict(flat, e_rc=g["e_rc"]))      # edges uploaded as flat arrays
    assert np.array_equal(from_flat, want)
    assert st["n_e"] == gold["n_e"] and int(want.sum(dtype=np.uint64)) == gold["count_sum"]
    assert hashlib.md5(_lines(flat, on_device)).hexdigest() == gold["lines_md5"]


@pytest.mark.gpu
def test_gpu_reference_entry_points(oracle, tmp_path):
    """kmer_count_on_edges(opt, g) + add_cnt_to_graph(g, table) exactly as build_coverage_process calls them, on FASTQ files
    and a struct asm_graph_t filled by build_graph_from_scratch; reads with N and lower case included."""
    from turingassembler_b200.api import AsmGraph, build_graph_from_scratch, free_asm_graph, load_library
    kind, kw, _ = CASES["P1"]
    r1, r2 = reads_for(kind, kw)
    r1 = [r.lower() if i % 7 == 0 else r for i, r in enumerate(r1)]
    f1, f2 = str(tmp_path / "R1.fq"), str(tmp_path / "R2.fq")
    _reads.write_fastq(f1, r1, 1)
    _reads.write_fastq(f2, r2, 2)
    lib = load_library()

    class Opt(C.Structure):                     # struct opt_proc_t, /root/reference/src/attribute.h:49-71 (include/tagpu_graph.h)
        _fields_ = [("n_threads", C.c_int), ("hash_size", C.c_int), ("k0", C.c_int), ("k1", C.c_int), ("k2", C.c_int),
                    ("split_len", C.c_int), ("lib_type", C.c_int), ("n_files", C.c_int),
                    ("files_1", C.POINTER(C.c_char_p)), ("files_2", C.POINTER(C.c_char_p)), ("files_I", C.c_void_p), ("var", C.c_void_p),
                    ("metagenomics", C.c_int), ("out_dir", C.c_char_p), ("in_file", C.c_char_p)]
    g = build_graph_from_scratch(31, 4, 32, [f1], [f2], str(tmp_path))
    opt = Opt()
    opt.n_threads, opt.n_files = 4, 1
    a1, a2 = (C.c_char_p * 1)(f1.encode()), (C.c_char_p * 1)(f2.encode())
    opt.files_1, opt.files_2 = a1, a2
    lib.kmer_count_on_edges.restype = C.c_void_p
    lib.kmer_count_on_edges.argtypes = [C.c_void_p, C.POINTER(AsmGraph)]
    lib.add_cnt_to_graph.argtypes = [C.POINTER(AsmGraph), C.c_void_p]
    table = lib.kmer_count_on_edges(C.byref(opt), C.byref(g))
    lib.add_cnt_to_graph(C.byref(g), table)
    # the same through the oracle on the struct's own edges
    n_e = g.n_e
    e_len = np.array([g.edges[e].seq_len for e in range(n_e)], np.uint32)
    nw = (e_len.astype(np.int64) + 15) >> 4
    e_off = (np.cumsum(nw) - nw).astype(np.uint64)
    e_seq = np.concatenate([np.ctypeslib.as_array(g.edges[e].seq, shape=(int(nw[e]),)) for e in range(n_e)]).astype(np.uint32)
    e_rc = np.array([g.edges[e].rc_id for e in range(n_e)], np.int64)
    got = np.array([g.edges[e].count for e in range(n_e)], np.uint64)
    want = oracle.coverage_recount(_reads.stream_of(r1, r2), e_len, e_off, e_seq, e_rc)
    free_asm_graph(g)
    assert np.array_equal(got, want) and int(want.sum()) > 0
