"""Contig-file mode of the stage entry points (n_files < 0; /root/reference/src/kmer_build.c:677-679,722-731,779-781): the
graph from reads + one contig file, the edge counts from a second count pass over the reads alone.  Golden vectors:
tests/golden/golden_contig.json, written by the UNMODIFIED reference function (oracle/contig_ref_main.c -> TA_contig_ref).

CPU: the oracle's restatement of the mode (count A, count B, graph of A with B's counts) equals the golden vectors.
GPU: the reference objects linked against libtagpu.so (TA_contig_gpu) equal them too — with and without the count pass."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import _oracle
import _reads
from _cases import CONTIG_CASES, contig_case, write_fasta

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_contig.json")))


def _files(c, d):
    f1, f2, fc = (str(d / x) for x in ("R1.fq", "R2.fq", "contigs.fa"))
    _reads.write_fastq(f1, c["r1"], 1)
    _reads.write_fastq(f2, c["r2"], 2)
    write_fasta(fc, c["contigs"])
    return f1, f2, fc


@pytest.mark.parametrize("name", sorted(CONTIG_CASES))
def test_oracle_restatement_matches_reference(oracle, tmp_path, name):
    c = contig_case(name)
    gold = GOLDEN[name]
    f1, f2, fc = _files(c, tmp_path)
    a = oracle.count(oracle.load_reads([f1, f2, fc]), c["k"] + 1)
    b = oracle.count(oracle.load_reads([f1, f2]), c["k"] + 1)
    in_b = {(int(h), int(l)): int(n) for h, l, n in zip(b["hi"], b["lo"], b["count"])}
    cnt = np.array([in_b.get((int(h), int(l)), 0) for h, l in zip(a["hi"], a["lo"])], dtype=np.uint32)
    for tag, counts in (("", cnt), ("nocount_", np.zeros_like(cnt))):
        g = oracle.graph(c["k"], a["hi"], a["lo"], counts)
        assert (g.contents.n_kmer, g.contents.n_v, g.contents.n_e) == (gold[tag + "n_kmers"], gold[tag + "n_v"], gold[tag + "n_e"])
        binp = str(tmp_path / f"{tag}o.bin")
        oracle.save_bin(g, binp)
        oracle.free_graph(g)
        for mode in (0, 1):
            bad, txt = _oracle.canon_text(oracle, binp, mode)
            assert bad == 0 and hashlib.md5(txt).hexdigest() == gold[f"{tag}canon{mode}_md5"], (tag, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CONTIG_CASES))
def test_dropin_binary_in_contig_mode(oracle, tmp_path, name):
    exe = os.path.join(os.path.dirname(_oracle.TA_GPU), "TA_contig_gpu")
    assert os.path.exists(exe), "oracle/_ref/TA_contig_gpu is missing: python -c 'import __graft_entry__ as g; g.build()'"
    c = contig_case(name)
    gold = GOLDEN[name]
    f1, f2, fc = _files(c, tmp_path)
    for without in (0, 1):
        tag = "nocount_" if without else ""
        binp = str(tmp_path / f"gpu{without}.bin")
        p = subprocess.run([exe, str(c["k"]), f1, f2, fc, str(tmp_path), binp, "4", str(without)], capture_output=True, text=True, timeout=300)
        log = p.stdout + p.stderr
        assert p.returncode == 0, log[-3000:]
        assert f"Number of kmer: {gold[tag + 'n_kmers']}" in log
        assert f"Number of nodes: {gold[tag + 'n_v']}; Number of edges: {gold[tag + 'n_e']}" in log
        if not without:
            assert f"Number of (k+1)-mer on edge: {gold['n_kp1_on_edge']}" in log
        for mode in (0, 1):
            bad, txt = _oracle.canon_text(oracle, binp, mode)
            assert bad == 0 and hashlib.md5(txt).hexdigest() == gold[f"{tag}canon{mode}_md5"], (tag, mode)
