"""build_local_assembly_graph (/root/reference/src/kmer_build.c:991-1044): the same stage re-entered per
gap with the two flanking edges of the global graph forced in (add_garbage) and their coverage imposed on the edges they
touch (assign_count_garbage).  Golden vectors: tests/golden/golden_local.json, produced by the UNMODIFIED reference function
through oracle/local_ref_main.c (tests/golden/make_golden.py:make_local_golden)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import _oracle
import _reads
from _cases import LOCAL_CASES, local_case

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden_local.json")))
TA_LOCAL_GPU = os.path.join(os.path.dirname(_oracle.TA_REF), "TA_local_gpu")


def oracle_local(oracle, lc, tmp_path, tag):
    cnt = oracle.count(lc["stream"], lc["lk"] + 1)
    g = oracle.graph_local(lc["lk"], cnt["hi"], cnt["lo"], cnt["count"], lc["contigs"], lc["covs"])
    binp = str(tmp_path / f"ora_local_{tag}.bin")
    oracle.save_bin(g, binp)
    info = dict(n_kmers=g.contents.n_kmer, n_v=g.contents.n_v, n_e=g.contents.n_e, n_kp1_on_edge=g.contents.n_kp1_on_edge)
    masks = oracle.graph_masks(g)
    oracle.free_graph(g)
    return binp, info, masks, cnt


@pytest.mark.parametrize("name", sorted(LOCAL_CASES))
def test_oracle_local_matches_reference_golden(oracle, name, tmp_path):
    gold = GOLDEN[name]
    lc = local_case(oracle, name, tmp_path)
    assert (lc["e1"], lc["e2"], [len(c) for c in lc["contigs"]]) == (gold["e1"], gold["e2"], gold["contig_len"])
    assert lc["covs"] == gold["covs"]
    binp, info, _, _ = oracle_local(oracle, lc, tmp_path, name)
    for f in ("n_kmers", "n_v", "n_e", "n_kp1_on_edge"):
        assert info[f] == gold[f], f
    for mode in (0, 1):
        bad, txt = _oracle.canon_text(oracle, binp, mode)
        assert bad == 0 and hashlib.md5(txt).hexdigest() == gold[f"canon{mode}_md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LOCAL_CASES))
def test_gpu_local_graph(tagpu, oracle, name, tmp_path):
    gold = GOLDEN[name]
    lc = local_case(oracle, name, tmp_path)
    tagpu.set_cutoff(2)
    st = tagpu.build_local_host(lc["stream"], lc["lk"], lc["contigs"], lc["covs"])
    binp, info, (khi, klo, kmask), cnt = oracle_local(oracle, lc, tmp_path, name)
    assert st["n_solid"] == cnt["hi"].size and st["n_instances"] == cnt["n_instances"]
    for f in ("n_kmers", "n_v", "n_e", "n_kp1_on_edge"):
        assert st[f] == info[f] == gold[f], f
    ghi, glo, gmask = tagpu.kmers()
    o = np.lexsort((glo, ghi))
    assert np.array_equal(ghi[o], khi) and np.array_equal(glo[o], klo) and np.array_equal(gmask[o], kmask)
    gpu_bin = str(tmp_path / f"gpu_local_{name}.bin")
    tagpu.write_graph_bin(gpu_bin)
    for mode in (0, 1):
        bad_o, txt_o = _oracle.canon_text(oracle, binp, mode)
        bad_g, txt_g = _oracle.canon_text(oracle, gpu_bin, mode)
        assert bad_o == 0 and bad_g == 0 and txt_o == txt_g
        assert hashlib.md5(txt_g).hexdigest() == gold[f"canon{mode}_md5"]
    # a plain build afterwards must not see the garbage any more
    ref = tagpu.build_host(lc["stream"], lc["lk"])
    g = oracle.graph(lc["lk"], cnt["hi"], cnt["lo"], cnt["count"])
    assert (ref["n_kmers"], ref["n_v"], ref["n_e"]) == (g.contents.n_kmer, g.contents.n_v, g.contents.n_e)
    oracle.free_graph(g)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(TA_LOCAL_GPU), reason="oracle/_ref/TA_local_gpu (reference objects + libtagpu.so's build_local_assembly_graph) not built")
def test_dropin_local_assembly_binary(oracle, tmp_path):
    """The reference's own load_asm_graph / test_asm_graph / save_asm_graph around OUR build_local_assembly_graph."""
    for name in sorted(LOCAL_CASES):
        gold = GOLDEN[name]
        lc = local_case(oracle, name, tmp_path)
        f1, f2 = str(tmp_path / f"{name}_R1.fq"), str(tmp_path / f"{name}_R2.fq")
        _reads.write_fastq(f1, lc["r1"], 1)
        _reads.write_fastq(f2, lc["r2"], 2)
        binp = str(tmp_path / f"{name}_local.bin")
        p = subprocess.run([TA_LOCAL_GPU, lc["g0_bin"], str(lc["e1"]), str(lc["e2"]), str(lc["lk"]), f1, f2, str(tmp_path), binp, "4"],
                           capture_output=True, text=True)
        log = p.stdout + p.stderr
        assert p.returncode == 0, log[-3000:]
        assert f"sum_count = {gold['sum_count']}" in log
        bad, txt = _oracle.canon_text(oracle, binp, 0)
        assert bad == 0 and hashlib.md5(txt).hexdigest() == gold["canon0_md5"]


@pytest.mark.gpu
def test_gpu_local_batch(oracle, tmp_path):
    """tagpu_build_local_batch: many gaps in flight on several contexts (own stream and host thread each).  60 jobs cycling
    through the three golden cases on 6 contexts: every job must report its case's golden counters, and the graphs filled
    for the first of each case must equal the oracle's canonically (a context must never see another gap's k-mers)."""
    import ctypes as C
    from turingassembler_b200.api import build_local_batch, free_asm_graph, load_library
    names = sorted(LOCAL_CASES)
    cases = {n: local_case(oracle, n, tmp_path) for n in names}
    jobs = [dict(stream=cases[n]["stream"], k=cases[n]["lk"], contigs=cases[n]["contigs"], covs=cases[n]["covs"])
            for i in range(60) for n in [names[i % len(names)]]]
    for n_ctx in (1, 6):
        stats, graphs = build_local_batch(jobs, n_ctx=n_ctx, fill=True)
        lib = load_library()
        for i, (st, g) in enumerate(zip(stats, graphs)):
            gold = GOLDEN[names[i % len(names)]]
            for f in ("n_kmers", "n_v", "n_e", "n_kp1_on_edge"):
                assert st[f] == gold[f], (i, f)
            assert (g.n_v, g.n_e) == (gold["n_v"], gold["n_e"])
            if i < len(names):
                # the filled struct, written through the .bin writer of the oracle side: sorted unitig lines must match the golden md5
                lines = []
                for e in range(g.n_e):
                    ed = g.edges[e]
                    if e > ed.rc_id:
                        continue
                    seq = "".join("ACGT"[(ed.seq[b >> 4] >> ((b & 15) << 1)) & 3] for b in range(ed.seq_len))
                    rc = seq[::-1].translate(str.maketrans("ACGT", "TGCA"))
                    lines.append(f"{min(seq, rc)}\t{ed.count}\t{ed.seq_len}")
                assert hashlib.md5("\n".join(sorted(lines)).encode()).hexdigest() == gold["canon0_md5"]
            free_asm_graph(g)
