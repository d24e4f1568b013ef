"""CPU tests: pin the oracle (oracle/*.c restatement) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py), and — when oracle/_ref/TA_ref is present — against the reference itself."""
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import _oracle
import _reads
from _cases import CASES, case_stream, reads_for

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def oracle_build(oracle, stream, k, tmp_path, tag):
    cnt = oracle.count(stream, k + 1)
    g = oracle.graph(k, cnt["hi"], cnt["lo"], cnt["count"])
    binp = str(tmp_path / f"ora_{tag}.bin")
    oracle.save_bin(g, binp)
    info = dict(n_kmers=g.contents.n_kmer, n_v=g.contents.n_v, n_e=g.contents.n_e,
                n_kp1_on_edge=g.contents.n_kp1_on_edge, n_solid=len(cnt["count"]), n_instances=cnt["n_instances"],
                n_distinct=cnt["n_distinct"], sum_solid=int(cnt["count"].astype(np.uint64).sum()))
    oracle.free_graph(g)
    return binp, info


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_oracle_matches_reference_golden(oracle, key, tmp_path):
    gold = GOLDEN[key]
    stream = case_stream(gold["case"])
    binp, info = oracle_build(oracle, stream, gold["k"], tmp_path, key)
    for f in ("n_kmers", "n_v", "n_e", "n_kp1_on_edge", "n_solid", "n_instances", "n_distinct"):
        assert info[f] == gold[f], f
    for mode in (0, 1):
        bad, txt = _oracle.canon_text(oracle, binp, mode)
        assert bad == 0
        assert hashlib.md5(txt).hexdigest() == gold[f"canon{mode}_md5"]
    gz = os.path.join(HERE, "golden", f"{key}_canon0.txt.gz")
    if os.path.exists(gz):
        assert gzip.open(gz).read() == _oracle.canon_text(oracle, binp, 0)[1]
    # identities that hold on every input
    if gold["n_e"]:
        assert gold["sum_count"] == info["sum_solid"] or gold["n_kp1_on_edge"] < gold["n_solid"]


def test_survey_known_answers(oracle):
    """The numbers recorded in SURVEY.md App. E for fixture P1 (reference log lines of the probe run)."""
    g31, g45 = GOLDEN["P1_k31"], GOLDEN["P1_k45"]
    assert (g31["n_instances"], g31["n_distinct"], g31["n_solid"], g31["sum_count"]) == (969119, 138362, 32264, 863021)
    assert (g31["n_kmers"], g31["n_v"], g31["n_e"], g31["n_kp1_on_edge"]) == (32239, 500, 550, 32264)
    assert g31["canon0_md5"] == "c79558baa5735bb7a55ac6a494a4678a"
    assert (g45["n_kmers"], g45["n_v"], g45["n_e"], g45["n_kp1_on_edge"], g45["sum_count"]) == (32399, 452, 450, 32398, 722219)
    assert g45["canon0_md5"] == "9c5c05dfb627459741d546ec8afd16b0"
    assert g31["fastq_md5"] == ["4bc56c228ad2a57e11508395871ed5c5", "3fbfe3066279b8fa05a5d076be3c1da0"]


def test_circular_genome_vanishes(oracle, tmp_path):
    """a repeat-free circular replicon has no node k-mer and yields an empty level-0 graph."""
    gold = GOLDEN["P3_circular_clean_k31"]
    assert gold["n_v"] == 0 and gold["n_e"] == 0 and gold["n_kmers"] == 5000


def test_count_edge_cases(oracle):
    # empty stream, stream shorter than K, N-only, window exactly K, lower case, K = 32 and K = 64 boundaries
    assert len(oracle.count(b"", 32)["count"]) == 0
    assert len(oracle.count(b"ACGT\n", 32)["count"]) == 0
    assert oracle.count(b"N" * 100, 32)["n_instances"] == 0
    s = b"ACGTTGCATGCATGCAAGCTTAGCTAGGATCCA"  # 33 bases
    assert oracle.count(s + b"\n" + s + b"\n", 32, ci=2)["n_instances"] == 4
    low = oracle.count(s.lower() + b"\n" + s + b"\n", 32, ci=2)
    assert list(low["count"]) == [2, 2]
    # reverse complement reads count towards the same canonical key
    rc = bytes({65: 84, 67: 71, 71: 67, 84: 65}[c] for c in reversed(s))
    both = oracle.count(s + b"\n" + rc + b"\n", 33, ci=2)
    assert list(both["count"]) == [2]
    # a break in the middle
    assert oracle.count(s[:16] + b"N" + s[16:] + b"\n", 18)["n_instances"] == 0
    long = (s * 4)[:128]
    assert oracle.count(long + b"\n", 64, ci=1)["n_instances"] == 65


def test_count_matches_numpy_bruteforce(oracle):
    rng = np.random.default_rng(0)
    stream = _reads.gen_stream(3000, 300, seed=2)
    for K in (18, 22, 32, 33, 46, 64):
        got = oracle.count(stream, K, ci=1)
        # brute force with python ints
        text = stream.tobytes().decode()
        code = {"A": 0, "C": 1, "G": 2, "T": 3}
        from collections import Counter
        c = Counter()
        for read in text.split("\n"):
            for i in range(len(read) - K + 1):
                w = read[i:i + K]
                if any(ch not in code for ch in w):
                    continue
                f = 0
                r = 0
                for j, ch in enumerate(w):
                    f = (f << 2) | code[ch]
                    r |= (3 - code[ch]) << (2 * j)
                c[min(f, r)] += 1
        keys = sorted(c)
        assert [int(h) << 64 | int(l) for h, l in zip(got["hi"], got["lo"])] == keys
        assert [int(x) for x in got["count"]] == [c[x] for x in keys]
        assert got["n_instances"] == sum(c.values())
    del rng


@pytest.mark.skipif(not os.path.exists(_oracle.TA_REF), reason="oracle/_ref/TA_ref (compiled reference) not present")
@pytest.mark.parametrize("name,k", [("P1", 31), ("P1", 45), ("M2_lowcov", 55)])
def test_oracle_vs_compiled_reference(oracle, name, k, tmp_path):
    kind, kw, _ = CASES[name]
    r1, r2 = reads_for(kind, kw)
    f1, f2 = str(tmp_path / "R1.fq"), str(tmp_path / "R2.fq")
    _reads.write_fastq(f1, r1, 1)
    _reads.write_fastq(f2, r2, 2)
    out = tmp_path / "ref"
    out.mkdir()
    p = subprocess.run([_oracle.TA_REF, "build_0", "-1", f1, "-2", f2, "-l", "ust", "-k0", str(k), "-t", "4", "-o", str(out)],
                       capture_output=True, text=True)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    ref_bin = str(out / f"graph_k_{k}_level_0.bin")
    ora_bin, _ = oracle_build(oracle, _reads.stream_of(r1, r2), k, tmp_path, "x")
    for mode in (0, 1):
        assert _oracle.canon_text(oracle, ref_bin, mode) == _oracle.canon_text(oracle, ora_bin, mode)


def test_digest_restatements_agree(oracle, tmp_path):
    """The edge digest has three statements: CUDA (csrc/tagpu_digest.cuh, checked on the GPU), C over a .bin
    (oracle/canon_dump.c) and numpy over flat arrays (tests/_digest.py).  The two CPU ones must agree, and the digest must
    not depend on edge numbering."""
    import ctypes as C
    import _digest
    stream = _reads.gen_stream(40000, 3000, seed=21)
    for k in (31, 45):
        cnt = oracle.count(stream, k + 1)
        g = oracle.graph(k, cnt["hi"], cnt["lo"], cnt["count"])
        binp = str(tmp_path / f"d{k}.bin")
        oracle.save_bin(g, binp)
        gc = g.contents
        n_e = gc.n_e
        e_len = np.ctypeslib.as_array(gc.e_len, shape=(n_e,)).copy()
        e_count = np.ctypeslib.as_array(gc.e_count, shape=(n_e,)).copy()
        e_off = np.ctypeslib.as_array(C.cast(gc.e_seq_off, C.POINTER(C.c_uint64)), shape=(n_e,)).copy()
        n_w = int((e_off + ((e_len.astype(np.uint64) + 15) >> 4)).max())
        e_seq = np.ctypeslib.as_array(C.cast(gc.e_seq, C.POINTER(C.c_uint32)), shape=(n_w,)).copy()
        oracle.free_graph(g)
        a = oracle.bin_digest(binp)
        b = _digest.edge_digest(e_len, e_count, e_off, e_seq)
        assert a == b and a["n_e"] == n_e > 0
        perm = np.random.default_rng(k).permutation(n_e)          # renumbered edges: same digest
        assert _digest.edge_digest(e_len[perm], e_count[perm], e_off[perm], e_seq) == a
        s1 = _digest.solid_digest(cnt["hi"], cnt["lo"], cnt["count"])
        p2 = np.random.default_rng(k + 1).permutation(cnt["hi"].size)
        assert _digest.solid_digest(cnt["hi"][p2], cnt["lo"][p2], cnt["count"][p2]) == s1
        half = cnt["hi"].size // 2                                  # shares of a sharded set add up
        sa = _digest.solid_digest(cnt["hi"][:half], cnt["lo"][:half], cnt["count"][:half])
        sb = _digest.solid_digest(cnt["hi"][half:], cnt["lo"][half:], cnt["count"][half:])
        assert (sa["solid_sum"] + sb["solid_sum"]) % (1 << 64) == s1["solid_sum"] and sa["solid_xor"] ^ sb["solid_xor"] == s1["solid_xor"]
