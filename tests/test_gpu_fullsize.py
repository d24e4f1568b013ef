"""Full-size (BASELINE.json configs[0]/[1]: E. coli-scale, 2 M pairs x 151 bp, k0 = 31 and 45) runs of the CUDA path, checked
through size-independent properties — the oracle is too slow for these sizes (SURVEY.md §8c "known-answer identities"):

  * #solid == "(k+1)-mer on edge" whenever no node-free cycle exists, and always sum_e (len_e - k) == 2 * n_kp1_on_edge;
  * sum of edge counts over e <= rc(e) (self-rc edges once... counted twice by the reference, App. A.7) relates to sum_solid;
  * rc_id is an involution, source/target are rc-symmetric, lengths and counts agree on an edge and its twin;
  * every edge's first k bases spell its source node and its last k bases its target node (test_asm_graph's checks);
  * a second run on the same input is bit-identical in every order-independent quantity (idempotence of the buffers), and
    a sample of the reads counted by the oracle gives counts <= the full-run counts for the same keys (monotonicity).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _seq_base(g, e, i):
    w = g["e_seq"][int(g["e_off"][e]) + (i >> 4)]
    return (int(w) >> ((i & 15) << 1)) & 3


@pytest.mark.parametrize("k", [31, 45])
def test_full_size_invariants(tagpu, oracle, k):
    import torch
    import bench
    wl = bench.WORKLOADS["C2"]
    d = bench.gen_reads_gpu(torch, wl, torch.device("cuda", 0))
    tagpu.set_cutoff(2)
    st = tagpu.build_device(d.data_ptr(), d.numel(), k)
    g = tagpu.graph()
    n_e = g["n_e"]
    # window count: 2 M pairs x 151 bp, 2 % of reads carry one N
    assert st["n_instances"] <= 4_000_000 * (151 - k) and st["n_instances"] > 0.97 * 4_000_000 * (151 - k)
    assert st["n_kp1_on_edge"] <= st["n_solid"] and st["n_solid"] - st["n_kp1_on_edge"] < 1000   # only node-free cycles are dropped
    length = g["e_len"].astype(np.int64)
    rc = g["e_rc"].astype(np.int64)
    assert np.array_equal(rc[rc], np.arange(n_e))                                   # involution
    assert np.array_equal(length[rc], length) and np.array_equal(g["e_count"][rc], g["e_count"])
    assert np.array_equal(g["e_src"][rc] ^ 1, g["e_dst"]) and np.array_equal(g["e_dst"][rc] ^ 1, g["e_src"])
    # every solid (k+1)-mer on an edge is one window of the edge and one of its twin (a self-rc edge holds both windows
    # itself; only a central palindromic (k+1)-mer of an odd self-rc edge breaks the pairing)
    assert abs(int((length - k).sum()) - 2 * st["n_kp1_on_edge"]) <= int((rc == np.arange(n_e)).sum())
    # counts: every solid (k+1)-mer on an edge adds its count to the edge and to its twin
    total = int(g["e_count"].astype(np.uint64).sum())
    hi, lo, cnt = tagpu.solid()
    assert int(cnt.astype(np.uint64).sum()) == st["sum_solid"]
    if st["n_kp1_on_edge"] == st["n_solid"]:
        assert total == 2 * st["sum_solid"]
    else:
        assert total <= 2 * st["sum_solid"]
    # node/edge consistency on a sample of edges: first k bases == source node k-mer is checked via the adjacency layout
    ebase, mask = g["node_ebase"].astype(np.int64), g["node_mask"]
    deg_f = np.array([bin(int(m) & 15).count("1") for m in range(256)])[mask]
    deg_r = np.array([bin(int(m) >> 4).count("1") for m in range(256)])[mask]
    assert int(deg_f.sum() + deg_r.sum()) == n_e
    src = g["e_src"].astype(np.int64)
    first = ebase[src >> 1] + np.where(src & 1, deg_f[src >> 1], 0)
    last = first + np.where(src & 1, deg_r[src >> 1], deg_f[src >> 1])
    e_ids = np.arange(n_e)
    assert np.all((e_ids >= first) & (e_ids < last))                                # every edge sits in its source's adj range
    rng = np.random.default_rng(0)
    for e in rng.integers(0, n_e, size=200):
        e = int(e)
        r = int(rc[e])
        L = int(length[e])
        a = [_seq_base(g, e, i) for i in range(L)]
        b = [_seq_base(g, r, i) for i in range(L)]
        assert a == [3 - x for x in reversed(b)]                                    # twin spells the reverse complement
    # second run: identical order-independent results
    st2 = tagpu.build_device(d.data_ptr(), d.numel(), k)
    for f in ("n_instances", "n_distinct", "n_solid", "sum_solid", "n_kmers", "n_v", "n_e", "n_kp1_on_edge", "n_seq_words"):
        assert st[f] == st2[f], f
    # monotonicity against the oracle on a slice of the reads: same keys, counts never larger than in the full run
    sample = d[: 20000 * 152].cpu().numpy()
    want = oracle.count(sample, k + 1, ci=2)
    order = np.lexsort((lo, hi))
    hi, lo, cnt = hi[order], lo[order], cnt[order]
    key_full = hi.astype(object) * (1 << 64) + lo.astype(object) if k + 1 > 32 else lo
    key_s = want["hi"].astype(object) * (1 << 64) + want["lo"].astype(object) if k + 1 > 32 else want["lo"]
    if k + 1 <= 32:
        pos = np.searchsorted(key_full, key_s)
        assert np.all(pos < key_full.size) and np.array_equal(key_full[pos], key_s)
        assert np.all(cnt[pos] >= want["count"])
    else:
        full = dict(zip(key_full.tolist(), cnt.tolist()))
        assert all(full.get(kk, 0) >= c for kk, c in zip(key_s.tolist(), want["count"].tolist()))


def _edge_fingerprints(g):
    """Numbering-independent multiset of the edges: (length, count, xor and wrapping sum of the 2-bit sequence words)."""
    off = g["e_off"].astype(np.int64)
    order = np.argsort(off, kind="stable")
    starts = off[order]
    assert starts[0] == 0 and np.all(np.diff(starts) > 0)
    words = g["e_seq"].astype(np.uint64)
    x = np.empty(g["n_e"], np.uint64)
    s = np.empty(g["n_e"], np.uint64)
    x[order] = np.bitwise_xor.reduceat(words, starts)
    s[order] = np.add.reduceat(words * np.uint64(0x9E3779B97F4A7C15), starts)
    keys = (s, x, g["e_count"].astype(np.uint64), g["e_len"].astype(np.uint64))
    o = np.lexsort(keys)
    return tuple(a[o] for a in keys)


@pytest.mark.parametrize("k", [31, 45])
def test_full_size_two_level_equals_one_level(tagpu, k):
    """Both graph stages (two-level with contraction inside the bucket groups, one-level with every k-mer in the HBM table)
    give the same graph on the full-size workload: same counters and the same multiset of (length, count, sequence) edges."""
    import torch
    import bench
    wl = bench.WORKLOADS["C2"]
    d = bench.gen_reads_gpu(torch, wl, torch.device("cuda", 0))
    tagpu.set_cutoff(2)
    was = tagpu.contract
    try:
        tagpu.set_contract(True)
        st2 = tagpu.build_device(d.data_ptr(), d.numel(), k)
        fp2 = _edge_fingerprints(tagpu.graph())
        tagpu.set_contract(False)
        st1 = tagpu.build_device(d.data_ptr(), d.numel(), k)
        fp1 = _edge_fingerprints(tagpu.graph())
    finally:
        tagpu.set_contract(was)
    for f in ("n_instances", "n_distinct", "n_solid", "sum_solid", "n_kmers", "n_v", "n_e", "n_kp1_on_edge", "n_seq_words"):
        assert st1[f] == st2[f], f
    for a, b in zip(fp1, fp2):
        assert np.array_equal(a, b)


def _ref_log_counters(text):
    """The reference's known-answer log lines (kmer_build.c:758,763,772; assembly_graph.c:1003)."""
    import re
    out = {}
    for name, pat in (("n_kmers", r"Number of kmer: (\d+)"), ("n_v", r"kmer_build\.c:763.*Number of nodes: (\d+)"),
                      ("n_e", r"kmer_build\.c:763.*Number of edges: (\d+)"), ("n_kp1_on_edge", r"Number of \(k\+1\)-mer on edge: (\d+)"),
                      ("sum_count", r"sum_count = (\d+)")):
        m = re.search(pat, text)
        assert m, f"reference log lacks {name}"
        out[name] = int(m.group(1))
    return out


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_full_size_parity_with_reference(tagpu, oracle, name):
    """BASELINE.json configs[0] / configs[1] at FULL size, bit-exact against the unmodified reference: the bench generator's
    read set is written as FASTQ, `oracle/_ref/TA_ref build_0 -t $(nproc)` builds graph_k_<k>_level_0.bin from it
    (/root/reference/src/kmer_build.c:714-786 behind src/process.c:47), and the GPU result must give the same canonical
    dumps (unitigs and topology), the same log counters, the same solid set with counts as
    ora_count_stream, and the digest bench.py prints must equal the digest of the reference's .bin (and the committed one)."""
    import json
    import shutil
    import subprocess
    import tempfile
    import torch
    import _digest
    import _oracle
    import bench
    assert os.path.exists(_oracle.TA_REF), "oracle/_ref/TA_ref is missing: run __graft_entry__.build() where /root/reference is mounted"
    wl = bench.WORKLOADS[name]
    k = wl["k"]
    d = bench.gen_reads_gpu(torch, wl, torch.device("cuda", 0))
    h = d.cpu().numpy()
    td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        f1, f2 = bench.write_fastq_pair(h, td)
        out = os.path.join(td, "ref")
        os.makedirs(out)
        p = subprocess.run([_oracle.TA_REF, "build_0", "-1", f1, "-2", f2, "-l", "ust", "-k0", str(k), "-t", str(os.cpu_count() or 4), "-o", out],
                           capture_output=True, text=True)
        assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
        ref = _ref_log_counters(p.stdout + p.stderr)
        ref_bin = os.path.join(out, f"graph_k_{k}_level_0.bin")
        tagpu.set_cutoff(2)
        st = tagpu.build_device(d.data_ptr(), d.numel(), k)
        assert (st["n_kmers"], st["n_v"], st["n_e"], st["n_kp1_on_edge"]) == (ref["n_kmers"], ref["n_v"], ref["n_e"], ref["n_kp1_on_edge"])
        g = tagpu.graph()
        rc = g["e_rc"].astype(np.int64)
        assert int(g["e_count"][np.arange(g["n_e"]) <= rc].sum(dtype=np.uint64)) == ref["sum_count"]
        gpu_bin = os.path.join(td, "gpu.bin")
        tagpu.write_graph_bin(gpu_bin)
        for mode in (0, 1):
            bad_r, txt_r = _oracle.canon_text(oracle, ref_bin, mode)
            bad_g, txt_g = _oracle.canon_text(oracle, gpu_bin, mode)
            assert bad_r == 0 and bad_g == 0
            assert txt_r == txt_g, f"canonical dump (mode {mode}) differs from the reference's"
            del txt_r, txt_g
        # solid set with counts
        want = oracle.count(h, k + 1, ci=2, threads=os.cpu_count() or 4)
        hi, lo, cnt = tagpu.solid()
        o = np.lexsort((lo, hi))
        assert st["n_instances"] == want["n_instances"] and st["n_distinct"] == want["n_distinct"]
        assert np.array_equal(hi[o], want["hi"]) and np.array_equal(lo[o], want["lo"]) and np.array_equal(cnt[o], want["count"])
        # digests: device == reference's .bin == oracle's solid set == committed golden
        dg = tagpu.digest()
        ref_dg = oracle.bin_digest(ref_bin)
        assert all(dg[f] == ref_dg[f] for f in ref_dg), (dg, ref_dg)
        want_s = _digest.solid_digest(want["hi"], want["lo"], want["count"])
        assert all(dg[f] == want_s[f] for f in want_s)
        gold_path = os.path.join(ROOT, "tests", "golden", "digest_fullsize.json")
        gold = json.load(open(gold_path)).get(name) if os.path.exists(gold_path) else None   # (a workload without an entry is only compared with the reference)
        print(f"DIGEST {name}: " + json.dumps({f: dg[f] for f in sorted(dg)}))
        if gold:
            assert all(dg[f] == gold[f] for f in gold), (dg, gold)
    finally:
        shutil.rmtree(td, ignore_errors=True)
