"""GPU parity tests (run on the B200 with -m gpu): the CUDA path, called through the C ABI of libtagpu.so, against the
CPU oracle on the same seeded inputs and against the golden vectors the unmodified reference produced.
Bar: bit-exact — solid (k+1)-mer set with counts, (k-mer, edge mask) table, and the unitig graph after canonical
sorting."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import _oracle
import _reads
from _cases import case_stream

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def sort_keys(hi, lo, *vals):
    order = np.lexsort((lo, hi))
    return (hi[order], lo[order]) + tuple(v[order] for v in vals)


def check_against_oracle(tagpu, oracle, stream, k, tmp_path, tag="x", ci=2):
    tagpu.set_cutoff(ci)
    st = tagpu.build_host(stream, k)
    want = oracle.count(stream, k + 1, ci=ci)
    # 1. solid set with counts
    hi, lo, cnt = sort_keys(*tagpu.solid())
    assert st["n_instances"] == want["n_instances"]
    assert st["n_distinct"] == want["n_distinct"]
    assert np.array_equal(hi, want["hi"]) and np.array_equal(lo, want["lo"]) and np.array_equal(cnt, want["count"])
    assert st["sum_solid"] == int(want["count"].astype(np.uint64).sum())
    # 2. k-mer table with edge masks (after a two-level graph stage the library rebuilds the full table on demand)
    g = oracle.graph(k, want["hi"], want["lo"], want["count"])
    khi, klo, kmask = oracle.graph_masks(g)
    ghi, glo, gmask = sort_keys(*tagpu.kmers())
    assert np.array_equal(ghi, khi) and np.array_equal(glo, klo) and np.array_equal(gmask, kmask)
    # 3. graph, canonically
    ora_bin, gpu_bin = str(tmp_path / f"ora_{tag}.bin"), str(tmp_path / f"gpu_{tag}.bin")
    oracle.save_bin(g, ora_bin)
    tagpu.write_graph_bin(gpu_bin)
    assert (st["n_kmers"], st["n_v"], st["n_e"], st["n_kp1_on_edge"]) == (
        g.contents.n_kmer, g.contents.n_v, g.contents.n_e, g.contents.n_kp1_on_edge)
    oracle.free_graph(g)
    for mode in (0, 1):
        bad_o, txt_o = _oracle.canon_text(oracle, ora_bin, mode)
        bad_g, txt_g = _oracle.canon_text(oracle, gpu_bin, mode)
        assert bad_o == 0 and bad_g == 0
        assert txt_o == txt_g
    # 4. the order-independent digests bench.py prints (computed on the device) equal the oracle's
    import _digest
    dg = tagpu.digest()
    want_s = _digest.solid_digest(want["hi"], want["lo"], want["count"])
    assert dg["solid_complete"] and all(dg[f] == want_s[f] for f in want_s)
    want_e = oracle.bin_digest(ora_bin)
    assert all(dg[f] == want_e[f] for f in want_e)
    return st, gpu_bin


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_golden_cases(tagpu, oracle, key, tmp_path):
    gold = GOLDEN[key]
    st, gpu_bin = check_against_oracle(tagpu, oracle, case_stream(gold["case"]), gold["k"], tmp_path, key)
    for f in ("n_kmers", "n_v", "n_e", "n_kp1_on_edge", "n_solid", "n_instances", "n_distinct"):
        assert st[f] == gold[f], f
    for mode in (0, 1):
        assert hashlib.md5(_oracle.canon_text(oracle, gpu_bin, mode)[1]).hexdigest() == gold[f"canon{mode}_md5"]
    gz = os.path.join(HERE, "golden", f"{key}_canon0.txt.gz")
    if os.path.exists(gz):
        assert gzip.open(gz).read() == _oracle.canon_text(oracle, gpu_bin, 0)[1]


@pytest.mark.parametrize("k", [17, 21, 30, 31, 32, 33, 45, 62, 63])
def test_every_key_width(tagpu, oracle, k, tmp_path):
    """k = 31 / 32 straddle the 64 -> 128-bit key switch (K = k + 1 = 32 / 33); k = 63 fills all 128 bits."""
    stream = _reads.gen_stream(40000, 3000, seed=100 + k, sub_err=0.004)
    check_against_oracle(tagpu, oracle, stream, k, tmp_path, f"k{k}")


@pytest.mark.parametrize("ci", [1, 2, 3, 5])
def test_cutoffs(tagpu, oracle, ci, tmp_path):
    stream = _reads.gen_stream(20000, 1500, seed=7, sub_err=0.01)
    check_against_oracle(tagpu, oracle, stream, 31, tmp_path, f"ci{ci}", ci=ci)
    tagpu.set_cutoff(2)


def test_edge_cases(tagpu, oracle, tmp_path):
    for name, stream in {
        "empty": b"",
        "short": b"ACGTACGT\n",
        "only_n": b"N" * 500 + b"\n",
        "one_window": b"ACGTTGCATGCATGCAAGCTTAGCTAGGATCCA\n" * 2,
        "no_trailing_newline": b"ACGTTGCATGCATGCAAGCTTAGCTAGGATCCAGGTT" * 3,
        "lower_case": (b"acgttgcatgcatgcaagcttagctaggatccaggtt" * 3 + b"\n") * 3,
        "ragged": b"\n".join(bytes(np.random.default_rng(i).choice(list(b"ACGT"), size=n).astype(np.uint8))
                             for i, n in enumerate([1, 31, 32, 33, 64, 65, 151, 500, 8191, 8192, 8193, 20000])) * 2,
        "homopolymer": (b"A" * 300 + b"\n") * 3 + (b"T" * 200 + b"\n") * 2,
        "crlf_and_junk": b"ACGTTGCATGCATGCAAGCTTAGCTAGGATCCAGGTT\r\n" * 4 + b"@#!!\n" + b"ACGTTGCATGCATGCAAGCTTAGCTAGGATCCAGGTT\n",
    }.items():
        check_against_oracle(tagpu, oracle, stream, 31, tmp_path, name)
        check_against_oracle(tagpu, oracle, stream, 45, tmp_path, name)


def test_hairpins_palindromes_and_tandem_repeats(tagpu, oracle, tmp_path):
    """Adversarial topology: reverse-complement palindromes (even (k+1)-mers are their own rc), hairpins where a k-mer
    is followed by its own reverse complement, tandem repeats (self-loops) and a clean cycle next to a branching one."""
    rng = np.random.default_rng(42)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    def rnd(n):
        return bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8))
    def rc(s):
        return s.translate(comp)[::-1]
    parts = []
    a = rnd(200)
    parts += [a + rc(a)] * 4                      # perfect palindrome of length 400
    b = rnd(90)
    parts += [rnd(50) + b + rc(b) + rnd(50)] * 3  # embedded hairpin
    unit = rnd(37)
    parts += [rnd(60) + unit * 9 + rnd(60)] * 3   # tandem repeat
    c = rnd(300)
    parts += [(c + c)[:520]] * 3                  # circular, no branch: vanishes
    d = rnd(250)
    parts += [d + rnd(80), rnd(80) + d, d + rnd(70), rnd(90) + d] * 2
    stream = b"\n".join(parts) + b"\n"
    for k in (21, 31, 45):
        check_against_oracle(tagpu, oracle, stream, k, tmp_path, f"adv{k}")


def test_device_pointer_path_and_skip_counts(tagpu, oracle, tmp_path):
    import torch
    stream = _reads.gen_stream(50000, 4000, seed=9)
    d = torch.from_numpy(stream.copy()).cuda()
    tagpu.set_stream(torch.cuda.current_stream().cuda_stream)
    st = tagpu.build_device(d.data_ptr(), d.numel(), 31)
    ref = tagpu.build_host(stream, 31)
    for f in ("n_instances", "n_solid", "n_kmers", "n_v", "n_e", "n_kp1_on_edge", "sum_solid"):
        assert st[f] == ref[f]
    tagpu.set_skip_counts(True)
    tagpu.build_device(d.data_ptr(), d.numel(), 31)
    g = tagpu.graph()
    assert g["n_e"] == ref["n_e"] and int(g["e_count"].sum()) == 0
    tagpu.set_skip_counts(False)
    tagpu.set_stream(0)


def test_reference_entry_points_on_files(oracle, tmp_path):
    """build_graph_from_scratch / KMC_build_kmer_database through the C ABI, on FASTQ files, like the reference calls them."""
    import ctypes as C
    from turingassembler_b200 import build_graph_from_scratch, kmc_build_kmer_database
    gold = GOLDEN["P1_k31"]
    from _cases import CASES, reads_for
    r1, r2 = reads_for(*CASES["P1"][:2])
    f1, f2 = str(tmp_path / "R1.fq"), str(tmp_path / "R2.fq")
    _reads.write_fastq(f1, r1, 1)
    _reads.write_fastq(f2, r2, 2)
    g = build_graph_from_scratch(31, 4, 32, [f1], [f2], str(tmp_path))
    assert (g.ksize, g.n_v, g.n_e, g.aux_flag, g.bin_size) == (31, gold["n_v"], gold["n_e"], 0, 0)
    sum_count = 0
    for e in range(g.n_e):
        ed = g.edges[e]
        assert g.edges[ed.rc_id].rc_id == e and ed.n_holes == 0 and not ed.p_holes and not ed.barcodes
        assert g.nodes[ed.source].rc_id == g.edges[ed.rc_id].target
        if e <= ed.rc_id:
            sum_count += ed.count
    assert sum_count == gold["sum_count"]
    for u in range(g.n_v):
        assert [g.edges[g.nodes[u].adj[j]].source for j in range(g.nodes[u].deg)] == [u] * g.nodes[u].deg
    # level-1 boundary: the KMC database the reference's reader would parse
    assert kmc_build_kmer_database(32, str(tmp_path), 4, 32, [f1, f2]) == 0
    want = oracle.count(_reads.stream_of(r1, r2), 32)
    suf = open(tmp_path / "KMC_32_count.kmc_suf", "rb").read()
    pre = open(tmp_path / "KMC_32_count.kmc_pre", "rb").read()
    assert suf[:4] == b"KMCS" and suf[-4:] == b"KMCS" and pre[:4] == b"KMCP" and pre[-4:] == b"KMCP"
    rec = np.frombuffer(suf[4:-4], dtype=np.uint8).reshape(-1, 7 + 4)     # (32 - 4) / 4 suffix bytes + 4 counter bytes
    assert rec.shape[0] == gold["n_solid"]
    lut = np.frombuffer(pre[4:4 + 8 * 257], dtype=np.uint64)
    prefix = np.repeat(np.arange(256, dtype=np.uint64), np.diff(lut).astype(np.int64))
    sfx = np.zeros(rec.shape[0], np.uint64)
    for j in range(7):
        sfx = (sfx << np.uint64(8)) | rec[:, j].astype(np.uint64)
    keys = (prefix << np.uint64(56)) | sfx
    assert np.array_equal(keys, want["lo"])
    assert np.array_equal(rec[:, 7:].copy().view(np.uint32).reshape(-1), want["count"])
    del C


@pytest.mark.skipif(not os.path.exists(_oracle.TA_GPU), reason="oracle/_ref/TA_gpu (reference linked against libtagpu.so) not built")
def test_dropin_reference_binary(oracle, tmp_path):
    """The unmodified reference objects, linked against libtagpu.so in place of their own build_initial_graph /
    libkmc.a, run `build_0`: the reference's own test_asm_graph validates our graph and its save_asm_graph writes it."""
    import subprocess
    from _cases import CASES, reads_for
    r1, r2 = reads_for(*CASES["P1"][:2])
    f1, f2 = str(tmp_path / "R1.fq"), str(tmp_path / "R2.fq")
    _reads.write_fastq(f1, r1, 1)
    _reads.write_fastq(f2, r2, 2)
    for k in (31, 45):
        out = tmp_path / f"o{k}"
        out.mkdir()
        p = subprocess.run([_oracle.TA_GPU, "build_0", "-1", f1, "-2", f2, "-l", "ust", "-k0", str(k), "-t", "4", "-o", str(out)],
                           capture_output=True, text=True)
        log = p.stdout + p.stderr
        assert p.returncode == 0, log[-3000:]
        gold = GOLDEN[f"P1_k{k}"]
        assert f"sum_count = {gold['sum_count']}" in log
        bad, txt = _oracle.canon_text(oracle, str(out / f"graph_k_{k}_level_0.bin"), 0)
        assert bad == 0 and hashlib.md5(txt).hexdigest() == gold["canon0_md5"]


_CAPACITY_CODE = (
    "import sys, numpy as np\n"
    "sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})\n"
    "import _oracle, _reads\n"
    "from turingassembler_b200 import Tagpu\n"
    "t = Tagpu(); ora = _oracle.load()\n"
    "for k, seed in ((31, 3), (45, 4)):\n"
    "    s = _reads.gen_stream(120000, 12000, seed=seed)\n"
    "    for rep in range(2):\n"                       # the second build reuses the buffers the first one had to grow
    "        st = t.build_host(s, k); want = ora.count(s, k + 1)\n"
    "        hi, lo, cnt = t.solid(); o = np.lexsort((lo, hi))\n"
    "        assert st['n_instances'] == want['n_instances'] and st['n_distinct'] == want['n_distinct']\n"
    "        assert np.array_equal(hi[o], want['hi']) and np.array_equal(lo[o], want['lo']) and np.array_equal(cnt[o], want['count'])\n"
    "        g = ora.graph(k, want['hi'], want['lo'], want['count'])\n"
    "        assert (st['n_kmers'], st['n_v'], st['n_e']) == (g.contents.n_kmer, g.contents.n_v, g.contents.n_e)\n"
    "print('CAPACITY-OK')\n")


@pytest.mark.parametrize("env", [{"TAGPU_REGION_CAP": "4"}, {"TAGPU_REGION_CAP": "4", "TAGPU_OVERFLOW_CAP": "1000"}, {"TAGPU_SOLID_CAP": "500"}],
                         ids=["overflow_list", "overflow_list_regrown", "solid_buffer_regrown"])
def test_capacity_paths(env):
    """The count stage sizes its buffers from estimates and recovers when an input does not fit them.
    TAGPU_REGION_CAP=4 leaves room for four records per bucket region, so nearly every super-k-mer record takes the overflow
    route (overflow list -> histogram -> scan -> scatter -> read back through ext_off in pass 2).  With TAGPU_OVERFLOW_CAP=1000
    that list is too small as well: pass 1 is repeated with a list sized from what the first attempt counted.
    TAGPU_SOLID_CAP=500 makes the solid (k+1)-mer buffers too small: pass 2 is repeated with larger ones.  Same results as the
    oracle in every case.  Subprocess: the knobs are read once per process."""
    import subprocess
    import sys
    code = _CAPACITY_CODE.format(root=os.path.dirname(HERE), here=HERE)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env), timeout=600)
    assert p.returncode == 0 and "CAPACITY-OK" in p.stdout, (p.stdout + p.stderr)[-3000:]


def test_skewed_low_complexity_reads(tagpu, oracle, tmp_path):
    """Inputs whose minimizers pile up in few buckets (SURVEY.md §8d asks for repeat families; real reads add poly-A tails and
    simple repeats): a genome of diverged copies of one 3 kbp unit with poly-A / dinucleotide / short tandem stretches, read
    at 60x.  A handful of minimizer sites then receive a large share of the windows — the overflow list, the oversized-group
    sub-classes of pass 2 and the uncontractible blocks of the graph stage all get exercised — and the result must still
    equal the oracle's."""
    rng = np.random.default_rng(77)
    unit = rng.integers(0, 4, 3000, dtype=np.uint8)
    parts = []
    for c in range(40):                                   # a 120 kbp repeat family at ~5 % divergence
        u = unit.copy()
        m = rng.random(u.size) < 0.05
        u[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
        parts.append(u)
        parts.append(rng.integers(0, 4, 500, dtype=np.uint8))
    for _ in range(30):                                   # low-complexity stretches: poly-A, (AC)n, (AAT)n, 200-400 bp each
        n = int(rng.integers(200, 400))
        parts.append(np.zeros(n, np.uint8))
        parts.append(rng.integers(0, 4, 300, dtype=np.uint8))
        parts.append(np.tile(np.array([0, 1], np.uint8), n // 2))
        parts.append(rng.integers(0, 4, 300, dtype=np.uint8))
        parts.append(np.tile(np.array([0, 0, 3], np.uint8), n // 3))
        parts.append(rng.integers(0, 4, 300, dtype=np.uint8))
    genome = np.concatenate(parts)
    stream = _reads.gen_stream(genome.size, int(genome.size * 60 / 302), seed=78, genome=genome)
    for k in (21, 31, 45):
        check_against_oracle(tagpu, oracle, stream, k, tmp_path, tag=f"skew{k}")


@pytest.mark.skipif(not os.path.exists(_oracle.TA_KMC), reason="oracle/_ref/TA_kmc (reference + libtagpu.so in place of libkmc.a) not built")
def test_dropin_library_boundary(oracle, tmp_path):
    """INTEGRATION.md option B: every reference object unmodified, libtagpu.so only supplies KMC_build_kmer_database.
    The GPU writes KMC_<k+1>_count.kmc_pre/.kmc_suf; the reference's own KMC_read_prefix / KMC_retrieve_kmer_multi parse
    them and its own kmhash / build_asm_graph_from_kmhash / assign_count_kedge_multi build the graph from them — which
    must be the graph of the golden vectors (rows a1-a3 of SURVEY.md §8)."""
    import subprocess
    from _cases import CASES, reads_for
    r1, r2 = reads_for(*CASES["P1"][:2])
    f1, f2 = str(tmp_path / "R1.fq"), str(tmp_path / "R2.fq")
    _reads.write_fastq(f1, r1, 1)
    _reads.write_fastq(f2, r2, 2)
    for k in (31, 45):
        out = tmp_path / f"o{k}"
        out.mkdir()
        p = subprocess.run([_oracle.TA_KMC, "build_0", "-1", f1, "-2", f2, "-l", "ust", "-k0", str(k), "-t", "4", "-o", str(out)],
                           capture_output=True, text=True)
        log = p.stdout + p.stderr
        assert p.returncode == 0, log[-3000:]
        gold = GOLDEN[f"P1_k{k}"]
        assert f"Number of kmer: {gold['n_kmers']}" in log and f"sum_count = {gold['sum_count']}" in log
        assert (out / f"KMC_{k + 1}_count.kmc_suf").exists()
        bad, txt = _oracle.canon_text(oracle, str(out / f"graph_k_{k}_level_0.bin"), 0)
        assert bad == 0 and hashlib.md5(txt).hexdigest() == gold["canon0_md5"]


def test_list_ranking_variant():
    """The work-efficient list ranking (Helman-JaJa; default above 48 M chain vertices) must give the same graphs as pointer
    jumping: the golden, adversarial-topology and edge-case tests again in a subprocess with TAGPU_LIST_RANKING=hj."""
    import subprocess
    import sys
    env = dict(os.environ, TAGPU_LIST_RANKING="hj")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_parity.py"), "-q", "-m", "gpu", "-x",
                        "-k", "golden or hairpins or edge_cases or every_key_width or one_level"], capture_output=True, text=True, env=env, timeout=1200)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]


@pytest.mark.parametrize("contract", [True, False], ids=["two_level", "one_level"])
@pytest.mark.parametrize("k", [21, 31, 32, 45, 63])
def test_graph_stage_variants(tagpu, oracle, k, contract, tmp_path):
    """Both graph stages give the reference's graph, bit-exact, on golden cases, adversarial topologies and mixed inputs:
    the two-level one (default; csrc/tagpu_contract.cuh: paths contracted inside the bucket groups, then the path-driven
    global stage) and the one-level one (every k-mer in the HBM table; used for local assembly and multi-GPU builds)."""
    was = tagpu.contract
    tagpu.set_contract(contract)
    try:
        for tag, stream in (("p1", case_stream("P1")), ("m1", case_stream("M1")), ("rnd", _reads.gen_stream(60000, 5000, seed=200 + k, sub_err=0.004))):
            check_against_oracle(tagpu, oracle, stream, k, tmp_path, f"ct_{tag}_{k}")
        if k % 2 == 0:
            return          # palindromic k-mers: the reference's own rc-link assert fires on such inputs (kmer_build.c:640)
        rng = np.random.default_rng(42)
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        rnd = lambda n: bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8))
        rc = lambda s: s.translate(comp)[::-1]
        a, b, unit, c, d = rnd(200), rnd(90), rnd(37), rnd(300), rnd(250)
        parts = [a + rc(a)] * 4 + [rnd(50) + b + rc(b) + rnd(50)] * 3 + [rnd(60) + unit * 9 + rnd(60)] * 3 + [(c + c)[:520]] * 3
        parts += [d + rnd(80), rnd(80) + d, d + rnd(70), rnd(90) + d] * 2 + [b"A" * 300] * 3 + [b"ACGT" * 60] * 3
        check_against_oracle(tagpu, oracle, b"\n".join(parts) + b"\n", k, tmp_path, f"ct_adv_{k}")
    finally:
        tagpu.set_contract(was)
