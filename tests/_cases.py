"""Golden-vector cases: seeded read sets shared by tests/golden/make_golden.py and the parity tests."""
import _reads

CASES = {
    # name: (generator, kwargs, [k...])
    "P1": ("p1", dict(genome_len=30000, seed=7), [21, 31, 45, 63]),
    "P3_circular_clean": ("p1", dict(genome_len=5000, seed=3, err=0.0, n_rate=0.0, repeat=False, circular=True), [31]),
    "M1": ("np", dict(genome_len=200000, n_pairs=20000, seed=11), [31, 45]),
    "M2_lowcov": ("np", dict(genome_len=100000, n_pairs=2000, seed=5, sub_err=0.01), [25, 55]),
}


def reads_for(kind, kw):
    if kind == "p1":
        kw = dict(kw)
        circ = kw.pop("circular", False)
        if circ:
            # SURVEY App. E fixture P3: error-free reads from G+G so the genome is effectively circular
            import random
            random.seed(kw["seed"])
            G = "".join(random.choice("ACGT") for _ in range(kw["genome_len"]))
            GG = G + G
            r1s, r2s = [], []
            for i in range(3000):
                p = random.randint(0, len(G) - 1)
                frag = GG[p:p + 400]
                if random.random() < 0.5:
                    frag = _reads._rc(frag)
                r1s.append(frag[:151])
                r2s.append(_reads._rc(frag)[:151])
            return r1s, r2s
        return _reads.p1_pairs(**kw)
    s = _reads.gen_stream(**kw).tobytes().decode().split("\n")[:-1]
    return s[: len(s) // 2], s[len(s) // 2:]


def case_stream(name, cases=None):
    """flat stream (files_1 ++ files_2 order) of a golden case"""
    kind, kw, _ = (cases or CASES)[name]
    r1, r2 = reads_for(kind, kw)
    return _reads.stream_of(r1, r2)


# ---- build_local_assembly_graph cases: global graph g0 = the oracle's level-0 graph of `case` at k0,
# flanking edges = the two longest edges with e < rc(e), local reads = the first `n_pairs` read pairs of the case, local k = lk
LOCAL_CASES = {
    "L1": dict(case="P1", k0=45, lk=31, n_pairs=1200),
    "L2": dict(case="M1", k0=45, lk=31, n_pairs=2500),
    "L3": dict(case="P1", k0=31, lk=21, n_pairs=600),
}


def local_case(oracle, name, workdir):
    """-> dict(g0_bin, e1, e2, contigs, covs, r1, r2, stream, lk)"""
    import os
    import _oracle
    spec = LOCAL_CASES[name]
    kind, kw, _ = CASES[spec["case"]]
    r1, r2 = reads_for(kind, kw)
    full = _reads.stream_of(r1, r2)
    cnt = oracle.count(full, spec["k0"] + 1)
    g = oracle.graph(spec["k0"], cnt["hi"], cnt["lo"], cnt["count"])
    g0_bin = os.path.join(str(workdir), f"g0_{name}.bin")
    oracle.save_bin(g, g0_bin)
    oracle.free_graph(g)
    g0 = _oracle.load_bin(g0_bin)
    cand = sorted((e for e, ed in enumerate(g0["edges"]) if ed and e < ed["rc"]), key=lambda e: (-g0["edges"][e]["seq_len"], e))
    e1, e2 = cand[0], cand[1]
    contigs = [g0["edges"][e]["seq"] for e in (e1, e2)]
    covs = [g0["edges"][e]["count"] * 1.0 / (g0["edges"][e]["seq_len"] - (g0["edges"][e]["n_holes"] + 1) * g0["ksize"]) for e in (e1, e2)]
    lr1, lr2 = r1[: spec["n_pairs"]], r2[: spec["n_pairs"]]
    return dict(g0_bin=g0_bin, e1=e1, e2=e2, contigs=contigs, covs=covs, r1=lr1, r2=lr2, stream=_reads.stream_of(lr1, lr2), lk=spec["lk"])


# ---- contig-file mode of build_graph_from_scratch (n_files < 0, /root/reference/src/kmer_build.c:722-731,779-781): reads of
# `case` + a contig FASTA: a novel sequence written twice (its (k+1)-mers become solid through the contig file alone and get
# no read count), two reads of the set (they lift (k+1)-mers seen once in the reads over the cutoff), and a sequence with N
CONTIG_CASES = {
    "G1": dict(case="P1", k=31),
    "G2": dict(case="P1", k=45),
    "G3": dict(case="M2_lowcov", k=25),
}


def contig_case(name):
    """-> dict(r1, r2, contigs (list of str), k)"""
    import random
    spec = CONTIG_CASES[name]
    kind, kw, _ = CASES[spec["case"]]
    r1, r2 = reads_for(kind, kw)
    rnd = random.Random(1000 + len(name) + spec["k"])
    novel = "".join(rnd.choice("ACGT") for _ in range(400))
    with_n = novel[:150] + "N" + r1[3][:120]
    return dict(r1=r1, r2=r2, contigs=[novel, novel, r1[5], r2[7], with_n, r1[5][:90] + r2[9]], k=spec["k"])


def write_fasta(path, seqs, width=70):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">contig%d\n" % i)
            for o in range(0, len(s), width):
                f.write(s[o:o + width] + "\n")
