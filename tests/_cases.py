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
