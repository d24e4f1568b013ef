"""Seeded synthetic read generators shared by tests, golden-vector scripts and bench.py."""
import random

import numpy as np

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def _rc(s):
    return "".join(_COMP[c] for c in reversed(s))


def p1_pairs(genome_len=30000, seed=7, circular=False, err=0.005, n_rate=0.02, cov=40, repeat=True):
    """SURVEY.md App. E fixture P1 generator (python `random`, exact byte-for-byte)."""
    random.seed(seed)
    G = "".join(random.choice("ACGT") for _ in range(genome_len))
    if repeat:
        G = G[:10000] + G[2000:2600] + G[10000:]
    L = 151
    n = int(len(G) * cov / (2 * L))
    r1s, r2s = [], []
    for i in range(n):
        ins = random.randint(300, 500)
        p = random.randint(0, len(G) - ins)
        frag = G[p:p + ins]
        if random.random() < 0.5:
            frag = _rc(frag)
        r1 = list(frag[:L])
        r2 = list(_rc(frag)[:L])
        for r in (r1, r2):
            for j in range(L):
                if random.random() < err:
                    r[j] = random.choice("ACGT")
            if random.random() < n_rate:
                r[random.randrange(L)] = "N"
        r1s.append("".join(r1))
        r2s.append("".join(r2))
    return r1s, r2s


def write_fastq(path, reads, mate):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            f.write("@r%d/%d\n%s\n+\n%s\n" % (i, mate, r, "I" * len(r)))


def stream_of(*read_lists):
    """files_1 ++ files_2 order, every read followed by a newline — the flat stream both oracle and GPU consume."""
    return ("\n".join(r for lst in read_lists for r in lst) + "\n").encode()


def gen_stream(genome_len, n_pairs, L=151, sub_err=0.005, n_rate=0.02, seed=1, n_repeats=8, repeat_len=600,
               circular=False, genome=None):
    """Vectorised generator for larger cases: uniform genome + planted repeats, paired reads with
    insert ~U[300,500], substitution errors, a few reads with one N.  Returns a uint8 array: R1 reads then R2 reads,
    each followed by '\\n' (so stride L+1)."""
    rng = np.random.default_rng(seed)
    if genome is None:
        g = rng.integers(0, 4, genome_len, dtype=np.uint8)
        for _ in range(n_repeats):
            if genome_len > 4 * repeat_len:
                a = int(rng.integers(0, genome_len - repeat_len))
                b = int(rng.integers(0, genome_len - repeat_len))
                g[b:b + repeat_len] = g[a:a + repeat_len]
    else:
        g = genome
    glen = g.size
    if circular:
        g = np.concatenate([g, g[:1000]])
    ins = rng.integers(300, 501, n_pairs)
    pos = (rng.random(n_pairs) * (g.size - ins)).astype(np.int64)
    flip = rng.random(n_pairs) < 0.5
    idx = np.arange(L, dtype=np.int64)[None, :]
    # mate 1 reads the fragment forward from its start, mate 2 reverse-complements from its end
    f_idx = pos[:, None] + idx
    r_idx = (pos + ins - 1)[:, None] - idx
    fwd = g[f_idx]
    rev = 3 - g[r_idx]
    r1 = np.where(flip[:, None], rev, fwd)
    r2 = np.where(flip[:, None], fwd, rev)
    out = np.empty((2 * n_pairs, L + 1), dtype=np.uint8)
    codes = np.concatenate([r1, r2], axis=0)
    errs = rng.random(codes.shape) < sub_err
    codes = np.where(errs, rng.integers(0, 4, codes.shape, dtype=np.uint8), codes)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    out[:, :L] = lut[codes]
    out[:, L] = ord("\n")
    with_n = np.nonzero(rng.random(2 * n_pairs) < n_rate)[0]
    out[with_n, rng.integers(0, L, with_n.size)] = ord("N")
    del glen
    return out.reshape(-1)
