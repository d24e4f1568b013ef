"""numpy restatement of the order-independent digests libtagpu computes on the device (csrc/tagpu_digest.cuh).
Test infrastructure: checks Tagpu.digest() against the oracle's solid set and flat graphs."""
import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
C2 = np.uint64(0xC2B2AE3D27D4EB4F)


def mix64(x):
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def solid_digest(hi, lo, count):
    """-> dict(solid_sum, solid_xor, solid_n) of a solid (k+1)-mer set given as (hi, lo, count) arrays, any order."""
    hi, lo = np.asarray(hi, np.uint64), np.asarray(lo, np.uint64)
    with np.errstate(over="ignore"):
        d = mix64(lo ^ mix64(hi + GOLD * (np.asarray(count, np.uint64) + np.uint64(1))))
        return {"solid_sum": int(d.sum(dtype=np.uint64)), "solid_xor": int(np.bitwise_xor.reduce(d)) if d.size else 0,
                "solid_n": int(d.size)}


def edge_digest(e_len, e_count, e_off, e_seq):
    """-> dict(edge_sum, edge_xor, edge_len_sum, edge_count_sum, n_e) of a flat graph (Tagpu.graph() arrays)."""
    e_len, e_off = np.asarray(e_len, np.int64), np.asarray(e_off, np.int64)
    e_count = np.asarray(e_count, np.uint64)
    n_e = e_len.size
    if n_e == 0:
        return {"edge_sum": 0, "edge_xor": 0, "edge_len_sum": 0, "edge_count_sum": 0, "n_e": 0}
    nw = (e_len + 15) >> 4
    start = np.cumsum(nw) - nw
    owner = np.repeat(np.arange(n_e), nw)
    idx = np.arange(int(nw.sum())) - start[owner]
    words = np.asarray(e_seq, np.uint64)[e_off[owner] + idx]
    with np.errstate(over="ignore"):
        term = mix64(words ^ (GOLD * (idx.astype(np.uint64) + np.uint64(1))))
        hw = np.add.reduceat(term, start)
        d = mix64(hw ^ mix64((e_len.astype(np.uint64) << np.uint64(32)) ^ (e_count * C2)))
        return {"edge_sum": int(d.sum(dtype=np.uint64)), "edge_xor": int(np.bitwise_xor.reduce(d)),
                "edge_len_sum": int(e_len.sum()), "edge_count_sum": int(e_count.sum(dtype=np.uint64)), "n_e": int(n_e)}
