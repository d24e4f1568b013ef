"""Packed read stream (include/tagpu.h): the host packer writes, tile by tile, exactly what the CUDA tile loader builds in
shared memory from an ASCII stream (csrc/tagpu_extract.cuh: tagpu_pack4 / tagpu_load_tile).

CPU part: tagpu_pack_stream (C, in libtagpu.so — no GPU needed) against a numpy restatement of that layout.
GPU part (-m gpu): builds from the packed stream give the same solid set, masks and graph as the oracle — i.e. as the
ASCII path — for both key widths, ragged ends and streams shorter than a tile."""
import numpy as np
import pytest

import _reads
from turingassembler_b200.api import pack_stream, packed_bytes

TILE_WORDS, TILE_BASES, TILE_BYTES = 256, 8192, 3072


def pack_numpy(stream: np.ndarray) -> np.ndarray:
    n = stream.size
    n_tiles = (n + TILE_BASES - 1) // TILE_BASES
    b = np.zeros(n_tiles * TILE_BASES, dtype=np.uint8)
    b[:n] = stream
    c = (b >> 1) & 3                       # A=0 C=1 G=3 T=2 (and whatever other bytes give: don't-care bits)
    c ^= c >> 1                            # A=0 C=1 G=2 T=3
    u = b & 0xDF
    ok = (u == ord("A")) | (u == ord("C")) | (u == ord("G")) | (u == ord("T"))
    past = np.arange(b.size) >= n
    c[past] = 0
    ok[past] = False
    c = c.reshape(-1, 32).astype(np.uint64)
    shifts = np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64)
    words = (c << shifts).sum(axis=1, dtype=np.uint64)     # disjoint bit fields: sum == or
    bad = ((~ok).reshape(-1, 32).astype(np.uint64) << (np.uint64(31) - np.arange(32, dtype=np.uint64))).sum(axis=1).astype(np.uint32)
    out = np.empty(n_tiles * TILE_BYTES, dtype=np.uint8)
    for t in range(n_tiles):
        o = t * TILE_BYTES
        out[o:o + TILE_WORDS * 8] = words[t * TILE_WORDS:(t + 1) * TILE_WORDS].view(np.uint8)
        out[o + TILE_WORDS * 8:o + TILE_BYTES] = bad[t * TILE_WORDS:(t + 1) * TILE_WORDS].view(np.uint8)
    return out


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 8191, 8192, 8193, 100_000, 3 * 8192 * 64 + 777])
def test_packer_matches_tile_layout(n):
    rng = np.random.default_rng(n)
    alphabet = np.frombuffer(b"ACGTacgtN\n\r@#0", dtype=np.uint8)
    weights = np.array([20, 20, 20, 20, 2, 2, 2, 2, 1, 2, 0.2, 0.2, 0.2, 0.2])
    stream = rng.choice(alphabet, size=n, p=weights / weights.sum()).astype(np.uint8)
    assert packed_bytes(n) == (n + TILE_BASES - 1) // TILE_BASES * TILE_BYTES
    for threads in (1, 5):
        got = pack_stream(stream, threads=threads)
        assert got.size == packed_bytes(n)
        assert np.array_equal(got, pack_numpy(stream))


def test_packer_every_byte_value():
    stream = np.tile(np.arange(256, dtype=np.uint8), 40)
    assert np.array_equal(pack_stream(stream, threads=2), pack_numpy(stream))


@pytest.mark.gpu
@pytest.mark.parametrize("k", [21, 31, 45, 63])
def test_packed_builds_match_oracle(tagpu, oracle, k, tmp_path):
    from test_gpu_parity import check_against_oracle, sort_keys
    for tag, stream in (("rnd", _reads.gen_stream(60000, 5000, seed=900 + k, sub_err=0.004)),
                        ("tiny", _reads.gen_stream(3000, 40, seed=901 + k)),
                        ("short", np.frombuffer(b"ACGTTGCAAGGCTTAACGGT" * 9 + b"\n", dtype=np.uint8))):
        stream = np.ascontiguousarray(np.frombuffer(bytes(stream), dtype=np.uint8))
        ascii_stats, _ = check_against_oracle(tagpu, oracle, stream, k, tmp_path, f"pk_a_{tag}_{k}")
        a_hi, a_lo, a_cnt = sort_keys(*tagpu.solid())
        packed = pack_stream(stream)
        st = tagpu.build_host_packed(packed, stream.size, k)
        p_hi, p_lo, p_cnt = sort_keys(*tagpu.solid())
        for key in ("n_instances", "n_distinct", "n_solid", "sum_solid", "n_kmers", "n_v", "n_e", "n_kp1_on_edge", "n_seq_words"):
            assert st[key] == ascii_stats[key], key
        assert np.array_equal(a_hi, p_hi) and np.array_equal(a_lo, p_lo) and np.array_equal(a_cnt, p_cnt)
        # and the graph itself, canonically, against the oracle's
        import _oracle
        want = oracle.count(stream, k + 1, ci=2)
        g = oracle.graph(k, want["hi"], want["lo"], want["count"])
        ora_bin, gpu_bin = str(tmp_path / f"pk_o_{tag}_{k}.bin"), str(tmp_path / f"pk_g_{tag}_{k}.bin")
        oracle.save_bin(g, ora_bin)
        tagpu.write_graph_bin(gpu_bin)
        oracle.free_graph(g)
        for mode in (0, 1):
            assert _oracle.canon_text(oracle, ora_bin, mode) == _oracle.canon_text(oracle, gpu_bin, mode)


@pytest.mark.gpu
def test_packed_large_stream_chunked_upload(tagpu, oracle):
    """More than two upload chunks (1366 tiles = 4.2 MB of packed data each): the chunked copy overlapped with pass 1,
    where the last tile of a chunk waits for its right neighbour's chunk."""
    stream = np.ascontiguousarray(np.frombuffer(bytes(_reads.gen_stream(2_000_000, 80_000, seed=77)), dtype=np.uint8))
    assert packed_bytes(stream.size) > 2 * 1366 * TILE_BYTES
    want = tagpu.build_host(stream, 45)
    got = tagpu.build_host_packed(pack_stream(stream), stream.size, 45)
    for key in ("n_instances", "n_distinct", "n_solid", "sum_solid", "n_kmers", "n_v", "n_e", "n_kp1_on_edge", "n_seq_words"):
        assert got[key] == want[key], key
