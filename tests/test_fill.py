"""Host half of the graph materialisation (tagpu_fill_asm_graph_from_flat, include/tagpu.h): flat arrays -> the reference's
struct asm_graph_t with one allocation per adj list and per edge sequence (SURVEY.md §8b ownership rules,
/root/reference/src/kmer_build.c:567-575,605-606).  Pure host code: runs without a GPU."""
import ctypes as C

import numpy as np

from turingassembler_b200.api import AsmGraph, FlatGraph, free_asm_graph, load_library


def test_fill_from_flat_roundtrip():
    lib = load_library()
    lib.tagpu_fill_asm_graph_from_flat.restype = C.c_int
    lib.tagpu_fill_asm_graph_from_flat.argtypes = [C.POINTER(FlatGraph), C.c_int, C.POINTER(AsmGraph)]
    rng = np.random.default_rng(5)
    for n_nodes in (0, 1, 7, 30000):
        mask = rng.integers(0, 256, n_nodes).astype(np.uint8)
        deg = np.array([bin(int(m)).count("1") for m in range(256)])[mask]
        ebase = (np.cumsum(deg) - deg).astype(np.uint32)
        n_e = int(deg.sum())
        e_len = rng.integers(22, 400, n_e).astype(np.uint32)
        words = ((e_len.astype(np.int64) + 15) >> 4)
        e_off = (np.cumsum(words) - words).astype(np.uint64)
        e_seq = rng.integers(0, 2 ** 32, int(words.sum()) + 1, dtype=np.uint64).astype(np.uint32)
        e_src = rng.integers(0, max(2 * n_nodes, 1), n_e).astype(np.uint32)
        e_dst = rng.integers(0, max(2 * n_nodes, 1), n_e).astype(np.uint32)
        e_rc = rng.permutation(n_e).astype(np.uint32)
        e_count = rng.integers(0, 2 ** 40, n_e).astype(np.uint64)
        fg = FlatGraph()
        fg.n_nodes, fg.n_e, fg.n_seq_words = n_nodes, n_e, int(words.sum())
        keep = dict(node_mask=mask, node_ebase=ebase, e_src=e_src, e_dst=e_dst, e_rc=e_rc, e_len=e_len, e_count=e_count, e_off=e_off, e_seq=e_seq)
        for name, a in keep.items():
            setattr(fg, name, a.ctypes.data)
        g = AsmGraph()
        assert lib.tagpu_fill_asm_graph_from_flat(C.byref(fg), 31, C.byref(g)) == 0
        assert (g.ksize, g.aux_flag, g.bin_size, g.n_v, g.n_e) == (31, 0, 0, 2 * n_nodes, n_e)
        e = 0
        for i in rng.permutation(n_nodes)[:200].tolist() if n_nodes else []:
            m = int(mask[i])
            first = int(ebase[i])
            for o, nib in ((0, m & 15), (1, m >> 4)):
                nd = g.nodes[2 * i + o]
                d = bin(nib).count("1")
                assert nd.rc_id == 2 * i + (o ^ 1) and nd.deg == d
                assert [nd.adj[a] for a in range(d)] == list(range(first, first + d))
                first += d
        for e in rng.permutation(n_e)[:300].tolist() if n_e else []:
            ed = g.edges[e]
            assert (ed.count, ed.seq_len, ed.n_holes, ed.source, ed.target, ed.rc_id) == (int(e_count[e]), int(e_len[e]), 0, int(e_src[e]), int(e_dst[e]), int(e_rc[e]))
            assert not ed.p_holes and not ed.l_holes and not ed.barcodes and bytes(ed.lock) == bytes(40)
            w = int(words[e])
            assert [ed.seq[x] for x in range(w)] == e_seq[int(e_off[e]):int(e_off[e]) + w].tolist()
        free_asm_graph(g)
        assert g.n_v == 0 and g.n_e == 0 and not g.nodes and not g.edges
