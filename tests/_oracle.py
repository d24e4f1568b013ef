"""ctypes access to the CPU oracle (oracle/liboracle.so). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
TA_REF = os.path.join(ORACLE_DIR, "_ref", "TA_ref")
TA_GPU = os.path.join(ORACLE_DIR, "_ref", "TA_gpu")
TA_KMC = os.path.join(ORACLE_DIR, "_ref", "TA_kmc")


class OraGraph(C.Structure):
    _fields_ = [
        ("ksize", C.c_int), ("n_kmer", C.c_int64),
        ("khi", C.POINTER(C.c_uint64)), ("klo", C.POINTER(C.c_uint64)), ("mask", C.POINTER(C.c_uint8)),
        ("n_v", C.c_int64), ("n_e", C.c_int64),
        ("node_rc", C.c_void_p), ("node_deg", C.c_void_p), ("node_adj_off", C.c_void_p), ("node_adj", C.c_void_p),
        ("e_src", C.c_void_p), ("e_dst", C.c_void_p), ("e_rc", C.c_void_p),
        ("e_count", C.POINTER(C.c_uint64)), ("e_len", C.POINTER(C.c_uint32)),
        ("e_seq_off", C.c_void_p), ("e_seq", C.c_void_p), ("n_kp1_on_edge", C.c_uint64),
    ]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.ora_count_stream.restype = C.c_int64
        lib.ora_count_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.ora_free.argtypes = [C.c_void_p]
        lib.ora_build_graph.restype = C.POINTER(OraGraph)
        lib.ora_build_graph.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ora_build_graph_local.restype = C.POINTER(OraGraph)
        lib.ora_build_graph_local.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                              C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
        lib.ora_graph_free.argtypes = [C.POINTER(OraGraph)]
        lib.ora_graph_save_bin.argtypes = [C.POINTER(OraGraph), C.c_char_p]
        lib.ora_canon_dump.restype = C.c_int
        lib.ora_canon_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.ora_bin_digest.restype = C.c_int
        lib.ora_bin_digest.argtypes = [C.c_char_p, C.POINTER(C.c_uint64)]
        lib.ora_coverage_recount.restype = C.c_int
        lib.ora_coverage_recount.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ora_load_reads.restype = C.c_int64
        lib.ora_load_reads.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p)]

    def count(self, stream, K, ci=2, threads=8):
        """-> dict(hi, lo, count (sorted by key), n_instances, n_distinct)"""
        a = np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else stream
        hi, lo, cnt = C.c_void_p(), C.c_void_p(), C.c_void_p()
        ni, nd = C.c_uint64(), C.c_uint64()
        n = self.lib.ora_count_stream(a.ctypes.data, a.size, K, ci, threads, C.byref(hi), C.byref(lo), C.byref(cnt),
                                      C.byref(ni), C.byref(nd))
        assert n >= 0
        def take(p, dt):
            out = np.ctypeslib.as_array(C.cast(p, C.POINTER(dt)), shape=(max(n, 1),))[:n].copy()
            self.lib.ora_free(p)
            return out
        return dict(hi=take(hi, C.c_uint64), lo=take(lo, C.c_uint64), count=take(cnt, C.c_uint32),
                    n_instances=ni.value, n_distinct=nd.value)

    def graph(self, k, hi, lo, count):
        hi, lo, count = np.ascontiguousarray(hi, np.uint64), np.ascontiguousarray(lo, np.uint64), np.ascontiguousarray(count, np.uint32)
        return self.lib.ora_build_graph(k, hi.size, hi.ctypes.data, lo.ctypes.data, count.ctypes.data)

    def graph_local(self, k, hi, lo, count, contigs, covs):
        """build_local_assembly_graph restatement: contigs = list of ACGT byte strings, covs = their coverages in g0"""
        hi, lo, count = np.ascontiguousarray(hi, np.uint64), np.ascontiguousarray(lo, np.uint64), np.ascontiguousarray(count, np.uint32)
        codes = [bytes(b"ACGT".index(ch) for ch in c) for c in contigs]
        n = len(codes)
        arr = (C.c_char_p * n)(*codes)
        return self.lib.ora_build_graph_local(k, hi.size, hi.ctypes.data, lo.ctypes.data, count.ctypes.data, n, arr,
                                              (C.c_uint32 * n)(*[len(c) for c in codes]), (C.c_double * n)(*covs))

    def graph_masks(self, g):
        n = g.contents.n_kmer
        f = lambda p: np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy()
        return f(g.contents.khi), f(g.contents.klo), f(g.contents.mask)

    def save_bin(self, g, path):
        assert self.lib.ora_graph_save_bin(g, os.fsencode(path)) == 0

    def free_graph(self, g):
        self.lib.ora_graph_free(g)

    def canon(self, bin_path, out_path, mode=0):
        """returns number of structural violations (0 = valid graph)"""
        return self.lib.ora_canon_dump(os.fsencode(bin_path), os.fsencode(out_path), mode)

    def bin_digest(self, bin_path):
        """Order-independent edge digest of a .bin (oracle/canon_dump.c) in the vocabulary of Tagpu.digest()."""
        out = (C.c_uint64 * 5)()
        assert self.lib.ora_bin_digest(os.fsencode(bin_path), out) == 0
        return {"edge_sum": out[0], "edge_xor": out[1], "edge_len_sum": out[2], "edge_count_sum": out[3], "n_e": out[4]}

    def coverage_recount(self, stream, e_len, e_off, e_seq, e_rc):
        """kmer_count_on_edges + add_cnt_to_graph (oracle/cov_oracle.c) on flat edge arrays -> uint64 count per edge"""
        a = np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else np.ascontiguousarray(stream)
        e_len = np.ascontiguousarray(e_len, np.uint32); e_off = np.ascontiguousarray(e_off, np.uint64)
        e_seq = np.ascontiguousarray(e_seq, np.uint32); e_rc = np.ascontiguousarray(e_rc, np.int64)
        out = np.zeros(e_len.size, np.uint64)
        assert self.lib.ora_coverage_recount(a.ctypes.data, a.size, e_len.size, e_len.ctypes.data, e_off.ctypes.data,
                                             e_seq.ctypes.data, e_rc.ctypes.data, out.ctypes.data) == 0
        return out

    def load_reads(self, files):
        arr = (C.c_char_p * len(files))()
        arr[:] = [os.fsencode(f) for f in files]
        out = C.c_void_p()
        n = self.lib.ora_load_reads(len(files), arr, C.byref(out))
        a = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(max(n, 1),))[:n].copy()
        self.lib.ora_free(out)
        return a


_cached = None


def load():
    global _cached
    if _cached is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so", "ta_oracle"], check=True, capture_output=True)
        _cached = Oracle(C.CDLL(LIB))
    return _cached


def canon_text(oracle, bin_path, mode=0):
    out = bin_path + f".canon{mode}.txt"
    bad = oracle.canon(bin_path, out, mode)
    with open(out, "rb") as f:
        return bad, f.read()


def load_bin(path):
    """Parses a graph .bin (save_asm_graph layout) -> dict(ksize, n_v, n_e, edges=[dict(...)])."""
    import struct
    b = open(path, "rb").read()
    assert b[:4] == b"asmg"
    aux, ksize, n_v, n_e = struct.unpack_from("<Iiqq", b, 4)
    o = 28
    for _ in range(n_v):
        _, deg = struct.unpack_from("<qq", b, o)
        o += 16 + 8 * deg
    edges = []
    for e in range(n_e):
        src, dst = struct.unpack_from("<qq", b, o)
        o += 16
        if src == -1:
            edges.append(None)
            continue
        rc, count, len8 = struct.unpack_from("<qQQ", b, o)
        o += 24
        seq_len = len8 & 0xffffffff
        nw = (seq_len + 15) >> 4
        words = struct.unpack_from(f"<{nw}I", b, o)
        o += 4 * nw
        (n_holes,) = struct.unpack_from("<I", b, o)
        o += 4 + 8 * n_holes
        seq = bytes(b"ACGT"[(words[i >> 4] >> ((i & 15) << 1)) & 3] for i in range(seq_len))
        edges.append(dict(src=src, dst=dst, rc=rc, count=count, seq_len=seq_len, n_holes=n_holes, seq=seq))
    return dict(ksize=ksize, n_v=n_v, n_e=n_e, edges=edges)


def load_bin_flat(path):
    """Graph .bin (save_asm_graph layout) -> flat edge arrays in file order:
    dict(ksize, n_v, n_e, e_src, e_dst, e_rc (int64), e_count (uint64), e_len (uint32), e_off (uint64, words), e_seq (uint32))."""
    import struct
    b = open(path, "rb").read()
    assert b[:4] == b"asmg"
    _, ksize, n_v, n_e = struct.unpack_from("<Iiqq", b, 4)
    o = 28
    for _ in range(n_v):
        (deg,) = struct.unpack_from("<q", b, o + 8)
        o += 16 + 8 * deg
    src, dst, rc = np.full(n_e, -1, np.int64), np.full(n_e, -1, np.int64), np.arange(n_e, dtype=np.int64)
    cnt, ln, off = np.zeros(n_e, np.uint64), np.zeros(n_e, np.uint32), np.zeros(n_e, np.uint64)
    words, n_w = [], 0
    for e in range(n_e):
        src[e], dst[e] = struct.unpack_from("<qq", b, o)
        o += 16
        if src[e] == -1:
            continue
        r, c, len8 = struct.unpack_from("<qQQ", b, o)
        o += 24
        rc[e], cnt[e], ln[e] = r, c, len8 & 0xffffffff
        nw = (int(ln[e]) + 15) >> 4
        off[e] = n_w
        words.append(np.frombuffer(b, dtype="<u4", count=nw, offset=o))
        n_w += nw
        o += 4 * nw
        (n_holes,) = struct.unpack_from("<I", b, o)
        o += 4 + 8 * n_holes
    seq = np.concatenate(words).astype(np.uint32) if words else np.zeros(1, np.uint32)
    return dict(ksize=ksize, n_v=n_v, n_e=n_e, e_src=src, e_dst=dst, e_rc=rc, e_count=cnt, e_len=ln, e_off=off, e_seq=seq)
