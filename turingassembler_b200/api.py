"""ctypes mirror of include/tagpu.h (and of the reference entry points libtagpu.so exports).

Function names and argument meaning follow the reference:
  * ``kmc_build_kmer_database``              -> KMC_build_kmer_database   (/root/reference/include/kmc_skipping.h:8-9)
  * ``build_graph_from_scratch``             -> /root/reference/src/kmer_build.h:17-19
  * ``build_graph_from_scratch_without_count`` -> /root/reference/src/kmer_build.h:20-22
Everything is executed by libtagpu.so on the GPU; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

IPC_HANDLE_BYTES = 64  # TAGPU_IPC_HANDLE_BYTES
# TAGPU_LIB: developer override (e.g. a build with -DTAGPU_TIMING); the product library is libtagpu.so next to this file
LIB_PATH = os.environ.get("TAGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtagpu.so")


class TagpuError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [
        ("n_instances", C.c_uint64),
        ("n_distinct", C.c_uint64),
        ("n_solid", C.c_uint64),
        ("sum_solid", C.c_uint64),
        ("n_kmers", C.c_uint64),
        ("n_v", C.c_uint64),
        ("n_e", C.c_uint64),
        ("n_seq_words", C.c_uint64),
        ("n_kp1_on_edge", C.c_uint64),
        ("error", C.c_uint64),
        ("jump_rounds", C.c_uint64),
        ("gpu_launches", C.c_uint64),
        ("ms_count", C.c_float),
        ("ms_graph", C.c_float),
        ("ms_total", C.c_float),
        ("record_bytes", C.c_uint32),
        ("n_records_local", C.c_uint64),
        ("n_records_peer", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class FlatGraph(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint64),
        ("n_e", C.c_uint64),
        ("n_seq_words", C.c_uint64),
        ("node_mask", C.c_void_p),
        ("node_ebase", C.c_void_p),
        ("e_src", C.c_void_p),
        ("e_dst", C.c_void_p),
        ("e_rc", C.c_void_p),
        ("e_len", C.c_void_p),
        ("e_count", C.c_void_p),
        ("e_off", C.c_void_p),
        ("e_seq", C.c_void_p),
    ]


# --- the reference's graph ABI (include/tagpu_graph.h) -----------------------------------------------------------
class BarcodeHash(C.Structure):
    _fields_ = [("size", C.c_uint32), ("n_item", C.c_uint32), ("n_unique", C.c_uint32),
                ("keys", C.c_void_p), ("cnts", C.c_void_p)]


class PthreadMutex(C.Structure):
    _fields_ = [("opaque", C.c_char * 40)]  # sizeof(pthread_mutex_t) on x86-64 glibc
    _align_ = 8


class AsmNode(C.Structure):
    _fields_ = [("rc_id", C.c_int64), ("deg", C.c_int64), ("adj", C.POINTER(C.c_int64))]


class AsmEdge(C.Structure):
    _fields_ = [
        ("count", C.c_uint64),
        ("seq", C.POINTER(C.c_uint32)),
        ("seq_len", C.c_uint32),
        ("n_holes", C.c_uint32),
        ("p_holes", C.c_void_p),
        ("l_holes", C.c_void_p),
        ("source", C.c_int64),
        ("target", C.c_int64),
        ("rc_id", C.c_int64),
        ("lock", C.c_uint64 * 5),
        ("barcodes", C.c_void_p),
        ("barcodes_scaf", BarcodeHash),
        ("barcodes_cov", BarcodeHash),
    ]


class AsmGraph(C.Structure):
    _fields_ = [
        ("ksize", C.c_int),
        ("bin_size", C.c_int),
        ("aux_flag", C.c_uint32),
        ("n_v", C.c_int64),
        ("n_e", C.c_int64),
        ("nodes", C.POINTER(AsmNode)),
        ("edges", C.POINTER(AsmEdge)),
        ("candidates", C.c_void_p),
    ]


class LocalJob(C.Structure):
    """struct tagpu_local_job (include/tagpu.h)"""
    _fields_ = [("reads", C.c_void_p), ("n_bytes", C.c_uint64), ("k", C.c_int), ("cutoff", C.c_int),
                ("contigs", C.c_void_p), ("n_contig_bytes", C.c_uint64), ("n_contigs", C.c_int),
                ("contig_off", C.POINTER(C.c_uint64)), ("contig_len", C.POINTER(C.c_uint32)), ("contig_cov", C.POINTER(C.c_double)),
                ("g", C.POINTER(AsmGraph)), ("rc", C.c_int), ("stats", Stats)]


_lib = None


def load_library() -> C.CDLL:
    """Loads libtagpu.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TagpuError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    lib.tagpu_create.restype = vp
    lib.tagpu_create.argtypes = [i32]
    lib.tagpu_destroy.argtypes = [vp]
    lib.tagpu_set_stream.argtypes = [vp, vp]
    lib.tagpu_set_cutoff.argtypes = [vp, i32]
    lib.tagpu_set_skip_counts.argtypes = [vp, i32]
    lib.tagpu_set_profile.argtypes = [vp, i32]
    lib.tagpu_set_contract.argtypes = [vp, i32]
    lib.tagpu_profile_json.restype = C.c_char_p
    lib.tagpu_profile_json.argtypes = [vp]
    lib.tagpu_last_error.restype = C.c_char_p
    lib.tagpu_last_error.argtypes = [vp]
    for name in ("tagpu_build_device", "tagpu_build_host", "tagpu_count_device", "tagpu_count_host",
                 "tagpu_build_device_packed", "tagpu_build_host_packed", "tagpu_count_host_packed"):
        f = getattr(lib, name)
        f.restype = i32
        f.argtypes = [vp, vp, u64, i32]
    lib.tagpu_build_local_host.restype = i32
    lib.tagpu_build_local_host.argtypes = [vp, vp, u64, i32, vp, u64, i32, C.POINTER(u64), C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    lib.build_local_assembly_graph.restype = None
    lib.build_local_assembly_graph.argtypes = [i32, i32, i32, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_char_p,
                                               C.POINTER(AsmGraph), C.POINTER(AsmGraph), C.c_int64, C.c_int64]
    lib.tagpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.tagpu_build_local_batch.restype = i32
    lib.tagpu_build_local_batch.argtypes = [i32, i32, C.POINTER(LocalJob), i32]
    lib.tagpu_copy_solid.restype = i32
    lib.tagpu_copy_solid.argtypes = [vp, vp, vp, vp]
    lib.tagpu_copy_kmers.restype = i32
    lib.tagpu_copy_kmers.argtypes = [vp, vp, vp, vp]
    lib.tagpu_copy_graph.restype = i32
    lib.tagpu_copy_graph.argtypes = [vp, C.POINTER(FlatGraph)]
    lib.tagpu_coverage_recount_host.restype = i32
    lib.tagpu_coverage_recount_host.argtypes = [vp, vp, u64, u64, vp, vp, vp, u64, vp, vp]
    lib.tagpu_digest.restype = i32
    lib.tagpu_digest.argtypes = [vp, C.POINTER(u64)]
    lib.tagpu_fill_asm_graph.restype = i32
    lib.tagpu_fill_asm_graph.argtypes = [vp, C.POINTER(AsmGraph)]
    lib.tagpu_free_asm_graph.restype = None
    lib.tagpu_free_asm_graph.argtypes = [C.POINTER(AsmGraph)]
    lib.tagpu_write_graph_bin.restype = i32
    lib.tagpu_write_graph_bin.argtypes = [vp, C.c_char_p]
    lib.tagpu_write_kmc_db.restype = i32
    lib.tagpu_write_kmc_db.argtypes = [vp, C.c_char_p]
    lib.tagpu_raw_ring.restype = vp
    lib.tagpu_raw_ring.argtypes = [vp, C.c_size_t]
    lib.tagpu_raw_begin.restype = i32
    lib.tagpu_raw_begin.argtypes = [vp, u64]
    lib.tagpu_raw_put.restype = i32
    lib.tagpu_raw_put.argtypes = [vp, u64, vp, u64, i32]
    lib.tagpu_raw_slot_wait.restype = i32
    lib.tagpu_raw_slot_wait.argtypes = [vp, i32]
    lib.tagpu_parse_fastq_device.restype = C.c_int64
    lib.tagpu_parse_fastq_device.argtypes = [vp, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_uint8), vp]
    lib.tagpu_build_fastq_device.restype = i32
    lib.tagpu_build_fastq_device.argtypes = [vp, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_uint8), i32, i32]
    lib.tagpu_load_reads.restype = C.c_int64
    lib.tagpu_load_reads.argtypes = [i32, C.POINTER(C.c_char_p), i32, C.POINTER(vp)]
    lib.tagpu_free_reads.argtypes = [vp]
    lib.tagpu_dist_plan.restype = i32
    lib.tagpu_dist_plan.argtypes = [vp, i32, i32, u64, i32, vp]
    lib.tagpu_dist_connect.restype = i32
    lib.tagpu_dist_connect.argtypes = [vp, vp]
    lib.tagpu_dist_partition.restype = i32
    lib.tagpu_dist_partition.argtypes = [vp, vp, u64]
    lib.tagpu_dist_partition_host.restype = i32
    lib.tagpu_dist_partition_host.argtypes = [vp, vp, u64]
    lib.tagpu_dist_partition_host_packed.restype = i32
    lib.tagpu_dist_partition_host_packed.argtypes = [vp, vp, u64]
    lib.tagpu_packed_bytes.restype = u64
    lib.tagpu_packed_bytes.argtypes = [u64]
    lib.tagpu_pack_stream.restype = i32
    lib.tagpu_pack_stream.argtypes = [vp, u64, vp, i32]
    lib.tagpu_dist_count.restype = i32
    lib.tagpu_dist_count.argtypes = [vp, C.POINTER(u64)]
    lib.tagpu_dist_graph.restype = i32
    lib.tagpu_dist_graph.argtypes = [vp, C.POINTER(u64), i32]
    lib.tagpu_dist_contract.restype = i32
    lib.tagpu_dist_contract.argtypes = [vp, C.POINTER(u64)]
    lib.tagpu_dist_graph_paths.restype = i32
    lib.tagpu_dist_graph_paths.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), i32]
    lib.tagpu_dist_step.restype = i32
    lib.tagpu_dist_step.argtypes = [vp, vp, vp, u64, i32, i32, C.POINTER(i32)]
    lib.tagpu_dist_close.argtypes = [vp]
    lib.tagpu_dist_disconnect.argtypes = [vp]
    lib.tagpu_dist_shard_range.restype = None
    lib.tagpu_dist_shard_range.argtypes = [vp, u64, i32, i32, C.POINTER(u64), C.POINTER(u64)]
    lib.KMC_build_kmer_database.restype = i32
    lib.KMC_build_kmer_database.argtypes = [i32, C.c_char_p, i32, i32, i32, C.POINTER(C.c_char_p)]
    for name in ("build_graph_from_scratch", "build_graph_from_scratch_without_count"):
        f = getattr(lib, name)
        f.restype = None
        f.argtypes = [i32, i32, i32, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_char_p, C.POINTER(AsmGraph)]
    _lib = lib
    return lib


def _char_pp(paths: Sequence[str]):
    arr = (C.c_char_p * len(paths))()
    arr[:] = [os.fsencode(p) for p in paths]
    return arr


class Tagpu:
    """One GPU context (tagpu_ctx). All heavy lifting happens inside libtagpu.so."""

    def __init__(self, device: int = -1, cutoff: int = 2):
        self.lib = load_library()
        self.ctx = self.lib.tagpu_create(device)
        if not self.ctx:
            raise TagpuError("tagpu_create failed: no usable CUDA device (libtagpu has no CPU fallback)")
        self.lib.tagpu_set_cutoff(self.ctx, cutoff)
        self.contract = os.environ.get("TAGPU_CONTRACT", "1") != "0"   # two-level graph stage (library default: on)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.tagpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise TagpuError(self.lib.tagpu_last_error(self.ctx).decode())

    def set_stream(self, cuda_stream: int):
        self.lib.tagpu_set_stream(self.ctx, C.c_void_p(cuda_stream))

    def set_cutoff(self, ci: int):
        self.lib.tagpu_set_cutoff(self.ctx, ci)

    def set_contract(self, on: bool):
        self.contract = bool(on)
        self.lib.tagpu_set_contract(self.ctx, int(on))

    def set_profile(self, on: bool):
        self.lib.tagpu_set_profile(self.ctx, int(on))

    def profile(self) -> dict:
        import json
        return json.loads(self.lib.tagpu_profile_json(self.ctx).decode())

    def set_skip_counts(self, skip: bool):
        self.lib.tagpu_set_skip_counts(self.ctx, int(skip))

    # ---- builds ----
    def build_device(self, d_ptr: int, n_bytes: int, k: int):
        """d_ptr: device address of the flat read stream (e.g. torch tensor .data_ptr())."""
        self._check(self.lib.tagpu_build_device(self.ctx, C.c_void_p(d_ptr), n_bytes, k))
        return self.stats()

    def count_device(self, d_ptr: int, n_bytes: int, K: int):
        self._check(self.lib.tagpu_count_device(self.ctx, C.c_void_p(d_ptr), n_bytes, K))
        return self.stats()

    def build_host(self, stream, k: int):
        """stream: bytes / numpy uint8 array / host address+length tuple."""
        ptr, n, keep = _host_buffer(stream)
        self._check(self.lib.tagpu_build_host(self.ctx, ptr, n, k))
        del keep
        return self.stats()

    def build_host_packed(self, packed, n_positions: int, k: int):
        """packed: numpy uint8 array (or host address) from pack_stream(); n_positions = bytes of the ASCII stream."""
        ptr = C.c_void_p(packed) if isinstance(packed, int) else C.c_void_p(packed.ctypes.data)
        self._check(self.lib.tagpu_build_host_packed(self.ctx, ptr, n_positions, k))
        return self.stats()

    def count_host_packed(self, packed, n_positions: int, K: int):
        ptr = C.c_void_p(packed) if isinstance(packed, int) else C.c_void_p(packed.ctypes.data)
        self._check(self.lib.tagpu_count_host_packed(self.ctx, ptr, n_positions, K))
        return self.stats()

    def build_device_packed(self, d_ptr: int, n_positions: int, k: int):
        self._check(self.lib.tagpu_build_device_packed(self.ctx, C.c_void_p(d_ptr), n_positions, k))
        return self.stats()

    def parse_fastq(self, blobs: Sequence[bytes]) -> bytes:
        """Raw FASTQ file contents -> the read stream, parsed on the device (tagpu_parse_fastq_device; what the files entry
        points do with plain FASTQ, here fed from memory: the bytes go up through the pinned ring in 1 MB pieces)."""
        n = len(blobs)
        off, ln, nl = (C.c_uint64 * n)(), (C.c_uint64 * n)(), (C.c_uint8 * n)()
        total = 0
        for i, b in enumerate(blobs):
            off[i], ln[i], nl[i] = total, len(b), 1 if b.endswith(b"\n") else 0
            total += (len(b) + 255) & ~255
        slot, n_slots = 1 << 20, 4
        ring = self.lib.tagpu_raw_ring(self.ctx, slot * n_slots)
        if not ring:
            raise TagpuError("cannot allocate the pinned ring")
        self._check(self.lib.tagpu_raw_begin(self.ctx, total))
        c = 0
        for i, b in enumerate(blobs):
            for o in range(0, len(b), slot):
                piece = b[o:o + slot]
                s_ = c % n_slots
                if c >= n_slots:
                    self._check(self.lib.tagpu_raw_slot_wait(self.ctx, s_))
                C.memmove(ring + s_ * slot, piece, len(piece))
                self._check(self.lib.tagpu_raw_put(self.ctx, off[i] + o, C.c_void_p(ring + s_ * slot), len(piece), s_))
                c += 1
        m = self.lib.tagpu_parse_fastq_device(self.ctx, n, off, ln, nl, None)
        if m < 0:
            raise TagpuError(f"tagpu_parse_fastq_device returned {m}: {self.lib.tagpu_last_error(self.ctx).decode()}")
        out = np.empty(max(m, 1), dtype=np.uint8)
        m2 = self.lib.tagpu_parse_fastq_device(self.ctx, n, off, ln, nl, C.c_void_p(out.ctypes.data))
        assert m2 == m
        return out[:m].tobytes()

    def build_local_host(self, stream, k: int, contigs: Sequence[bytes], contig_cov: Sequence[float]):
        """build_local_assembly_graph on host buffers: reads + flanking contigs (ACGT bytes) with their coverages."""
        ptr, n, keep = _host_buffer(stream)
        txt = b"".join(c + b"\n" for c in contigs)
        offs, o = [], 0
        for c in contigs:
            offs.append(o)
            o += len(c) + 1
        nc = len(contigs)
        self._check(self.lib.tagpu_build_local_host(
            self.ctx, ptr, n, k, C.c_char_p(txt), len(txt), nc, (C.c_uint64 * nc)(*offs),
            (C.c_uint32 * nc)(*[len(c) for c in contigs]), (C.c_double * nc)(*contig_cov)))
        del keep
        return self.stats()

    def count_host(self, stream, K: int):
        ptr, n, keep = _host_buffer(stream)
        self._check(self.lib.tagpu_count_host(self.ctx, ptr, n, K))
        del keep
        return self.stats()

    # ---- multi-GPU phases (include/tagpu.h "multi-GPU"; orchestration in turingassembler_b200/dist.py) ----
    def dist_plan(self, rank: int, world: int, n_total_bytes: int, k: int) -> bytes:
        h = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self.lib.tagpu_dist_plan(self.ctx, rank, world, n_total_bytes, k, h))
        return h.raw

    def dist_connect(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        self._check(self.lib.tagpu_dist_connect(self.ctx, C.c_char_p(blob)))

    def dist_partition(self, d_ptr: int, n_local_bytes: int):
        self._check(self.lib.tagpu_dist_partition(self.ctx, C.c_void_p(d_ptr), n_local_bytes))

    def dist_partition_host(self, h_ptr: int, n_local_bytes: int):
        self._check(self.lib.tagpu_dist_partition_host(self.ctx, C.c_void_p(h_ptr), n_local_bytes))

    def dist_partition_host_packed(self, h_ptr: int, n_local_positions: int):
        self._check(self.lib.tagpu_dist_partition_host_packed(self.ctx, C.c_void_p(h_ptr), n_local_positions))

    def dist_count(self):
        out = (C.c_uint64 * 4)()
        self._check(self.lib.tagpu_dist_count(self.ctx, out))
        return list(out)

    def dist_graph(self, all_stats: Sequence[int], with_graph: bool = True):
        arr = (C.c_uint64 * len(all_stats))(*all_stats)
        self._check(self.lib.tagpu_dist_graph(self.ctx, arr, int(with_graph)))
        return self.stats()

    def dist_contract(self):
        """Level 1 of the two-level graph stage on this rank's solid set -> [paths, interior words, hidden k-mers, ok]."""
        out = (C.c_uint64 * 4)()
        self._check(self.lib.tagpu_dist_contract(self.ctx, out))
        return list(out)

    def dist_graph_paths(self, all_stats: Sequence[int], all_paths: Sequence[int], gather_solid: bool = False):
        a = (C.c_uint64 * len(all_stats))(*all_stats)
        b = (C.c_uint64 * len(all_paths))(*all_paths)
        self._check(self.lib.tagpu_dist_graph_paths(self.ctx, a, b, 3 if gather_solid else 1))
        return self.stats()

    def dist_step(self, shm, ptr: int, n: int, src_kind: int, flags: int):
        """One whole multi-GPU step in a single call (tagpu_dist_step) -> (global stats, used_paths)."""
        used = C.c_int(0)
        self._check(self.lib.tagpu_dist_step(self.ctx, C.c_void_p(shm), C.c_void_p(ptr), n, src_kind, flags, C.byref(used)))
        return self.stats(), used.value

    def dist_disconnect(self):
        self.lib.tagpu_dist_disconnect(self.ctx)

    def dist_close(self):
        self.lib.tagpu_dist_close(self.ctx)

    def stats(self) -> dict:
        st = Stats()
        self.lib.tagpu_get_stats(self.ctx, C.byref(st))
        return st.as_dict()

    # ---- results ----
    def solid(self):
        n = self.stats()["n_solid"]
        hi, lo, cnt = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint32)
        self._check(self.lib.tagpu_copy_solid(self.ctx, hi.ctypes.data, lo.ctypes.data, cnt.ctypes.data))
        return hi, lo, cnt

    def kmers(self):
        n = self.stats()["n_kmers"]
        hi, lo, mask = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint8)
        self._check(self.lib.tagpu_copy_kmers(self.ctx, hi.ctypes.data, lo.ctypes.data, mask.ctypes.data))
        return hi, lo, mask

    def graph(self) -> dict:
        st = self.stats()
        nn, ne, nw = st["n_v"] // 2, st["n_e"], st["n_seq_words"]
        arrs = {
            "node_mask": np.zeros(nn + 1, np.uint8), "node_ebase": np.zeros(nn + 1, np.uint32),
            "e_src": np.zeros(ne + 1, np.uint32), "e_dst": np.zeros(ne + 1, np.uint32),
            "e_rc": np.zeros(ne + 1, np.uint32), "e_len": np.zeros(ne + 1, np.uint32),
            "e_count": np.zeros(ne + 1, np.uint64), "e_off": np.zeros(ne + 1, np.uint64),
            "e_seq": np.zeros(nw + 1, np.uint32),
        }
        fg = FlatGraph()
        for name, a in arrs.items():
            setattr(fg, name, a.ctypes.data)
        self._check(self.lib.tagpu_copy_graph(self.ctx, C.byref(fg)))
        out = {"n_nodes": nn, "n_e": ne, "n_seq_words": nw}
        out.update({name: a[: {"node_mask": nn, "node_ebase": nn, "e_seq": nw}.get(name, ne)] for name, a in arrs.items()})
        return out

    def coverage_recount(self, stream, edges: dict | None = None) -> np.ndarray:
        """Coverage recount (kmer_count_on_edges + add_cnt_to_graph, /root/reference/src/coverage/kmer_count.c:198-240,113-135)
        of the reads in `stream` on the edges of the last build (edges=None: they are still on the device) or on flat edge
        arrays dict(e_len, e_off [32-bit words], e_seq, e_rc).  -> uint64 count per edge."""
        ptr, n, keep = _host_buffer(stream)
        if edges is None:
            n_e = self.stats()["n_e"]
            out = np.zeros(n_e + 1, np.uint64)
            self._check(self.lib.tagpu_coverage_recount_host(self.ctx, ptr, n, 0, None, None, None, 0, None, out.ctypes.data))
        else:
            e_len = np.ascontiguousarray(edges["e_len"], np.uint32)
            e_off = np.ascontiguousarray(edges["e_off"], np.uint64)
            e_seq = np.ascontiguousarray(edges["e_seq"], np.uint32)
            e_rc = np.ascontiguousarray(edges["e_rc"], np.uint32)
            n_e = int(e_len.size)
            out = np.zeros(n_e + 1, np.uint64)
            self._check(self.lib.tagpu_coverage_recount_host(self.ctx, ptr, n, n_e, e_len.ctypes.data, e_off.ctypes.data, e_seq.ctypes.data,
                                                             int(e_seq.size), e_rc.ctypes.data, out.ctypes.data))
        del keep
        return out[:n_e]

    def digest(self) -> dict:
        """Order-independent digests of the last build, computed on the device (tagpu_digest, include/tagpu.h)."""
        out = (C.c_uint64 * 9)()
        self._check(self.lib.tagpu_digest(self.ctx, out))
        return {"solid_sum": out[0], "solid_xor": out[1], "solid_n": out[2], "solid_complete": bool(out[3]),
                "edge_sum": out[4], "edge_xor": out[5], "edge_len_sum": out[6], "edge_count_sum": out[7], "n_e": out[8]}

    def copy_graph_into(self, fg: "FlatGraph"):
        """tagpu_copy_graph into caller-owned (e.g. pinned) host arrays described by a FlatGraph of addresses."""
        self._check(self.lib.tagpu_copy_graph(self.ctx, C.byref(fg)))
        return fg

    def write_graph_bin(self, path: str):
        self._check(self.lib.tagpu_write_graph_bin(self.ctx, os.fsencode(path)))

    def write_kmc_db(self, working_dir: str):
        self._check(self.lib.tagpu_write_kmc_db(self.ctx, os.fsencode(working_dir)))

    def fill_asm_graph(self) -> AsmGraph:
        g = AsmGraph()
        self._check(self.lib.tagpu_fill_asm_graph(self.ctx, C.byref(g)))
        return g


def _host_buffer(stream):
    if isinstance(stream, (bytes, bytearray)):
        a = np.frombuffer(stream, dtype=np.uint8)
    elif isinstance(stream, np.ndarray):
        a = np.ascontiguousarray(stream.view(np.uint8))
    elif isinstance(stream, tuple):
        return C.c_void_p(stream[0]), int(stream[1]), None
    else:
        raise TypeError(f"unsupported stream type {type(stream)}")
    return C.c_void_p(a.ctypes.data), int(a.size), a


def load_reads(files: Sequence[str], n_threads: int = 4):
    """FASTQ/FASTA(.gz) -> (pinned host address, n_bytes); release with free_reads."""
    lib = load_library()
    out = C.c_void_p()
    n = lib.tagpu_load_reads(len(files), _char_pp(files), n_threads, C.byref(out))
    return out.value, n


def shard_range(stream: np.ndarray, rank: int, world: int):
    """[begin, end) of `rank`'s share of a host read stream, cut at read boundaries (tagpu_dist_shard_range)."""
    lib = load_library()
    a = np.ascontiguousarray(stream.view(np.uint8))
    b, e = C.c_uint64(), C.c_uint64()
    lib.tagpu_dist_shard_range(C.c_void_p(a.ctypes.data), a.size, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value


def packed_bytes(n_positions: int) -> int:
    return int(load_library().tagpu_packed_bytes(n_positions))


def pack_stream(stream, threads: int = 8, out: np.ndarray | None = None) -> np.ndarray:
    """ASCII read stream (bytes / uint8 array) -> packed read stream (include/tagpu.h), packed on the host by libtagpu.so."""
    ptr, n, keep = _host_buffer(stream)
    lib = load_library()
    if out is None:
        out = np.empty(int(lib.tagpu_packed_bytes(n)), dtype=np.uint8)
    if lib.tagpu_pack_stream(ptr, n, C.c_void_p(out.ctypes.data), threads) != 0:
        raise TagpuError("tagpu_pack_stream failed (host and device tile layouts disagree)")
    del keep
    return out


def build_local_batch(jobs, n_ctx: int = 8, device: int = 0, fill: bool = False):
    """tagpu_build_local_batch: `jobs` = list of dict(stream=bytes/ndarray, k=int, contigs=[bytes], covs=[float]).
    -> (list of stats dicts, list of AsmGraph or None).  The builds run concurrently on n_ctx contexts."""
    lib = load_library()
    arr = (LocalJob * len(jobs))()
    keep, graphs = [], []
    for j, job in zip(arr, jobs):
        a = np.frombuffer(job["stream"], dtype=np.uint8) if isinstance(job["stream"], (bytes, bytearray)) else np.ascontiguousarray(job["stream"])
        txt = b"".join(c + b"\n" for c in job["contigs"])
        nc = len(job["contigs"])
        offs, o = [], 0
        for c in job["contigs"]:
            offs.append(o)
            o += len(c) + 1
        off_a, len_a, cov_a = (C.c_uint64 * nc)(*offs), (C.c_uint32 * nc)(*[len(c) for c in job["contigs"]]), (C.c_double * nc)(*job["covs"])
        txt_b = C.create_string_buffer(txt, len(txt) + 1)
        g = AsmGraph() if fill else None
        keep.append((a, txt_b, off_a, len_a, cov_a, g))
        graphs.append(g)
        j.reads, j.n_bytes, j.k, j.cutoff = a.ctypes.data, a.size, job["k"], job.get("cutoff", 2)
        j.contigs, j.n_contig_bytes, j.n_contigs = C.cast(txt_b, C.c_void_p), len(txt), nc
        j.contig_off, j.contig_len, j.contig_cov = off_a, len_a, cov_a
        j.g = C.pointer(g) if fill else None
    rc = lib.tagpu_build_local_batch(device, n_ctx, arr, len(jobs))
    if rc != 0:
        raise TagpuError(f"tagpu_build_local_batch: {sum(1 for j in arr if j.rc)} of {len(jobs)} jobs failed")
    del keep
    return [j.stats.as_dict() for j in arr], graphs


def free_asm_graph(g: AsmGraph):
    """Releases a graph returned by build_graph_from_scratch / Tagpu.fill_asm_graph (tagpu_free_asm_graph)."""
    load_library().tagpu_free_asm_graph(C.byref(g))


def free_reads(addr: int):
    load_library().tagpu_free_reads(C.c_void_p(addr))


# ---- the reference's entry points, same names and argument order -------------------------------------------------
def kmc_build_kmer_database(ksize: int, working_dir: str, n_threads: int, mmem: int, files: Sequence[str]) -> int:
    lib = load_library()
    return lib.KMC_build_kmer_database(ksize, os.fsencode(working_dir), n_threads, mmem, len(files), _char_pp(files))


def build_graph_from_scratch(ksize: int, n_threads: int, mmem: int, files_1: Sequence[str], files_2: Sequence[str],
                             work_dir: str) -> AsmGraph:
    lib = load_library()
    g = AsmGraph()
    lib.build_graph_from_scratch(ksize, n_threads, mmem, len(files_1), _char_pp(files_1), _char_pp(files_2),
                                 os.fsencode(work_dir), C.byref(g))
    return g


def build_graph_from_scratch_without_count(ksize: int, n_threads: int, mmem: int, files_1: Sequence[str],
                                           files_2: Sequence[str], work_dir: str) -> AsmGraph:
    lib = load_library()
    g = AsmGraph()
    lib.build_graph_from_scratch_without_count(ksize, n_threads, mmem, len(files_1), _char_pp(files_1),
                                               _char_pp(files_2), os.fsencode(work_dir), C.byref(g))
    return g
