"""Multi-GPU orchestration of the k-mer stage: one process per GPU (include/tagpu.h "multi-GPU").

This module is host plumbing only.  torch.distributed supplies rendezvous, the barriers between phases and the two tiny
host-side exchanges (IPC handles once, 4 + 4 values per step); the data path — super-k-mer records travelling to the GPU
that owns their bucket, and the contracted paths (or the solid sets) travelling back — is done by libtagpu.so itself over
NVLink peer memory (peer loads inside k_count_buckets and k_gather_paths, peer copies in tagpu_dist_graph).  The reference has no counterpart: it is a single
process (pthreads only).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np


def shard_reads(n_reads: int, rank: int, world: int):
    """Contiguous 1/world slice of the read list for `rank`: [first, last).  Every read belongs to exactly one rank."""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def owner_of_bucket(bucket: int, n_buckets: int, world: int) -> int:
    """Rank that counts `bucket` (mirror of PartCfg.per_rank in csrc/tagpu_count.cuh): contiguous ranges."""
    per_rank = (n_buckets + world - 1) // world
    return bucket // per_rank


class DistTagpu:
    """Drives the tagpu_dist_* phases of one rank.  `tagpu` is a turingassembler_b200.Tagpu bound to this rank's GPU."""

    def __init__(self, tagpu, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist
        self.t, self.rank, self.world, self.group = tagpu, rank, world, group
        self.torch, self.dist = torch, dist
        self.on_gpu = dist.get_backend(group) == "nccl"
        self.dev = torch.device("cuda", torch.cuda.current_device()) if self.on_gpu else torch.device("cpu")
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._stats = torch.zeros(4, dtype=torch.int64, device=self.dev)
        self._all = torch.zeros(4 * world, dtype=torch.int64, device=self.dev)
        self._dirty = False           # the other ranks may still be pulling paths out of this rank's regions
        self._shm = None
        self._open_shm()

    def _open_shm(self):
        """Intra-node rendezvous (tagpu_shm_*, include/tagpu.h): barrier and counter exchange over a shared-memory segment,
        a few microseconds each instead of a collective launch + device round trip.  Rank 0 picks the name; if the
        segment cannot be opened everywhere (ranks on different hosts) the torch.distributed collectives stay in use."""
        import ctypes as C
        import uuid
        if os.environ.get("TAGPU_NO_SHM"):
            return
        from .api import load_library
        lib = self._lib = load_library()
        lib.tagpu_shm_open.restype = C.c_void_p
        lib.tagpu_shm_open.argtypes = [C.c_char_p, C.c_int, C.c_int]
        lib.tagpu_shm_barrier.argtypes = [C.c_void_p]
        lib.tagpu_shm_allgather.restype = C.c_int
        lib.tagpu_shm_allgather.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_uint64)]
        lib.tagpu_shm_close.argtypes = [C.c_void_p]
        name = [f"tagpu_{os.getpid()}_{uuid.uuid4().hex[:12]}" if self.rank == 0 else None]
        self.dist.broadcast_object_list(name, src=self.dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        shm = lib.tagpu_shm_open(name[0].encode(), self.rank, self.world)
        ok = self.torch.tensor([1 if shm else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()):
            self._shm = shm
        elif shm:
            lib.tagpu_shm_close(shm)

    def plan(self, n_total_bytes: int, k: int):
        """Same n_total_bytes / k on every rank.  Allocates this rank's arena and maps everybody else's."""
        self.t.dist_disconnect()          # a previous plan: unmap the peers, and only free our arena once everybody has
        self.barrier()
        handle = self.t.dist_plan(self.rank, self.world, n_total_bytes, k)
        handles = exchange_bytes(self.dist, handle, self.world, self.group)
        self.t.dist_connect(handles)
        self.barrier()

    def barrier(self):
        # A 4-byte all-reduce followed by a host wait: a true barrier whatever stream the library launches on (the
        # tagpu_dist_* phases synchronise their own stream before returning, so "every rank called barrier()" means
        # "every rank's previous phase is complete and visible").
        if self._shm:
            self._sync()
            self._lib.tagpu_shm_barrier(self._shm)
            return
        self.dist.all_reduce(self._flag, group=self.group)
        self._sync()

    def build(self, ptr: int, n_local_bytes: int, with_graph: bool = True, host: bool = False, gather_solid: bool = True,
              packed: bool = False) -> dict:
        """One pass of the hot path over this rank's slice of the reads (device address, or pinned host address with
        host=True); returns the GLOBAL stats on every rank.  gather_solid=False leaves the solid (k+1)-mers sharded over
        their owner ranks when the two-level graph stage runs (only the contracted paths travel)."""
        t = self.t
        if self._shm and hasattr(t, "dist_step"):
            # the whole step in one C call (tagpu_dist_step): barriers and counter exchanges over the shared-memory segment
            kind = 2 if (host and packed) else 1 if host else 0
            flags = (1 if with_graph else 0) | (2 if gather_solid else 0) | (4 if self._dirty else 0)
            st, used_paths = t.dist_step(self._shm, ptr, n_local_bytes, kind, flags)
            self._dirty = bool(used_paths)
            return st
        if self._dirty:
            self.barrier()
            self._dirty = False
        if host and packed:           # ptr: this rank's slice as a packed read stream of n_local_bytes positions
            t.dist_partition_host_packed(ptr, n_local_bytes)
        elif host:
            t.dist_partition_host(ptr, n_local_bytes)
        else:
            t.dist_partition(ptr, n_local_bytes)
        self.barrier()
        local = t.dist_count()
        all_stats = self._gather(local)
        if with_graph and t.contract:
            # level 1 on every rank's own solid set; the all-gather of the 4 values is the barrier before the pull
            all_paths = self._gather(t.dist_contract())
            if os.environ.get("TAGPU_DIST_DEBUG") and self.rank == 0:
                print("tagpu dist: stats", all_stats, "paths", all_paths, flush=True)
            if all(all_paths[4 * r + 3] for r in range(self.world)):
                self._dirty = True
                return t.dist_graph_paths(all_stats, all_paths, gather_solid)
        return t.dist_graph(all_stats, with_graph)

    def _gather(self, local):
        """all-gather of 4 counters per rank -> flat list in rank order; doubles as a barrier"""
        if self._shm:
            import ctypes as C
            mine = (C.c_uint64 * 4)(*[int(v) for v in local])
            out = (C.c_uint64 * (4 * self.world))()
            if self._lib.tagpu_shm_allgather(self._shm, mine, 4, out) != 0:
                raise RuntimeError("tagpu_shm_allgather failed")
            return list(out)
        return gather_stats(self.dist, self._stats, self._all, local, self.group)

    def _sync(self):
        if self.on_gpu:
            self.torch.cuda.current_stream().synchronize()

    def close(self):
        self._dirty = False
        self.t.dist_disconnect()
        self.barrier()
        self.t.dist_close()
        if self._shm:
            self.barrier()
            self._lib.tagpu_shm_close(self._shm)
            self._shm = None


def exchange_bytes(dist, blob: bytes, world: int, group=None):
    """all-gather of one fixed-size byte string per rank (the 64-byte CUDA IPC handles)."""
    out = [None] * world
    dist.all_gather_object(out, blob, group=group)
    return out


def gather_stats(dist, buf, all_buf, local, group=None):
    """all-gather of the 4 per-rank counters (n_instances, n_distinct, n_solid, sum_solid) -> flat list, rank order.
    The collective doubles as the barrier between counting and the solid-set gather."""
    import torch
    buf.copy_(torch.tensor([int(v) for v in local], dtype=torch.int64))
    dist.all_gather_into_tensor(all_buf, buf, group=group) if buf.is_cuda else _gather_cpu(dist, all_buf, buf, group)
    return [int(v) for v in all_buf.cpu().tolist()]


def _gather_cpu(dist, all_buf, buf, group):
    parts = list(all_buf.view(-1, buf.numel()).unbind(0))
    dist.all_gather(parts, buf, group=group)


def sum_stats(all_stats, world: int):
    a = np.asarray(all_stats, dtype=np.uint64).reshape(world, 4)
    return dict(zip(("n_instances", "n_distinct", "n_solid", "sum_solid"), (int(x) for x in a.sum(axis=0))))
