"""Multi-GPU path.

CPU part (gloo, world_size 2, runs everywhere): the host-side logic of the sharded path — the read-stream split at read
boundaries (tagpu_dist_shard_range in libtagpu.so), owner partitioning of the key space, the stats all-gather of
turingassembler_b200/dist.py — is exact: per-rank shard counts, exchanged to their owners and merged, equal the oracle's
count of the whole stream.  The oracle stands in for the GPU kernels here (test infrastructure; no GPU in the CPU suite).

GPU part (-m gpu, needs >= 2 GPUs, else skipped): tools/dist_check.py under torchrun on 2 ranks, CUDA path vs oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import _oracle
import _reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, stream_bytes, K, ci, out_q):
    import torch
    import torch.distributed as dist
    from turingassembler_b200.api import shard_range
    from turingassembler_b200.dist import exchange_bytes, gather_stats, owner_of_bucket, sum_stats

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stream = np.frombuffer(stream_bytes, dtype=np.uint8)
    ora = _oracle.load()
    b, e = shard_range(stream, rank, world)
    mine = ora.count(stream[b:e].copy(), K, ci=1, threads=2)                  # every (key, count) of this rank's reads
    n_buckets = 1 << 10
    bucket = ((mine["lo"] * np.uint64(0x9E3779B97F4A7C15)) ^ (mine["hi"] * np.uint64(0xD6E8FEB86659FD93))) >> np.uint64(54)
    owner = np.array([owner_of_bucket(int(x), n_buckets, world) for x in bucket], dtype=np.int64) if bucket.size else np.zeros(0, np.int64)
    # "all-to-all": every rank publishes one message per destination
    msgs = []
    for dst in range(world):
        m = owner == dst
        msgs.append((mine["hi"][m].tobytes(), mine["lo"][m].tobytes(), mine["count"][m].tobytes()))
    allmsgs = [None] * world
    dist.all_gather_object(allmsgs, msgs)
    hi = np.concatenate([np.frombuffer(allmsgs[src][rank][0], np.uint64) for src in range(world)])
    lo = np.concatenate([np.frombuffer(allmsgs[src][rank][1], np.uint64) for src in range(world)])
    cnt = np.concatenate([np.frombuffer(allmsgs[src][rank][2], np.uint32) for src in range(world)]).astype(np.uint64)
    order = np.lexsort((lo, hi))
    hi, lo, cnt = hi[order], lo[order], cnt[order]
    if hi.size:
        new = np.ones(hi.size, bool)
        new[1:] = (hi[1:] != hi[:-1]) | (lo[1:] != lo[:-1])
        idx = np.flatnonzero(new)
        total = np.add.reduceat(cnt, idx)
        hi, lo = hi[idx], lo[idx]
    else:
        total = cnt
    solid = total >= ci
    local = [mine["n_instances"], int(hi.size), int(solid.sum()), int(total[solid].sum())]
    buf, all_buf = torch.zeros(4, dtype=torch.int64), torch.zeros(4 * world, dtype=torch.int64)
    all_stats = gather_stats(dist, buf, all_buf, local)
    handles = exchange_bytes(dist, bytes([rank]) * 64, world)
    assert handles == [bytes([r]) * 64 for r in range(world)]
    parts = [None] * world
    dist.all_gather_object(parts, (hi[solid].tobytes(), lo[solid].tobytes(), total[solid].astype(np.uint32).tobytes()))
    if rank == 0:
        out_q.put((sum_stats(all_stats, world), (b, e), parts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("K,world", [(32, 2), (46, 2), (22, 3)])
def test_sharded_count_is_exact_gloo(K, world):
    import torch.multiprocessing as mp
    stream = _reads.gen_stream(30000, 1500, seed=21 + K)
    ora = _oracle.load()
    want = ora.count(stream, K, ci=2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + K + world
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, stream.tobytes(), K, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    tot, (b0, e0), parts = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert b0 == 0 and 0 < e0 < stream.size and stream[e0 - 1] == ord("\n")
    assert tot["n_instances"] == want["n_instances"]
    assert tot["n_distinct"] == want["n_distinct"]
    assert tot["n_solid"] == want["hi"].size
    assert tot["sum_solid"] == int(want["count"].astype(np.uint64).sum())
    hi = np.concatenate([np.frombuffer(p[0], np.uint64) for p in parts])
    lo = np.concatenate([np.frombuffer(p[1], np.uint64) for p in parts])
    cnt = np.concatenate([np.frombuffer(p[2], np.uint32) for p in parts])
    o = np.lexsort((lo, hi))
    assert np.array_equal(hi[o], want["hi"]) and np.array_equal(lo[o], want["lo"]) and np.array_equal(cnt[o], want["count"])


def test_shard_range_covers_stream_once():
    from turingassembler_b200.api import shard_range
    from turingassembler_b200.dist import shard_reads
    rng = np.random.default_rng(3)
    for trial in range(20):
        lens = rng.integers(0, 300, size=int(rng.integers(1, 60)))
        stream = np.frombuffer(b"".join(bytes(rng.choice(list(b"ACGTN"), size=n).astype(np.uint8)) + b"\n" for n in lens), np.uint8)
        if trial % 3 == 0:
            stream = stream[:-1]                      # no trailing newline
        for world in (1, 2, 3, 5, 8):
            cuts = [shard_range(stream, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == stream.size
            for (b0, e0), (b1, e1) in zip(cuts, cuts[1:]):
                assert e0 == b1 and b0 <= e0
            for b, e in cuts[1:]:
                assert b == stream.size or b == 0 or stream[b - 1] == ord("\n")
    for n, world in ((10, 3), (7, 8), (4_000_000, 8)):
        cuts = [shard_reads(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))


class _FakeTagpu:
    """Records the tagpu_dist_* phases DistTagpu drives (no GPU): the protocol of include/tagpu.h "multi-GPU"."""

    def __init__(self, rank, contract, refuse_rank):
        self.rank, self.contract, self.refuse_rank, self.calls = rank, contract, refuse_rank, []

    def dist_disconnect(self):
        self.calls.append("disconnect")

    def dist_plan(self, rank, world, n_total, k):
        self.calls.append("plan")
        return bytes([rank]) * 64

    def dist_connect(self, handles):
        assert handles == [bytes([r]) * 64 for r in range(len(handles))]
        self.calls.append("connect")

    def dist_partition(self, ptr, n):
        self.calls.append("partition")

    def dist_count(self):
        self.calls.append("count")
        return [100 + self.rank, 50 + self.rank, 10 + self.rank, 40 + self.rank]

    def dist_contract(self):
        self.calls.append("contract")
        return [3 + self.rank, 2, 7, 0 if self.rank == self.refuse_rank else 1]

    def dist_graph_paths(self, all_stats, all_paths, gather_solid):
        self.calls.append(("graph_paths", tuple(all_stats), tuple(all_paths), gather_solid))
        return {"n_solid": sum(all_stats[2::4])}

    def dist_graph(self, all_stats, with_graph):
        self.calls.append(("graph", tuple(all_stats), with_graph))
        return {"n_solid": sum(all_stats[2::4])}

    def dist_close(self):
        self.calls.append("close")


def _protocol_worker(rank, world, port, out_q, no_shm=False):
    import torch.distributed as dist
    from turingassembler_b200.dist import DistTagpu

    if no_shm:
        os.environ["TAGPU_NO_SHM"] = "1"      # barriers and counter exchange through torch.distributed collectives
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    logs = {}
    for name, contract, refuse in (("two_level", True, -1), ("refused", True, 1), ("one_level", False, -1)):
        t = _FakeTagpu(rank, contract, refuse)
        d = DistTagpu(t, rank, world)
        assert (d._shm is None) == no_shm
        d.plan(1000, 31)
        st1 = d.build(0, 10, gather_solid=False)
        st2 = d.build(0, 10)
        d.close()
        assert st1["n_solid"] == st2["n_solid"] == sum(10 + r for r in range(world))
        logs[name] = t.calls
    out_q.put((rank, logs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("no_shm", [False, True], ids=["shm_rendezvous", "collectives"])
def test_phase_protocol_gloo(no_shm):
    """Host-side phase order of one multi-GPU build on 2 gloo ranks: the two-level graph stage (contract -> all-gather ->
    pull paths), its collective fallback when any rank could not contract, and the one-level stage — with the barriers and
    the counter exchange over the shared-memory segment of libtagpu.so (tagpu_shm_*) and over torch.distributed."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_protocol_worker, args=(r, world, 29671 + int(no_shm), q, no_shm)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    stats = tuple(v for r in range(world) for v in (100 + r, 50 + r, 10 + r, 40 + r))
    for rank in range(world):
        paths = lambda refuse: tuple(v for r in range(world) for v in (3 + r, 2, 7, 0 if r == refuse else 1))
        head = ["disconnect", "plan", "connect"]
        assert got[rank]["two_level"] == head + [
            "partition", "count", "contract", ("graph_paths", stats, paths(-1), False),
            "partition", "count", "contract", ("graph_paths", stats, paths(-1), True), "disconnect", "close"]
        # one rank refused: EVERY rank falls back to the one-level stage, with the same global stats
        assert got[rank]["refused"] == head + ["partition", "count", "contract", ("graph", stats, True)] * 2 + ["disconnect", "close"]
        assert got[rank]["one_level"] == head + ["partition", "count", ("graph", stats, True)] * 2 + ["disconnect", "close"]


@pytest.mark.gpu
def test_two_rank_parity_on_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-4000:]
    assert p.stdout.count("PARITY") == 12 and "MISMATCH" not in p.stdout     # 4 cases x (two-level, solid sharded, one-level)
