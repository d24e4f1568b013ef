#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank takes its slice of a seeded read stream, the ranks count + build through the tagpu_dist_* phases, and rank 0
compares the global result (solid set with counts, k-mer masks, canonical graph) with the CPU oracle.  Exit code 0 = parity."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import _oracle  # noqa: E402
import _reads  # noqa: E402
from turingassembler_b200 import Tagpu  # noqa: E402
from turingassembler_b200.api import shard_range  # noqa: E402
from turingassembler_b200.dist import DistTagpu  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t = Tagpu(local)
    t.set_stream(torch.cuda.current_stream().cuda_stream)
    d = DistTagpu(t, rank, world)
    ok = True
    for genome, pairs, k, seed in ((60000, 6000, 31, 5), (60000, 6000, 45, 6), (300000, 40000, 45, 7), (20000, 300, 21, 8)):
        stream = _reads.gen_stream(genome, pairs, seed=seed)
        b, e = shard_range(stream, rank, world)
        mine = torch.from_numpy(stream[b:e].copy()).cuda()
        d.plan(stream.size, k)
        # two-level graph stage with the solid set gathered / left sharded, then the one-level stage; each twice (the
        # second pass exercises the re-zeroing of the cursors and the reuse of the regions the paths were parked in)
        for mode, contract, gather in (("two-level", True, True), ("two-level, solid sharded", True, False), ("one-level", False, True)):
            t.set_contract(contract)
            for rep in range(2):
                st = d.build(mine.data_ptr(), mine.numel(), gather_solid=gather)
            if rank == 0:
                ora = _oracle.load()
                want = ora.count(stream, k + 1)
                good = st["n_instances"] == want["n_instances"] and st["n_distinct"] == want["n_distinct"] and st["n_solid"] == want["hi"].size
                g = ora.graph(k, want["hi"], want["lo"], want["count"])
                if gather:
                    hi, lo, cnt = t.solid()
                    o = np.lexsort((lo, hi))
                    good = good and np.array_equal(hi[o], want["hi"]) and np.array_equal(lo[o], want["lo"]) and np.array_equal(cnt[o], want["count"])
                    khi, klo, kmask = ora.graph_masks(g)
                    ghi, glo, gmask = t.kmers()
                    o = np.lexsort((glo, ghi))
                    good = good and np.array_equal(ghi[o], khi) and np.array_equal(glo[o], klo) and np.array_equal(gmask[o], kmask)
                else:
                    try:
                        t.solid()
                        good = False if world > 1 else good      # must refuse: the solid set is not on this rank
                    except Exception:
                        pass
                good = good and (st["n_kmers"], st["n_v"], st["n_e"], st["n_kp1_on_edge"]) == (
                    g.contents.n_kmer, g.contents.n_v, g.contents.n_e, g.contents.n_kp1_on_edge)
                tmp = f"/tmp/dist_check_{os.getpid()}"
                ora.save_bin(g, tmp + "_o.bin")
                t.write_graph_bin(tmp + "_g.bin")
                ora.free_graph(g)
                for m in (0, 1):
                    bo, to = _oracle.canon_text(ora, tmp + "_o.bin", m)
                    bg, tg = _oracle.canon_text(ora, tmp + "_g.bin", m)
                    good = good and bo == 0 and bg == 0 and to == tg
                print(f"dist_check world={world} genome={genome} pairs={pairs} k={k} [{mode}]: n_inst={st['n_instances']} n_solid={st['n_solid']} "
                      f"n_v={st['n_v']} n_e={st['n_e']} -> {'PARITY' if good else 'MISMATCH'}", flush=True)
                ok = ok and good
        t.set_contract(True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    d.close()
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
