#!/usr/bin/env python
"""Developer tool: distribution of bucket sizes after the partition pass."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from turingassembler_b200 import Tagpu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
wl = bench.WORKLOADS[name]
d = bench.gen_reads_gpu(torch, wl["genome_len"], wl["n_pairs"], wl["seed"], torch.device("cuda", 0))
t = Tagpu(0)
st = t.build_device(d.data_ptr(), d.numel(), wl["k"])
buf = np.zeros(1 << 22, np.uint64)
t.lib.tagpu_debug_bucket_cursors.restype = C.c_uint64
t.lib.tagpu_debug_bucket_cursors.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
n = t.lib.tagpu_debug_bucket_cursors(t.ctx, buf.ctypes.data, buf.size)
cur = buf[:n]
rec = (cur & np.uint64(0xffffffff)).astype(np.int64)
inst = (cur >> np.uint64(32)).astype(np.int64)
nz = rec > 0
print("buckets", n, "non-empty", int(nz.sum()), "records", int(rec.sum()), "instances", int(inst.sum()))
for nm, a in (("records", rec), ("instances", inst)):
    q = np.percentile(a, [0, 1, 10, 50, 90, 99, 99.9, 100])
    print(nm, "mean %.1f" % a.mean(), "pct[0,1,10,50,90,99,99.9,100] =", [int(x) for x in q])
top = np.argsort(-inst)[:10]
print("top buckets by instances:", [(int(b), int(inst[b]), int(rec[b])) for b in top])
print("windows per record: %.2f" % (inst.sum() / max(rec.sum(), 1)))
for thr in (12288, 24576, 49152, 98304):
    print(f"buckets with instances > {thr}: {int((inst > thr).sum())}, holding {inst[inst > thr].sum() / inst.sum():.3f} of all instances")
