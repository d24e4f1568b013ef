#!/usr/bin/env python
"""Developer tool: per-kernel CUDA-event breakdown of one workload (python tools/prof_run.py [C2|C1|small] [reps])."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from turingassembler_b200 import Tagpu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
d = bench.gen_reads_gpu(torch, wl, dev)
t = Tagpu(0)
for _ in range(2):
    st = t.build_device(d.data_ptr(), d.numel(), wl["k"])
t.set_profile(True)
tot = {}
for _ in range(reps):
    st = t.build_device(d.data_ptr(), d.numel(), wl["k"])
    for k_, v in t.profile().items():
        a = tot.setdefault(k_, [0.0, 0])
        a[0] += v["ms"] / reps
        a[1] = v["launches"]
print(json.dumps({k_: v for k_, v in st.items() if not k_.startswith("ms_")}))
print(f"count {st['ms_count']:.3f} ms  graph {st['ms_graph']:.3f} ms  total {st['ms_total']:.3f} ms (last rep, with event overhead)")
for k_, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:10.4f} ms  x{n:<4d} {k_}")
print(f"{sum(v[0] for v in tot.values()):10.4f} ms  sum of kernels")
