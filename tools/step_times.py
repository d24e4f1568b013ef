#!/usr/bin/env python
"""Developer tool: per-step stage times of repeated device-resident builds (python tools/step_times.py [workload] [steps])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from turingassembler_b200 import Tagpu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
wl = bench.WORKLOADS[name]
d = bench.gen_reads_gpu(torch, wl, torch.device("cuda", 0))
t = Tagpu(0)
if os.environ.get("STEP_TIMES_TORCH_STREAM"):
    s = torch.cuda.Stream()
    t.set_stream(s.cuda_stream)
rows = []
for _ in range(steps):
    st = t.build_device(d.data_ptr(), d.numel(), wl["k"])
    rows.append((st["ms_count"], st["ms_graph"], st["ms_total"]))
print(" ".join(f"{g:.2f}" for _, g, _ in rows))
print("median count %.3f graph %.3f total %.3f" % tuple(sorted(r[i] for r in rows)[len(rows) // 2] for i in range(3)))
