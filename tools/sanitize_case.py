#!/usr/bin/env python
"""Developer tool: one small build of every flavour (k=31 / k=45, plain and local-assembly, overflow path) for
compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _reads  # noqa: E402
from turingassembler_b200 import Tagpu  # noqa: E402

t = Tagpu(0)
for k, seed in ((31, 1), (45, 2), (63, 3), (21, 4)):
    s = _reads.gen_stream(40000, 4000, seed=seed)
    st = t.build_host(s, k)
    print(k, st["n_instances"], st["n_solid"], st["n_v"], st["n_e"], flush=True)
contig = bytes(_reads.gen_stream(3000, 1, seed=9)[:140])
st = t.build_local_host(_reads.gen_stream(20000, 1500, seed=5), 31, [contig, contig[20:120]], [12.5, 30.0])
print("local", st["n_solid"], st["n_v"], st["n_e"], flush=True)
t.close()
print("SANITIZE-DONE")
