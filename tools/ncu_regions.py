#!/usr/bin/env python
"""Developer tool: warp-instruction and sample shares of one kernel by source-line range (python tools/ncu_regions.py rep kernel_substr file:lo-hi=name ...)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    rng, name = a.split("=")
    f, r = rng.split(":")
    lo, hi = r.split("-")
    regions.append((f, int(lo), int(hi), name))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
tot = {}
cur_fn = cur_file = None
for r in csv.reader(src.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": continue
    if r[0].isdigit() and len(r) > 8 and r[2] == "-" and kern in (cur_fn or ""):
        f = lambda x: float(x) if x.replace(".", "").isdigit() else 0
        line = int(r[0])
        name = "other:" + cur_file
        for (rf, lo, hi, nm) in regions:
            if rf == cur_file and lo <= line <= hi: name = nm; break
        d = tot.setdefault(name, [0, 0, 0])
        d[0] += f(r[7]); d[1] += f(r[6]); d[2] += f(r[8])
ti = sum(v[0] for v in tot.values()); ts = sum(v[1] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]/ti*100:5.1f}% inst {v[1]/ts*100:5.1f}% smp  act {v[2]/max(v[0],1):4.1f}  {k}")
