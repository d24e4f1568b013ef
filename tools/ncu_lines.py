#!/usr/bin/env python
"""Developer tool: per-source-line instruction / stall-sample shares from an .ncu-rep (python tools/ncu_lines.py rep [sort=smp|inst])."""
import csv
import subprocess
import sys

rep = sys.argv[1]
sort_by = sys.argv[2] if len(sys.argv) > 2 else "smp"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
g = lambda r, n: r[hdr.index(n)] if n in hdr else "?"
for r in rows[2:]:
    print(g(r, "Kernel Name")[:34], "| dur us", g(r, "gpu__time_duration.sum"), "| warp-inst", g(r, "smsp__inst_executed.sum"),
          "| issue%", g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), "| dram R/W", g(r, "dram__bytes_read.sum"),
          g(r, "dram__bytes_write.sum"), "| smem conflicts", g(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
          "/", g(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
data, stalls = {}, {}
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1].split("(")[0]
        continue
    if r[0] == "Line No":
        h2 = r
        continue
    if r[0] and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        f = lambda x: float(x) if x.replace(".", "").isdigit() else 0
        key = (cur_fn, cur_file.split("/")[-1], int(r[0]))
        d = data.setdefault(key, [0, 0, r[1], 0])
        d[0] += f(r[7]); d[1] += f(r[6]); d[3] += f(r[8])
        st = stalls.setdefault(cur_fn, {})
        for i, name in enumerate(h2):
            if name.startswith("stall_") and "Not Issued" not in name:
                st[name] = st.get(name, 0) + f(r[i])
for fn in sorted(set(k[0] for k in data)):
    items = [(k, v) for k, v in data.items() if k[0] == fn]
    tot = sum(v[0] for _, v in items); tots = sum(v[1] for _, v in items)
    print("=====", fn, "warp-inst", tot, "samples", tots)
    s = stalls[fn]; ssum = sum(s.values()) or 1
    print("   stalls:", ", ".join(f"{k[6:]} {v / ssum * 100:.0f}%" for k, v in sorted(s.items(), key=lambda kv: -kv[1])[:8]))
    idx = 1 if sort_by == "smp" else 0
    for k, v in sorted(items, key=lambda kv: -kv[1][idx])[:26]:
        print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / max(tots, 1) * 100:5.1f}% smp  act {v[3] / max(v[0], 1):4.1f} | {k[1]}:{k[2]:<4d} | {v[2][:92]}")
