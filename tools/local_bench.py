#!/usr/bin/env python
"""Developer tool: latency of one build_local_assembly_graph call (row f1) on the GPU vs the reference's CPU function.
    python tools/local_bench.py [L1|L2|L3] [reps]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle  # noqa: E402
import _reads  # noqa: E402
from _cases import local_case  # noqa: E402
from turingassembler_b200 import Tagpu  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "L2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ora = _oracle.load()
with tempfile.TemporaryDirectory() as td:
    lc = local_case(ora, name, td)
    t = Tagpu(0)
    for _ in range(3):
        st = t.build_local_host(lc["stream"], lc["lk"], lc["contigs"], lc["covs"])
    calls = []
    for _ in range(reps):
        t0 = time.perf_counter()
        st = t.build_local_host(lc["stream"], lc["lk"], lc["contigs"], lc["covs"])
        calls.append((time.perf_counter() - t0) * 1e3)
    calls.sort()
    gpu_ms = calls[len(calls) // 2]
    print(f"{name}: {len(lc['r1']) * 2} reads, lk={lc['lk']}, contigs {[len(c) for c in lc['contigs']]}: n_instances={st['n_instances']} "
          f"n_solid={st['n_solid']} n_v={st['n_v']} n_e={st['n_e']}")
    print(f"GPU  tagpu_build_local_host: median {gpu_ms:.3f} ms per call [{calls[0]:.3f} .. {calls[-1]:.3f}] (host wall clock, {reps} calls; "
          f"device {st['ms_total']:.3f} ms in the last)")
    exe = os.path.join(os.path.dirname(_oracle.TA_REF), "TA_local_ref")
    if os.path.exists(exe):
        f1, f2 = os.path.join(td, "R1.fq"), os.path.join(td, "R2.fq")
        _reads.write_fastq(f1, lc["r1"], 1)
        _reads.write_fastq(f2, lc["r2"], 2)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            p = subprocess.run([exe, lc["g0_bin"], str(lc["e1"]), str(lc["e2"]), str(lc["lk"]), f1, f2, td, os.path.join(td, "o.bin"), str(os.cpu_count())],
                               capture_output=True, text=True)
            best = min(best, time.perf_counter() - t0)
        print(f"CPU  reference build_local_assembly_graph via TA_local_ref (process start + load g0 + FASTQ + build + save, {os.cpu_count()} threads): {best * 1e3:.1f} ms")

    # ---- many gaps in flight (tagpu_build_local_batch): the caller's loop runs over thousands of gaps.
    # Wall-clock times of sub-millisecond builds on a shared host vary by several x from one repetition to the next
    # (profiles/r2_local_batch_ab.txt), so every configuration is repeated and reported as median [min .. max].
    from turingassembler_b200.api import build_local_batch
    n_jobs = int(os.environ.get("LOCAL_BENCH_JOBS", "1024"))
    n_rep = int(os.environ.get("LOCAL_BENCH_REPS", "7"))
    job = dict(stream=lc["stream"], k=lc["lk"], contigs=lc["contigs"], covs=lc["covs"])
    for n_ctx in (1, 4, 8, 16):
        build_local_batch([job] * min(n_jobs, 4 * n_ctx), n_ctx)          # warm the contexts
        per_gap = []
        for _ in range(n_rep):
            t0 = time.perf_counter()
            stats, _ = build_local_batch([job] * n_jobs, n_ctx)
            per_gap.append((time.perf_counter() - t0) / n_jobs * 1e3)
            assert all(s_["n_e"] == st["n_e"] and s_["n_solid"] == st["n_solid"] for s_ in stats)
        per_gap.sort()
        med = per_gap[len(per_gap) // 2]
        print(f"GPU  tagpu_build_local_batch: {n_jobs} gaps on {n_ctx:2d} contexts, {n_rep} repetitions: median {med:.3f} ms per gap "
              f"[{per_gap[0]:.3f} .. {per_gap[-1]:.3f}], median {1e3 / med:.0f} gaps/s")
