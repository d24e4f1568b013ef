#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: k-mers/sec counted + graph built; HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C1|small]

One "step" = one pass of the whole path (count -> solid set -> edge masks -> unitig graph, flat arrays in HBM) over one
batch of synthetic reads.  At N = 1 the workload is BASELINE.json configs[1] ("C2": E. coli-scale, 2 M pairs x 151 bp,
k0 = 45, 128-bit keys).  `value` is measured with the (ASCII) read stream already resident in HBM; `e2e` is the same metric
through the host-buffer C-ABI call (pinned host stream -> H2D -> build -> stats back); its host buffer is the packed read
stream of include/tagpu.h by default (`--host-format ascii` for the byte-per-base stream).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: genome_len, n_pairs, k   ("small" is for quick checks only)
    "C2": dict(genome_len=4_641_652, n_pairs=2_000_000, k=45, seed=1),
    "C1": dict(genome_len=4_641_652, n_pairs=2_000_000, k=31, seed=1),
    "small": dict(genome_len=400_000, n_pairs=150_000, k=45, seed=1),
    # C4's coverage (42x) and repeat density on a genome that fits one GPU: developer runs of the C4 regime
    "C4s": dict(genome_len=25_000_000, n_pairs=3_500_000, k=45, seed=4, n_repeats=2_000),
    # BASELINE.json configs[2]: metagenomic mock community, ~20 M pairs, uneven coverage, k0 = 45 (multi-GPU sized: the
    # 6 GB read stream needs >= 2 B200s with today's region sizing)
    "C3": dict(n_genomes=20, n_pairs=20_000_000, k=45, seed=3, genome_len=90_000_000),
    # BASELINE.json configs[3]: human chr1-scale, ~250 Mbp, 40x, ~35 M pairs, k0 = 45, hash-partitioned over 8 B200s
    # (planted repeats stand in for the repeat families of the spec)
    "C4": dict(genome_len=250_000_000, n_pairs=35_000_000, k=45, seed=4, n_repeats=20_000),
}
L = 151
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the last `ncu --set full` capture of the C2 workload
# (profiles/), keyed by kernel; None until such a capture exists for the current kernels
TRAFFIC = {"k_count_buckets<W>": 0.973014e9 + 0.140346e9, "k_partition<W>": 0.610491e9 + 0.909045e9}  # profiles/r1_ncu_top_kernels_raw.txt


N_CHUNKS = 16   # the read set is generated in 16 independently seeded chunks of pairs, so a rank can make just its share


def make_genomes(torch, wl, device):
    """Concatenated synthetic genome(s) of a workload + (start, length, weight) of each replicon.  Uniform bases plus
    planted 600 bp repeats so that branching exists; C3 = 20 genomes, log-uniform lengths, log-normal
    abundances (a metagenomic mock community)."""
    g = torch.Generator(device=device)
    g.manual_seed(wl["seed"])
    if wl.get("n_genomes", 1) > 1:
        cpu = torch.Generator().manual_seed(wl["seed"])
        lens = torch.exp(torch.rand(wl["n_genomes"], generator=cpu) * (torch.log(torch.tensor(8e6)) - torch.log(torch.tensor(1e6))) +
                         torch.log(torch.tensor(1e6))).long()
        abund = torch.exp(torch.randn(wl["n_genomes"], generator=cpu) * 1.5)
    else:
        lens = torch.tensor([wl["genome_len"]])
        abund = torch.ones(1)
    total = int(lens.sum())
    genome = torch.randint(0, 4, (total,), generator=g, device=device, dtype=torch.uint8)
    n_rep = wl.get("n_repeats", 40) * len(lens)
    src = torch.randint(0, total - 600, (n_rep,), generator=g, device=device)
    dst = torch.randint(0, total - 600, (n_rep,), generator=g, device=device)
    for a, b in zip(src.tolist(), dst.tolist()):
        genome[b:b + 600] = genome[a:a + 600].clone()
    starts = torch.cumsum(lens, 0) - lens
    weight = (abund * lens.double()).double()
    return genome, starts.to(device), lens.to(device), (weight / weight.sum()).to(device)


def gen_reads_gpu(torch, wl, device, chunks=None, sub_err=0.005, n_rate=0.02):
    """Synthetic paired reads generated on the device (SURVEY.md §8d shape: insert ~U[300,500], 0.5 % substitutions, 2 % of
    reads carry one N).  Returns a uint8 tensor of reads, each followed by a newline: for every chunk of pairs its R1 reads,
    then its R2 reads.  `chunks` = range of chunk ids to generate (default: all N_CHUNKS); chunk c is seeded by
    (seed, c) alone, so every rank of a multi-GPU run produces exactly its slice of the same read set."""
    genome, starts, lens, weight = make_genomes(torch, wl, device)
    n_pairs = wl["n_pairs"]
    chunks = range(N_CHUNKS) if chunks is None else chunks
    lut = torch.tensor(list(b"ACGT"), device=device, dtype=torch.uint8)
    idx = torch.arange(L, device=device)[None, :]
    parts = []
    for c in chunks:
        p0, p1 = n_pairs * c // N_CHUNKS, n_pairs * (c + 1) // N_CHUNKS
        g = torch.Generator(device=device)
        g.manual_seed(wl["seed"] * 1000003 + c + 1)
        for s in range(p0, p1, 250_000):
            m = min(250_000, p1 - s)
            which = torch.multinomial(weight, m, replacement=True, generator=g) if len(lens) > 1 else torch.zeros(m, dtype=torch.long, device=device)
            ins = torch.randint(300, 501, (m,), generator=g, device=device)
            pos = starts[which] + (torch.rand(m, generator=g, device=device, dtype=torch.float64) * (lens[which] - ins)).long()
            flip = torch.rand(m, generator=g, device=device) < 0.5
            fwd = genome[pos[:, None] + idx]
            rev = 3 - genome[(pos + ins - 1)[:, None] - idx]
            out = torch.empty((2, m, L + 1), device=device, dtype=torch.uint8)
            for mate, codes in ((0, torch.where(flip[:, None], rev, fwd)), (1, torch.where(flip[:, None], fwd, rev))):
                err = torch.rand((m, L), generator=g, device=device) < sub_err
                rnd = torch.randint(0, 4, (m, L), generator=g, device=device, dtype=torch.uint8)
                codes = torch.where(err, rnd, codes)
                chars = lut[codes.long()]
                with_n = torch.rand(m, generator=g, device=device) < n_rate
                n_pos = torch.randint(0, L, (m,), generator=g, device=device)
                rows = torch.nonzero(with_n).squeeze(1)
                chars[rows, n_pos[rows]] = ord("N")
                out[mate, :, :L] = chars
            out[:, :, L] = ord("\n")
            parts.append(out.reshape(-1))
    return torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device=device)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def algorithmic_bytes(st, k, n_stream):
    """Bytes = N_i (B_in + W + 8) + N_distinct (W + 4) + N_solid (3W + 28) + N_kmer (W + 33.5)."""
    K = k + 1
    W = 8 if K <= 32 else 16
    b_in = n_stream / max(st["n_instances"], 1)          # ASCII bytes actually read per window
    count = st["n_instances"] * (b_in + W + 8) + st["n_distinct"] * (W + 4)
    graph = st["n_solid"] * (3 * W + 28) + st["n_kmers"] * (W + 33.5)
    return count, graph


def files_e2e(torch, h_stream, k, reps=3):
    """The reference-facing call on FILES: build_graph_from_scratch(k, n_threads, mmem, 1, &R1.fq, &R2.fq, dir, &g)
    (/root/reference/src/kmer_build.h:17-19) on FASTQ files in a RAM disk -> the caller's struct asm_graph_t in host memory
    (SURVEY.md §8d T_e2e: parse FASTQ, upload, count, build, copy back, one malloc per node / edge).  Informative only."""
    import numpy as np
    from turingassembler_b200 import build_graph_from_scratch
    a = h_stream.numpy().reshape(-1, L + 1)
    n = a.shape[0] // 2
    td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    paths = []
    try:
        for mate, block in ((1, a[:n]), (2, a[n:])):
            rec = np.empty((n, 12 + L + 1 + 2 + L + 1), np.uint8)
            ids = np.arange(n)
            rec[:, 0] = ord("@")
            for d in range(9):
                rec[:, 1 + d] = ord("0") + (ids // 10 ** (8 - d)) % 10
            rec[:, 10] = ord("/"); rec[:, 11] = ord("0") + mate
            rec[:, 12:12 + L + 1] = block                      # sequence + newline
            o = 12 + L + 1
            rec[:, o] = ord("+"); rec[:, o + 1] = 10
            rec[:, o + 2:o + 2 + L] = ord("I"); rec[:, o + 2 + L] = 10
            p = os.path.join(td, f"R{mate}.fq")
            # header line needs its own newline: "@000000001/1\n"
            hdr = rec[:, :12]
            with open(p, "wb") as f:
                f.write(np.concatenate([hdr, np.full((n, 1), 10, np.uint8), rec[:, 12:]], axis=1).tobytes())
            paths.append(p)
        threads = os.cpu_count() or 4
        times = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            g = build_graph_from_scratch(k, threads, 32, [paths[0]], [paths[1]], td)
            times.append(time.perf_counter() - t0)
        sec = sum(times[1:]) / reps
        return {"ms_per_step": sec * 1e3, "n_e": int(g.n_e), "threads": threads,
                "what": "build_graph_from_scratch on 2 FASTQ files (RAM disk) -> struct asm_graph_t in host memory"}
    finally:
        for p in paths:
            os.remove(p)
        os.rmdir(td)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_cpu(workload, steps, warmup):
    """Times the reference's own CPU implementation of the path on this box's host cores, all threads, on a BOUNDED sample
    of the workload (same read length, error model and coverage; a fifth of the genome and of the reads):
    oracle/_ref/TA_ref build_0 = the unmodified reference sources + oracle/kmc_cpu.c standing in for the absent libkmc.a
    (kind "reference"); the oracle port's own driver if that binary was not built (kind "port").
    FASTQ files on a RAM disk -> graph_k_<k>_level_0.bin, i.e. the reference's whole stage including its file I/O."""
    import _oracle
    import _reads
    wl = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    n_pairs = min(wl["n_pairs"], 400_000)
    genome_len = max(wl["genome_len"] * n_pairs // wl["n_pairs"], 20_000)
    stream = _reads.gen_stream(genome_len, n_pairs, seed=wl["seed"], n_repeats=8)
    reads = stream.reshape(-1, L + 1)[:, :L]
    ora = _oracle.load()
    n_inst = ora.count(stream, wl["k"] + 1, ci=2, threads=cores)["n_instances"]
    have_ref = os.path.exists(_oracle.TA_REF)
    exe = _oracle.TA_REF if have_ref else os.path.join(ROOT, "oracle", "ta_oracle")
    times = []
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        f1, f2 = os.path.join(td, "R1.fq"), os.path.join(td, "R2.fq")
        q = b"+\n" + b"I" * L + b"\n"
        for path, block, mate in ((f1, reads[:n_pairs], 1), (f2, reads[n_pairs:], 2)):
            with open(path, "wb") as f:
                f.write(b"".join(b"@r%d/%d\n" % (i, mate) + r.tobytes() + b"\n" + q for i, r in enumerate(block)))
        for step in range(warmup + steps):
            out = os.path.join(td, f"out{step}")
            os.makedirs(out)
            cmd = [exe, "build_0", "-1", f1, "-2", f2, "-l", "ust", "-k0", str(wl["k"]), "-t", str(cores), "-o", out]
            t0 = time.perf_counter()
            p = subprocess.run(cmd, capture_output=True, text=True)
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                sys.stderr.write((p.stdout + p.stderr)[-2000:])
                raise SystemExit(1)
            subprocess.run(["rm", "-rf", out])
            if step >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    sample = (f"{2 * n_pairs} reads x {L} bp from a {genome_len} bp genome ({n_inst} (k+1)-mer instances), "
              f"FASTQ files -> graph_k_{wl['k']}_level_0.bin via build_0 -t {cores}")
    return {"value": n_inst / sec, "unit": "kmers/s", "cores": cores, "kind": "reference" if have_ref else "port", "sample": sample}, sec


def run_reference(args):
    """--impl reference: see reference_cpu(); rank 0 alone runs and prints, the other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = WORKLOADS[args.workload]
    base, sec = reference_cpu(args.workload, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": "kmers_per_sec_counted_and_graph_built", "value": base["value"], "unit": "kmers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u128" if wl["k"] + 1 > 32 else "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload} (bounded sample): {base['sample']}"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tagpu", choices=["tagpu", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the informative FASTQ-files end-to-end measurement")
    ap.add_argument("--host-format", default="packed", choices=["ascii", "packed"],
                    help="e2e leg: the pinned host buffer holds the ASCII read stream, or the packed one (include/tagpu.h; packed outside the timed region)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from turingassembler_b200 import Tagpu

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = os.environ.get("TAGPU_NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    wl = WORKLOADS[args.workload]
    k = wl["k"]

    # Strong scaling (north_star): the SAME read set is split over the ranks.  The set is generated in N_CHUNKS
    # independently seeded chunks of pairs; rank r generates (only) chunks [r N_CHUNKS / world, (r + 1) N_CHUNKS / world).
    n_total = 2 * wl["n_pairs"] * (L + 1)
    if world > 1:
        from turingassembler_b200.dist import DistTagpu
        d_stream = gen_reads_gpu(torch, wl, dev, range(N_CHUNKS * rank // world, N_CHUNKS * (rank + 1) // world))
    else:
        d_stream = gen_reads_gpu(torch, wl, dev)
    n_stream = d_stream.numel()
    h_stream = torch.empty(n_stream, dtype=torch.uint8, pin_memory=True)
    h_stream.copy_(d_stream)
    torch.cuda.synchronize()
    h_packed = None
    if args.host_format == "packed":
        from turingassembler_b200.api import pack_stream, packed_bytes
        h_packed = torch.empty(packed_bytes(n_stream), dtype=torch.uint8, pin_memory=True)
        pack_stream((h_stream.data_ptr(), n_stream), threads=min(16, os.cpu_count() or 1), out=h_packed.numpy())

    t = Tagpu(local_rank)
    stream = torch.cuda.Stream(device=dev)      # one explicit stream for the library's kernels, the copies and the timing events
    torch.cuda.set_stream(stream)
    t.set_stream(stream.cuda_stream)
    if world > 1:
        dt = DistTagpu(t, rank, world)
        dt.plan(n_total, k)

    def step_device():
        if world > 1:
            return dt.build(d_stream.data_ptr(), n_stream, gather_solid=False)
        return t.build_device(d_stream.data_ptr(), n_stream, k)

    def step_host():
        # end to end: this rank's reads start in pinned HOST memory; H2D copy, build, stats back to the host
        if h_packed is not None:
            if world > 1:
                return dt.build(h_packed.data_ptr(), n_stream, host=True, gather_solid=False, packed=True)
            return t.build_host_packed(h_packed.data_ptr(), n_stream, k)
        if world > 1:
            return dt.build(h_stream.data_ptr(), n_stream, host=True, gather_solid=False)
        return t.build_host((h_stream.data_ptr(), n_stream), k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        st = step_device()
    t.set_profile(True)   # one CUDA-event pair around every kernel launch, on the launching stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_count = ms_graph = 0.0
    launches = 0
    kern = {}
    ev0.record(stream)
    for _ in range(args.steps):
        st = step_device()
        ms_count += st["ms_count"]
        ms_graph += st["ms_graph"]
        launches += st["gpu_launches"]
        for name, v in t.profile().items():
            a = kern.setdefault(name, [0.0, 0])
            a[0] += v["ms"]
            a[1] += v["launches"]
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    t.set_profile(False)
    step_host()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        st_e = step_host()
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join()

    if world > 1:
        ln = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(ln)
        launches = int(ln.item())
    tm = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    h2d_packed = 0
    if h_packed is not None:
        hb = torch.tensor([h_packed.numel()], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(hb)
        h2d_packed = int(hb.item())
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, ms_e2e = tm.tolist()
    if world > 1:
        dt.close()
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    n_inst = st["n_instances"]
    ms_step = ms / args.steps
    value = n_inst / (ms_step * 1e-3)
    b_count, b_graph = algorithmic_bytes(st, k, n_total)
    peak, peak_src = peaks()
    count_ms = ms_count / args.steps
    # dominant kernel = the one with the largest share of the step; its algorithmic bytes:
    #   k_count_buckets: N_i (W + 8) + N_distinct (W + 4)   (key compare + count read/write, first touch per distinct key)
    #   k_partition:     N_i B_in                            (the ASCII stream is read exactly once)
    W = 8 if k + 1 <= 32 else 16
    # (per launch = per rank: 1/world of the instances, distinct keys and stream bytes)
    kbytes = {"k_count_buckets<W>": (st["n_instances"] * (W + 8) + st["n_distinct"] * (W + 4)) / world, "k_partition<W>": float(n_total) / world}
    kernels = {name: {"ms_per_launch": v[0] / max(v[1], 1), "launches_per_step": v[1] / args.steps,
                      "share_of_step": v[0] / args.steps / ms_step} for name, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
    top = next(iter(kernels))
    top_ms = kernels[top]["ms_per_launch"]
    top_bytes = kbytes.get(top, b_count)
    achieved = top_bytes / (top_ms * 1e-3) / 1e9
    line = {
        "metric": "kmers_per_sec_counted_and_graph_built", "value": value, "unit": "kmers/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u128" if k + 1 > 32 else "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl.get('n_genomes', 1)} genome(s), {wl['genome_len']} bp, {wl['n_pairs']} pairs x {L} bp, k0={k} "
                               f"(K={k + 1}), cutoff 2; {n_total} stream bytes resident in HBM (> L2, no flush needed)",
                   "n_instances": n_inst, "n_distinct": st["n_distinct"], "n_solid": st["n_solid"], "n_kmers": st["n_kmers"],
                   "n_v": st["n_v"], "n_e": st["n_e"],
                   "parallelism": (f"{world} ranks: reads split 1/{world} per rank, (k+1)-mer buckets hash-partitioned to owner GPUs "
                                   f"(k_count_buckets reads every rank's records through NVLink peer loads); graph stage two-level: every rank "
                                   f"contracts its own solid (k+1)-mers into unbranched paths, the paths are pulled over NVLink (k_gather_paths), "
                                   f"the path-level global stage runs on every rank; the solid set stays sharded over its owners") if world > 1 else
                                  "1 gpu; graph stage two-level (k_contract inside the bucket groups, then the path-level global stage)"},
        "stage_ms": {"count": count_ms, "graph": ms_graph / args.steps},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": TRAFFIC.get(top), "kernel": top, "ms_per_launch": top_ms,
                     "algorithmic_bytes": top_bytes, "peak_source": peak_src,
                     "count_stage": {"achieved": b_count / (count_ms * 1e-3) / 1e9, "frac": b_count / (count_ms * 1e-3) / 1e9 / peak,
                                     "algorithmic_bytes": b_count},
                     "whole_path": {"achieved": (b_count + b_graph) / (ms_step * 1e-3) / 1e9,
                                    "frac": (b_count + b_graph) / (ms_step * 1e-3) / 1e9 / peak}},
        "e2e": {"value": st_e["n_instances"] / (ms_e2e / args.steps * 1e-3), "unit": "kmers/s",
                "h2d_bytes_per_step": n_total if h_packed is None else h2d_packed, "d2h_bytes_per_step": 8 * 140 * world,
                "ms_per_step": ms_e2e / args.steps, "host_format": args.host_format,
                "n_solid": st_e["n_solid"], "n_e": st_e["n_e"]},
        "gpu_launches": launches,
        "kernels": kernels,
        "clocks": sampler.summary(),
    }
    if world == 1 and not args.no_files:
        fe = files_e2e(torch, h_stream, k)
        fe["value"] = n_inst / (fe["ms_per_step"] * 1e-3)
        fe["unit"] = "kmers/s"
        line["e2e_files"] = fe
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = reference_cpu(args.workload, 2, 1)[0]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
