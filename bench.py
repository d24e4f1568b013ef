#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: k-mers/sec counted + graph built; HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C1|small]

One "step" = one pass of the whole path (count -> solid set -> edge masks -> unitig graph, flat arrays in HBM) over one
batch of synthetic reads.  At N = 1 the workload is BASELINE.json configs[1] ("C2": E. coli-scale, 2 M pairs x 151 bp,
k0 = 45, 128-bit keys).  `value` is measured with the (ASCII) read stream already resident in HBM.  `e2e` at N = 1 is the
reference-facing call on FILES (build_graph_from_scratch: FASTQ on a RAM disk -> struct asm_graph_t in host memory, wall
clock) — the same input the reference arm gets; `e2e_host_stream` / `e2e_host_stream_packed` are the host-BUFFER legs
(pinned read stream -> H2D -> build -> whole flat graph D2H), which is also what `e2e` is at N > 1.  `--impl reference`
runs the unmodified reference on the same read set.  See DESIGN.md "Measurement".
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
TRAFFIC = {"C2": {"k_count_buckets<W>": 1.370291e9 + 0.143480e9, "k_partition<W>": 0.610127e9 + 0.908919e9,   # profiles/r2_ncu_top_kernels_raw.txt
                  # all kernel launches of one build (profiles/r2_dram_bytes_C2.csv; the 12 memsets of a step, ~0.1 GB, are not in it)
                  "whole_path": 4.311797e9}}


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


def write_fastq_pair(reads, td):
    """Writes a (2 n_pairs, L + 1) uint8 array of reads (R1 block then R2 block, each row = sequence + newline) as the two
    FASTQ files of SURVEY.md §8d (`@<id>/<mate>`, sequence, `+`, L x `I`) into directory td -> [R1.fq, R2.fq]."""
    import numpy as np
    a = reads.reshape(-1, L + 1)
    n = a.shape[0] // 2
    paths = []
    for mate, block in ((1, a[:n]), (2, a[n:])):
        o = 13 + L + 1
        rec = np.empty((n, o + 2 + L + 1), np.uint8)
        ids = np.arange(n)
        rec[:, 0] = ord("@")
        for d in range(9):
            rec[:, 1 + d] = ord("0") + (ids // 10 ** (8 - d)) % 10
        rec[:, 10] = ord("/"); rec[:, 11] = ord("0") + mate; rec[:, 12] = 10
        rec[:, 13:o] = block                                  # sequence + newline
        rec[:, o] = ord("+"); rec[:, o + 1] = 10
        rec[:, o + 2:o + 2 + L] = ord("I"); rec[:, o + 2 + L] = 10
        p = os.path.join(td, f"R{mate}.fq")
        rec.tofile(p)
        paths.append(p)
    return paths


def workload_string(name):
    """One description of a workload, used verbatim as config.workload by BOTH arms (tagpu and --impl reference)."""
    wl = WORKLOADS[name]
    return (f"{name}: {wl.get('n_genomes', 1)} synthetic genome(s), {wl['genome_len']} bp, {wl['n_pairs']} pairs x {L} bp "
            f"(0.5% substitutions, 2% of reads with one N), k0={wl['k']} (K={wl['k'] + 1}), cutoff 2, whole read set")


def make_reads(torch, wl, device, chunks=None):
    """The workload's read set (or the given chunks of it): generated on the GPU when there is one — the same bytes in both
    arms — else with torch's CPU generator (same distribution, other bytes; developer runs without a GPU)."""
    return gen_reads_gpu(torch, wl, device, chunks)


def files_e2e(h_stream, k, steps, warmup=1):
    """End to end through the reference-facing call on FILES: build_graph_from_scratch(k, n_threads, mmem, 1, &R1.fq,
    &R2.fq, dir, &g) (/root/reference/src/kmer_build.h:17-19) on FASTQ files in a RAM disk -> the caller's struct
    asm_graph_t in host memory (SURVEY.md §8d T_e2e: parse FASTQ, H2D, count, build, D2H, one malloc per node / edge).
    Wall clock around every call; the graph of a step is released (tagpu_free_asm_graph) outside the timed region."""
    from turingassembler_b200 import build_graph_from_scratch
    from turingassembler_b200.api import free_asm_graph
    td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    paths = []
    try:
        paths = write_fastq_pair(h_stream.numpy(), td)
        threads = os.cpu_count() or 4
        times = []
        n_e = n_v = seq_bytes = 0
        for _ in range(warmup + steps):
            t0 = time.perf_counter()
            g = build_graph_from_scratch(k, threads, 32, [paths[0]], [paths[1]], td)
            times.append(time.perf_counter() - t0)
            n_e, n_v = int(g.n_e), int(g.n_v)
            free_asm_graph(g)
        sec = sum(times[warmup:]) / steps
        return {"ms_per_step": sec * 1e3, "n_e": n_e, "n_v": n_v, "threads": threads,
                "fastq_bytes": sum(os.path.getsize(p) for p in paths),
                "what": "build_graph_from_scratch on 2 FASTQ files (RAM disk) -> struct asm_graph_t in host memory (wall clock)"}
    finally:
        for p in paths:
            os.remove(p)
        os.rmdir(td)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_cpu(workload, steps, warmup, reads=None):
    """Times the reference's own CPU implementation of the path on this box's host cores, all threads, on the WHOLE workload
    (the same read set as the tagpu arm): oracle/_ref/TA_ref build_0 = the unmodified reference sources + oracle/kmc_cpu.c
    standing in for the absent libkmc.a (kind "reference"); the oracle port's own driver if that binary was not built
    (kind "port").  FASTQ files on a RAM disk -> graph_k_<k>_level_0.bin, i.e. the reference's whole stage, process start
    and file I/O included.  reads: the read set as a uint8 array (generated here when None)."""
    import _oracle
    wl = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    if reads is None:
        import torch
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        reads = make_reads(torch, wl, dev).cpu().numpy()
    ora = _oracle.load()
    n_inst = ora.count(reads, wl["k"] + 1, ci=2, threads=cores)["n_instances"]
    have_ref = os.path.exists(_oracle.TA_REF)
    exe = _oracle.TA_REF if have_ref else os.path.join(ROOT, "oracle", "ta_oracle")
    times = []
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        f1, f2 = write_fastq_pair(reads, td)
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
    sample = (f"the whole workload, {len(times)} run(s): {reads.size // (L + 1)} reads x {L} bp ({n_inst} (k+1)-mer instances), "
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
        "config": {"workload": workload_string(args.workload)},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class HostGraph:
    """Pinned host arrays for tagpu_copy_graph (the flat graph of include/tagpu.h), allocated once with head-room."""

    FIELDS = (("node_mask", "uint8", "n"), ("node_ebase", "int32", "n"), ("e_src", "int32", "e"), ("e_dst", "int32", "e"),
              ("e_rc", "int32", "e"), ("e_len", "int32", "e"), ("e_count", "int64", "e"), ("e_off", "int64", "e"), ("e_seq", "int32", "w"))

    def __init__(self, torch, st):
        from turingassembler_b200.api import FlatGraph
        cap = {"n": st["n_v"] // 2 * 5 // 4 + 1024, "e": st["n_e"] * 5 // 4 + 1024, "w": st["n_seq_words"] * 5 // 4 + 1024}
        self.cap, self.fg, self.t = cap, FlatGraph(), {}
        for name, dt, kind in self.FIELDS:
            self.t[name] = torch.empty(cap[kind], dtype=getattr(torch, dt), pin_memory=True)
            setattr(self.fg, name, self.t[name].data_ptr())

    def fetch(self, tagpu, st):
        assert st["n_v"] // 2 <= self.cap["n"] and st["n_e"] <= self.cap["e"] and st["n_seq_words"] <= self.cap["w"]
        tagpu.copy_graph_into(self.fg)
        return st["n_v"] // 2 * 5 + st["n_e"] * 32 + st["n_seq_words"] * 4          # bytes copied device -> host


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tagpu", choices=["tagpu", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the FASTQ-files end-to-end leg (e2e then falls back to the host-stream leg)")
    ap.add_argument("--one-level", action="store_true", help="one-level graph stage (developer comparison)")
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
        d_stream = make_reads(torch, wl, dev, range(N_CHUNKS * rank // world, N_CHUNKS * (rank + 1) // world))
    else:
        d_stream = make_reads(torch, wl, dev)
    n_stream = d_stream.numel()
    h_stream = torch.empty(n_stream, dtype=torch.uint8, pin_memory=True)
    h_stream.copy_(d_stream)
    torch.cuda.synchronize()
    from turingassembler_b200.api import pack_stream, packed_bytes
    h_packed = torch.empty(packed_bytes(n_stream), dtype=torch.uint8, pin_memory=True)
    pack_stream((h_stream.data_ptr(), n_stream), threads=min(16, os.cpu_count() or 1), out=h_packed.numpy())

    t = Tagpu(local_rank)
    if args.one_level:
        t.set_contract(False)
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

    host_graph = [None]

    def step_host(packed):
        # end to end on host buffers: this rank's reads start in pinned HOST memory (ASCII, or the packed stream of
        # include/tagpu.h); H2D copy, build, and the whole flat graph copied back into pinned host arrays (rank 0)
        if world > 1:
            st = (dt.build(h_packed.data_ptr(), n_stream, host=True, gather_solid=False, packed=True) if packed else
                  dt.build(h_stream.data_ptr(), n_stream, host=True, gather_solid=False))
        else:
            st = t.build_host_packed(h_packed.data_ptr(), n_stream, k) if packed else t.build_host((h_stream.data_ptr(), n_stream), k)
        nb = 0
        if rank == 0:
            if host_graph[0] is None:
                host_graph[0] = HostGraph(torch, st)
            nb = host_graph[0].fetch(t, st)
        return st, nb

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
    # order-independent digest of the result of the last timed step (outside every timed region)
    dg = t.digest()
    if world > 1:
        part = torch.tensor([dg["solid_sum"] - (1 << 64) if dg["solid_sum"] >= (1 << 63) else dg["solid_sum"], dg["solid_n"]], device=dev, dtype=torch.int64)
        xs = torch.tensor([dg["solid_xor"] - (1 << 64) if dg["solid_xor"] >= (1 << 63) else dg["solid_xor"]], device=dev, dtype=torch.int64)
        if not dg["solid_complete"]:
            dist.all_reduce(part)                                        # wrapping int64 sums = sums mod 2^64
            gathered = [torch.zeros_like(xs) for _ in range(world)]
            dist.all_gather(gathered, xs)
            x = 0
            for g_ in gathered:
                x ^= int(g_.item()) & ((1 << 64) - 1)
            dg["solid_sum"], dg["solid_n"], dg["solid_xor"] = int(part[0].item()) & ((1 << 64) - 1), int(part[1].item()), x
    # e2e legs on host buffers (CUDA events on the launching stream, barrier on both sides)
    e2e_stream = {}
    for name, packed in (("ascii", False), ("packed", True)):
        for _ in range(2):
            step_host(packed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            st_e, d2h = step_host(packed)
        e1.record(stream)
        barrier()
        e2e_stream[name] = (e0.elapsed_time(e1), d2h)
    sampler.stop_flag = True
    sampler.join()

    peer_bytes = st.get("n_records_peer", 0) * st.get("record_bytes", 0)      # this rank's k_count_buckets, per launch
    peer_max = peer_sum = peer_bytes
    if world > 1:
        ln = torch.tensor([launches, peer_bytes], device=dev, dtype=torch.int64)
        dist.all_reduce(ln)
        launches, peer_sum = int(ln[0].item()), int(ln[1].item())
        pm = torch.tensor([peer_bytes], device=dev, dtype=torch.int64)
        dist.all_reduce(pm, op=dist.ReduceOp.MAX)
        peer_max = int(pm.item())
    tm = torch.tensor([ms, e2e_stream["ascii"][0], e2e_stream["packed"][0]], device=dev, dtype=torch.float64)
    hb = torch.tensor([h_packed.numel(), n_stream], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(hb)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    h2d_packed, h2d_ascii = (int(x) for x in hb.tolist())
    ms, ms_e2e_ascii, ms_e2e_packed = tm.tolist()
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
    # per-kernel algorithmic bytes, per launch = per rank (1/world of the instances, distinct keys, stream):
    #   k_count_buckets: N_i (W + 8) + N_distinct (W + 4)   (key compare + count read/write, first touch per distinct key)
    #   k_partition:     N_i B_in                            (the ASCII stream is read exactly once)
    W = 8 if k + 1 <= 32 else 16
    kbytes = {"k_count_buckets<W>": (st["n_instances"] * (W + 8) + st["n_distinct"] * (W + 4)) / world, "k_partition<W>": float(n_total) / world}
    kernels = {name: {"ms_per_launch": v[0] / max(v[1], 1), "launches_per_step": v[1] / args.steps,
                      "share_of_step": v[0] / args.steps / ms_step} for name, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
    for name, kb in kbytes.items():
        if name in kernels:
            ach = kb / (kernels[name]["ms_per_launch"] * 1e-3) / 1e9
            kernels[name].update({"algorithmic_bytes": kb, "equivalent_gbs": ach, "equivalent_frac_of_hbm_peak": ach / peak,
                                  "dram_traffic_bytes": TRAFFIC.get(args.workload, {}).get(name) if world == 1 else None})
    top = next(iter(kernels))
    whole = (b_count + b_graph) / (ms_step * 1e-3) / 1e9
    traffic_all = TRAFFIC.get(args.workload, {}).get("whole_path") if world == 1 else None
    wstr = workload_string(args.workload)
    gold_path = os.path.join(ROOT, "tests", "golden", "digest_fullsize.json")
    gold = json.load(open(gold_path)).get(args.workload) if os.path.exists(gold_path) else None
    digest = {f: dg[f] for f in ("solid_sum", "solid_xor", "solid_n", "edge_sum", "edge_xor", "edge_len_sum", "edge_count_sum", "n_e")}
    digest["matches_reference_golden"] = (all(digest[f] == gold[f] for f in gold) if gold else None)
    line = {
        "metric": "kmers_per_sec_counted_and_graph_built", "value": value, "unit": "kmers/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u128" if k + 1 > 32 else "u64", "data": "synthetic",
        "config": {"workload": wstr},
        "timed_region": f"{n_total} ASCII stream bytes resident in HBM (> L2, no flush needed) -> flat graph arrays in HBM",
        "result": {"n_instances": n_inst, "n_distinct": st["n_distinct"], "n_solid": st["n_solid"], "n_kmers": st["n_kmers"],
                   "n_v": st["n_v"], "n_e": st["n_e"], "digest": digest},
        "parallelism": (f"{world} ranks: reads split 1/{world} per rank, (k+1)-mer buckets hash-partitioned to owner GPUs "
                        f"(k_count_buckets reads every rank's records through NVLink peer loads); graph stage two-level: every rank "
                        f"contracts its own solid (k+1)-mers into unbranched paths, the paths are pulled over NVLink (k_gather_paths), "
                        f"the path-level global stage runs on every rank; the solid set stays sharded over its owners") if world > 1 else
                       "1 gpu; graph stage two-level (k_contract inside the bucket groups, then the path-level global stage)",
        "stage_ms": {"count": count_ms, "graph": ms_graph / args.steps},
        # the fraction is taken on the WHOLE path (T_core), algorithmic bytes / step time / HBM peak
        # (N GPUs: the whole job's bytes against N x the per-GPU peak)
        "roofline": {"bound": "hbm", "achieved": whole, "peak": peak * world, "unit": "GB/s", "frac": whole / (peak * world),
                     "traffic": traffic_all, "scope": "whole path (count + graph), SURVEY.md §8d T_core"
                                                      + (f"; aggregate of {world} GPUs against {world} x the per-GPU peak" if world > 1 else ""),
                     "algorithmic_bytes": b_count + b_graph, "peak_source": peak_src, "dominant_kernel": top,
                     "count_stage": {"achieved": b_count / (count_ms * 1e-3) / 1e9, "frac": b_count / (count_ms * 1e-3) / 1e9 / (peak * world),
                                     "algorithmic_bytes": b_count}},
        "gpu_launches": launches,
        "kernels": kernels,
        "clocks": sampler.summary(),
    }
    e_ascii = {"value": n_inst / (ms_e2e_ascii / args.steps * 1e-3), "unit": "kmers/s", "h2d_bytes_per_step": h2d_ascii,
               "d2h_bytes_per_step": e2e_stream["ascii"][1], "ms_per_step": ms_e2e_ascii / args.steps,
               "what": "pinned host ASCII read stream -> H2D -> build -> whole flat graph D2H into pinned host arrays (CUDA events)"}
    e_packed = {"value": n_inst / (ms_e2e_packed / args.steps * 1e-3), "unit": "kmers/s", "h2d_bytes_per_step": h2d_packed,
                "d2h_bytes_per_step": e2e_stream["packed"][1], "ms_per_step": ms_e2e_packed / args.steps,
                "what": "same from the PACKED host stream (include/tagpu.h; 2-bit packing done by the ingest side, outside the timed region)"}
    if world > 1:
        # NVLink: what the exchange moves is counted by the kernels themselves (records pass 2 reads out of the OTHER ranks'
        # regions x record size; `nvidia-smi nvlink -gt d` reports N/A on this pod, and ncu is a single-GPU tool here)
        t_cb = kernels.get("k_count_buckets<W>", {}).get("ms_per_launch")
        t_gp = kernels.get("k_gather_paths<W>", {}).get("ms_per_launch")
        path_bytes = st["n_solid"] * 0   # (filled below when the path counters are available)
        line["nvlink"] = {"count_stage_bytes_pulled_per_step_all_gpus": peer_sum, "count_stage_bytes_pulled_max_gpu": peer_max,
                          "rx_gbs_max_gpu_during_k_count_buckets": (peer_max / (t_cb * 1e-3) / 1e9) if t_cb else None,
                          "k_gather_paths_ms": t_gp, "peak_gbs_per_direction": 900.0,
                          "frac_of_peak_during_k_count_buckets": (peer_max / (t_cb * 1e-3) / 1e9 / 900.0) if t_cb else None,
                          "source": "bytes counted by the kernels (peer records x record size); nvidia-smi nvlink counters read N/A on this pod"}
        del path_bytes
    line["e2e_host_stream"] = e_ascii
    line["e2e_host_stream_packed"] = e_packed
    if world == 1 and not args.no_files:
        fe = files_e2e(h_stream, k, min(args.steps, 10))
        line["e2e"] = {"value": n_inst / (fe["ms_per_step"] * 1e-3), "unit": "kmers/s", "h2d_bytes_per_step": n_stream,
                       "d2h_bytes_per_step": e2e_stream["ascii"][1], "ms_per_step": fe["ms_per_step"], "steps": min(args.steps, 10),
                       "what": fe["what"], "threads": fe["threads"], "fastq_bytes": fe["fastq_bytes"], "n_e": fe["n_e"],
                       "same_boundary_as_reference_arm": "input yes (the same FASTQ files); output: struct asm_graph_t in memory, the reference arm also writes the .bin"}
    else:
        line["e2e"] = dict(e_ascii)     # multi-GPU: no files entry point; reads start in each rank's pinned host memory
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = reference_cpu(args.workload, 1, 0, h_stream.numpy())[0]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
