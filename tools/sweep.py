#!/usr/bin/env python
"""BASELINE.json configs[4] ("C5"): k-mer count + edge-build throughput sweep, k = 21/31/45/63 x read-set sizes, on ONE GPU:
    python tools/sweep.py [--reads 1000000,4000000,16000000] [--ks 21,31,45,63] [--steps 5] > profiles/<name>.jsonl
One JSON line per (k, reads): device-resident step time (CUDA events), k-mers/s, and the whole-path fraction of the HBM
roofline on SURVEY.md §8(d)'s algorithmic bytes.  The genome is C1's (4.64 Mbp), so coverage grows with the read count.
(Multi-GPU points of the sweep: bench.py --workload under torchrun.)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
from turingassembler_b200 import Tagpu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", default="1000000,4000000,16000000")
    ap.add_argument("--ks", default="21,31,45,63")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    t = Tagpu(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    t.set_stream(stream.cuda_stream)
    peak, peak_src = bench.peaks()
    for n_reads in (int(x) for x in args.reads.split(",")):
        wl = dict(bench.WORKLOADS["C1"], n_pairs=n_reads // 2)
        d = bench.gen_reads_gpu(torch, wl, dev)
        for k in (int(x) for x in args.ks.split(",")):
            for _ in range(args.warmup):
                st = t.build_device(d.data_ptr(), d.numel(), k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                st = t.build_device(d.data_ptr(), d.numel(), k)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            b_count, b_graph = bench.algorithmic_bytes(st, k, d.numel())
            print(json.dumps({
                "workload": f"C5: {n_reads} reads x {bench.L} bp from the {wl['genome_len']} bp genome, k0={k}", "k": k, "reads": n_reads,
                "n_gpus": 1, "ms_per_step": ms, "kmers_per_s": st["n_instances"] / (ms * 1e-3), "n_instances": st["n_instances"],
                "n_distinct": st["n_distinct"], "n_solid": st["n_solid"], "n_e": st["n_e"], "stage_ms": {"count": st["ms_count"], "graph": st["ms_graph"]},
                "roofline_frac_whole_path": (b_count + b_graph) / (ms * 1e-3) / 1e9 / peak, "hbm_peak_gbs": peak, "peak_source": peak_src,
                "steps": args.steps, "warmup": args.warmup}), flush=True)
        del d
        torch.cuda.empty_cache()
    t.close()


if __name__ == "__main__":
    main()
