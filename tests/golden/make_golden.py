#!/usr/bin/env python
"""Generates tests/golden/golden.json (+ the small canonical unitig files) by running the UNMODIFIED reference
(oracle/_ref/TA_ref = reference sources + oracle/kmc_cpu.c for the absent libkmc.a) on seeded synthetic reads.

Run in the build container only (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
The fixtures travel to the GPU box; the reference does not have to.
"""
import gzip
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle  # noqa: E402
import _reads  # noqa: E402

from _cases import CASES, reads_for  # noqa: E402


def main():
    ora = _oracle.load()
    assert os.path.exists(_oracle.TA_REF), "build oracle/_ref/TA_ref first: make -C oracle ref"
    golden = {}
    for name, (kind, kw, ks) in CASES.items():
        r1, r2 = reads_for(kind, kw)
        with tempfile.TemporaryDirectory() as td:
            f1, f2 = os.path.join(td, "R1.fq"), os.path.join(td, "R2.fq")
            _reads.write_fastq(f1, r1, 1)
            _reads.write_fastq(f2, r2, 2)
            fq_md5 = [hashlib.md5(open(f, "rb").read()).hexdigest() for f in (f1, f2)]
            for k in ks:
                out = os.path.join(td, f"out{k}")
                os.makedirs(out)
                p = subprocess.run([_oracle.TA_REF, "build_0", "-1", f1, "-2", f2, "-l", "ust", "-k0", str(k), "-t", "1", "-o", out],
                                   capture_output=True, text=True)
                log = p.stdout + p.stderr
                assert p.returncode == 0, log[-2000:]
                g = lambda pat: int(re.search(pat, log).group(1))
                rec = dict(
                    k=k, fastq_md5=fq_md5, n_reads=len(r1) + len(r2),
                    n_kmers=g(r"Number of kmer: (\d+)"), n_v=g(r"Number of nodes: (\d+)"), n_e=g(r"Number of edges: (\d+)"),
                    n_kp1_on_edge=g(r"\(k\+1\)-mer on edge: (\d+)"), sum_count=g(r"sum_count = (\d+)"),
                )
                m = re.search(r"\[oracle-kmc\] K=\d+ instances=(\d+) distinct=(\d+) solid=(\d+)", log)
                rec.update(n_instances=int(m.group(1)), n_distinct=int(m.group(2)), n_solid=int(m.group(3)))
                binp = os.path.join(out, f"graph_k_{k}_level_0.bin")
                for mode in (0, 1):
                    bad, txt = _oracle.canon_text(ora, binp, mode)
                    assert bad == 0
                    rec[f"canon{mode}_md5"] = hashlib.md5(txt).hexdigest()
                    if mode == 0 and len(txt) < 200000:
                        with gzip.GzipFile(os.path.join(HERE, f"{name}_k{k}_canon0.txt.gz"), "wb", mtime=0) as gz:
                            gz.write(txt)
                golden[f"{name}_k{k}"] = dict(case=name, gen=kind, gen_kwargs=kw, **rec)
                print(name, k, rec)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    make_local_golden(ora)
    make_contig_golden(ora)


def make_contig_golden(ora):
    """Contig-file mode of build_graph_from_scratch (n_files < 0): the UNMODIFIED reference function, reached through
    oracle/contig_ref_main.c (oracle/_ref/TA_contig_ref), on the cases of tests/_cases.py:CONTIG_CASES
    -> tests/golden/golden_contig.json (with and without the count pass)."""
    from _cases import CONTIG_CASES, contig_case, write_fasta
    exe = os.path.join(os.path.dirname(_oracle.TA_REF), "TA_contig_ref")
    assert os.path.exists(exe), "build oracle/_ref/TA_contig_ref first: make -C oracle ref"
    out = {}
    for name in CONTIG_CASES:
        c = contig_case(name)
        with tempfile.TemporaryDirectory() as td:
            f1, f2, fc = (os.path.join(td, x) for x in ("R1.fq", "R2.fq", "contigs.fa"))
            _reads.write_fastq(f1, c["r1"], 1)
            _reads.write_fastq(f2, c["r2"], 2)
            write_fasta(fc, c["contigs"])
            rec = dict(k=c["k"], contig_len=[len(x) for x in c["contigs"]])
            for without in (0, 1):
                binp = os.path.join(td, f"contig{without}.bin")
                p = subprocess.run([exe, str(c["k"]), f1, f2, fc, td, binp, "1", str(without)], capture_output=True, text=True)
                log = p.stdout + p.stderr
                assert p.returncode == 0, log[-2000:]
                g = lambda pat: int(re.search(pat, log).group(1))
                tag = "nocount_" if without else ""
                rec.update({tag + "n_kmers": g(r"Number of kmer: (\d+)"), tag + "n_v": g(r"Number of nodes: (\d+)"),
                            tag + "n_e": g(r"Number of edges: (\d+)")})
                if not without:
                    rec["n_kp1_on_edge"] = g(r"\(k\+1\)-mer on edge: (\d+)")
                for mode in (0, 1):
                    bad, txt = _oracle.canon_text(ora, binp, mode)
                    assert bad == 0
                    rec[f"{tag}canon{mode}_md5"] = hashlib.md5(txt).hexdigest()
            out[name] = rec
            print(name, rec)
    with open(os.path.join(HERE, "golden_contig.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def make_local_golden(ora):
    """build_local_assembly_graph (row f1): the UNMODIFIED reference function, reached through oracle/local_ref_main.c
    (oracle/_ref/TA_local_ref), on the cases of tests/_cases.py:LOCAL_CASES -> tests/golden/golden_local.json."""
    from _cases import LOCAL_CASES, local_case
    exe = os.path.join(os.path.dirname(_oracle.TA_REF), "TA_local_ref")
    assert os.path.exists(exe), "build oracle/_ref/TA_local_ref first: make -C oracle ref"
    out = {}
    for name in LOCAL_CASES:
        with tempfile.TemporaryDirectory() as td:
            lc = local_case(ora, name, td)
            f1, f2 = os.path.join(td, "R1.fq"), os.path.join(td, "R2.fq")
            _reads.write_fastq(f1, lc["r1"], 1)
            _reads.write_fastq(f2, lc["r2"], 2)
            binp = os.path.join(td, "local.bin")
            p = subprocess.run([exe, lc["g0_bin"], str(lc["e1"]), str(lc["e2"]), str(lc["lk"]), f1, f2, td, binp, "1"],
                               capture_output=True, text=True)
            log = p.stdout + p.stderr
            assert p.returncode == 0, log[-2000:]
            g = lambda pat: int(re.search(pat, log).group(1))
            rec = dict(lk=lc["lk"], e1=lc["e1"], e2=lc["e2"], contig_len=[len(c) for c in lc["contigs"]], covs=lc["covs"],
                       n_kmers=g(r"Number of kmer: (\d+)"), n_v=g(r"Number of nodes: (\d+)"), n_e=g(r"Number of edges: (\d+)"),
                       n_kp1_on_edge=g(r"\(k\+1\)-mer on edge: (\d+)"), sum_count=g(r"sum_count = (\d+)"))
            for mode in (0, 1):
                bad, txt = _oracle.canon_text(ora, binp, mode)
                assert bad == 0
                rec[f"canon{mode}_md5"] = hashlib.md5(txt).hexdigest()
            out[name] = rec
            print(name, rec)
    with open(os.path.join(HERE, "golden_local.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def coverage_lines(flat, counts):
    """One line per edge, `sequence<TAB>count`, sorted: the numbering-independent form of a coverage recount."""
    import numpy as np
    out = []
    for e in range(flat["n_e"]):
        n, o = int(flat["e_len"][e]), int(flat["e_off"][e])
        w = flat["e_seq"][o:o + ((n + 15) >> 4)].astype(np.uint64)
        seq = "".join("ACGT"[(int(w[i >> 4]) >> ((i & 15) << 1)) & 3] for i in range(n))
        out.append(f"{seq}\t{int(counts[e])}")
    return "\n".join(sorted(out)).encode()


def make_coverage_golden():
    """tests/golden/golden_coverage.json: the UNMODIFIED reference's `build_coverage` (kmer_count_on_edges + add_cnt_to_graph,
    /root/reference/src/coverage/kmer_count.c) on its own level-0 graph of the seeded cases -> md5 of the sorted
    `sequence<TAB>count` lines, edge count and count sum.  Pins oracle/cov_oracle.c (CPU test) and the GPU path."""
    import numpy as np
    out = {}
    for name, ks in (("P1", [31, 45]), ("M2_lowcov", [25])):
        kind, kw, _ = CASES[name]
        r1, r2 = reads_for(kind, kw)
        with tempfile.TemporaryDirectory() as td:
            f1, f2 = os.path.join(td, "R1.fq"), os.path.join(td, "R2.fq")
            _reads.write_fastq(f1, r1, 1)
            _reads.write_fastq(f2, r2, 2)
            for k in ks:
                o0, oc = os.path.join(td, f"l0_{k}"), os.path.join(td, f"cov_{k}")
                os.makedirs(o0)
                os.makedirs(oc)
                for cmd in ([_oracle.TA_REF, "build_0", "-1", f1, "-2", f2, "-l", "ust", "-k0", str(k), "-t", "1", "-o", o0],
                            [_oracle.TA_REF, "build_coverage", "-i", os.path.join(o0, f"graph_k_{k}_level_0.bin"), "-1", f1, "-2", f2, "-l", "ust",
                             "-t", "2", "-o", oc]):
                    p = subprocess.run(cmd, capture_output=True, text=True)
                    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
                flat = _oracle.load_bin_flat(os.path.join(oc, f"graph_k_{k}_coverage_built.bin"))
                lines = coverage_lines(flat, flat["e_count"])
                out[f"{name}_k{k}"] = dict(case=name, k=k, n_e=int(flat["n_e"]), count_sum=int(flat["e_count"].sum(dtype=np.uint64)),
                                           lines_md5=hashlib.md5(lines).hexdigest())
                print(name, k, out[f"{name}_k{k}"])
    with open(os.path.join(HERE, "golden_coverage.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def make_struct_layout():
    """tests/golden/struct_layout.txt: sizeof / offsetof of the reference's graph structs, printed by a C program compiled
    against the reference's own headers (tests/test_abi.py compares include/tagpu_graph.h with it, and never rewrites it)."""
    import pathlib
    import test_abi
    with tempfile.TemporaryDirectory() as td:
        ref = test_abi._layout(pathlib.Path(td), "ref", '#include "assembly_graph.h"\n#include "attribute.h"',
                               ["-I", test_abi.REF, "-I", os.path.join(test_abi.REF, "src")])
    with open(os.path.join(HERE, "struct_layout.txt"), "w") as f:
        f.write(ref)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "layout":
        make_struct_layout()
    elif len(sys.argv) > 1 and sys.argv[1] == "coverage":
        make_coverage_golden()
    else:
        main()
