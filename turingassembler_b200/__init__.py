"""turingassembler_b200 — B200-native k-mer counting + de Bruijn graph build behind TuringAssembler's C entry points.

The product is ``libtagpu.so`` (hand-written sm_100a CUDA kernels + a C host layer, C ABI in ``include/tagpu.h``).
This package is only the thin ctypes mirror of that ABI used by the tests and ``bench.py``; it holds no compute
and has no CPU fallback: importing :mod:`turingassembler_b200.api` without a built ``libtagpu.so`` raises.
"""
from .api import (  # noqa: F401
    LIB_PATH,
    AsmGraph,
    Tagpu,
    TagpuError,
    build_graph_from_scratch,
    build_graph_from_scratch_without_count,
    kmc_build_kmer_database,
    load_library,
)

__all__ = [
    "LIB_PATH",
    "AsmGraph",
    "Tagpu",
    "TagpuError",
    "build_graph_from_scratch",
    "build_graph_from_scratch_without_count",
    "kmc_build_kmer_database",
    "load_library",
]
