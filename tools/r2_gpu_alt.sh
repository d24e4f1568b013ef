#!/bin/sh
# developer helper (ON the GPU box): compare builds of the library on C2 (per-kernel CUDA-event times)
for lib in "$@"; do
    echo "== $lib"
    TAGPU_LIB=$PWD/turingassembler_b200/$lib python tools/prof_run.py C2 5 2>&1 | grep -E "count |k_count_buckets|k_partition|k_contract<|sum of kernels" | head -6
done
