#!/bin/sh
# developer helper (ON the GPU box): default vs alternative k_count_buckets geometry, C2 and C1
for lib in libtagpu.so libtagpu_alt.so; do
  for wl in C2 C1; do
    echo "== $lib $wl"
    TAGPU_LIB=$PWD/turingassembler_b200/$lib python tools/prof_run.py $wl 5 2>&1 | grep -E "count |k_count_buckets|k_partition|sum of kernels" | head -5
  done
done
