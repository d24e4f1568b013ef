#!/bin/sh
# developer helper (ON the GPU box): single-context and batched local builds, work-sized grids vs full-device grids, interleaved
# on the same box so that box-to-box differences cancel (3 rounds)
for round in 1 2 3; do
  for mode in sized full; do
    if [ $mode = full ]; then export TAGPU_FULL_GRIDS=1; else unset TAGPU_FULL_GRIDS; fi
    for c in L1 L2; do
      echo "== round $round grids=$mode $c"
      LOCAL_BENCH_JOBS=512 python tools/local_bench.py $c 50 2>&1 | grep -E "tagpu_build_local_host|contexts" | grep -E "local_host| 1 contexts| 8 contexts|16 contexts"
    done
  done
done
