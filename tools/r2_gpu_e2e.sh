#!/bin/sh
# developer helper (ON the GPU box): files -> asm_graph_t end to end, fused vs two-pass ingest
for mode in fused twopass; do
  if [ $mode = twopass ]; then export TAGPU_INGEST_TWO_PASS=1; else unset TAGPU_INGEST_TWO_PASS; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_$mode.json 2> gpurun_out/e2e_$mode.err
  python - <<PY
import json
l = json.loads(open("gpurun_out/e2e_$mode.json").read().strip().split("\n")[-1])
print("$mode", "e2e files ms", round(l["e2e"]["ms_per_step"], 2), "k_count_buckets", round(l["kernels"]["k_count_buckets<W>"]["ms_per_launch"], 3), "step", round(l["ms_per_step"], 3))
PY
  grep "tagpu\] k=" gpurun_out/e2e_$mode.err | tail -2
done
