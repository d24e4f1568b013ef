#!/bin/sh
# developer helper (runs ON the GPU box): parity tests, then a short bench, then the phase breakdown of k_count_buckets
tag=${1:-r2x}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
tail -12 gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench1.json 2> gpurun_out/${tag}_bench1.err
python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/${tag}_bench1.json").read().strip().split("\n")[-1])
    print("ms_per_step", l["ms_per_step"], "stage_ms", l["stage_ms"], "frac", l["roofline"]["frac"], "digest ok", l["result"]["digest"]["matches_reference_golden"])
    for k, v in list(l["kernels"].items())[:8]:
        print(" ", k, round(v["ms_per_launch"], 4), v["launches_per_step"])
    for k in ("e2e", "e2e_host_stream", "e2e_host_stream_packed"):
        print(" ", k, round(l[k]["ms_per_step"], 2), "ms")
except Exception as e:
    print("bench failed:", e)
PY
tail -4 gpurun_out/${tag}_bench1.err
TAGPU_LIB=$PWD/turingassembler_b200/libtagpu_timing.so python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-files > /dev/null 2> gpurun_out/${tag}_timing.err
grep "tagpu timing" gpurun_out/${tag}_timing.err | tail -2
