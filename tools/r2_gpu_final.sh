#!/bin/sh
# developer helper (runs ON the GPU box): the round's final 1-GPU evidence — parity suite, the bench line as the driver runs it,
# the C1 and C3 lines, the k = 21/31/45/63 sweep
tag=${1:-r2fin}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_C2_1gpu.json 2> gpurun_out/${tag}_bench_C2_1gpu.err; echo "C2 rc=$?"
python bench.py --steps 20 --warmup 5 --workload C1 --no-cpu-baseline > gpurun_out/${tag}_bench_C1_1gpu.json 2> gpurun_out/${tag}_bench_C1_1gpu.err; echo "C1 rc=$?"
python bench.py --steps 5 --warmup 3 --workload C3 --no-cpu-baseline --no-files > gpurun_out/${tag}_bench_C3_1gpu.json 2> gpurun_out/${tag}_bench_C3_1gpu.err; echo "C3 rc=$?"
python - <<PY
import json
for wl in ("C2", "C1", "C3"):
    try:
        l = json.loads([x for x in open("gpurun_out/${tag}_bench_%s_1gpu.json" % wl).read().strip().split("\n") if x.startswith("{")][-1])
        print(wl, "ms_per_step", round(l["ms_per_step"], 3), "stage_ms", l["stage_ms"], "frac", round(l["roofline"]["frac"], 4), "digest ok", l["result"]["digest"]["matches_reference_golden"],
              "e2e ms", round(l["e2e"].get("ms_per_step", 0), 2) if "e2e" in l else None)
    except Exception as e:
        print(wl, "unreadable:", e)
PY
