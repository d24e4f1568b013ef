#!/bin/sh
# developer helper (ON a multi-GPU box): strong-scaling bench lines of several workloads at N GPUs (+ parity check at 8)
N=${1:-2}; tag=${2:-r2x}; shift 2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
if [ "$N" = 8 ]; then
	run tools/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.log 2>&1
	echo "dist_check rc=$? PARITY $(grep -c PARITY gpurun_out/${tag}_dist_check_${N}gpu.log) MISMATCH $(grep -c MISMATCH gpurun_out/${tag}_dist_check_${N}gpu.log)"
fi
for wl in "$@"; do
	steps=20; [ "$wl" = C2 ] || steps=5
	timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps $steps --warmup 3 --workload $wl > gpurun_out/${tag}_bench_${wl}_${N}gpu.json 2> gpurun_out/${tag}_bench_${wl}_${N}gpu.err
	python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/${tag}_bench_${wl}_${N}gpu.json").read().strip().split("\n") if x.startswith("{")][-1])
    print("$wl N", l["n_gpus"], "ms_per_step", round(l["ms_per_step"], 3), "stage_ms", {k: round(v, 2) for k, v in l["stage_ms"].items()}, "e2e ms", round(l["e2e"]["ms_per_step"], 2), "digest", l["result"]["digest"]["matches_reference_golden"])
except Exception as e:
    print("$wl: bench line unreadable:", e)
PY
done
