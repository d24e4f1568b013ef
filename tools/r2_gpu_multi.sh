#!/bin/sh
# developer helper (ON a multi-GPU box): parity of the sharded path against the oracle, then strong-scaling bench lines,
# with the NVLink data counters read before and after the bench (nvidia-smi nvlink -gt d: KiB per link since reset)
N=${1:-2}; tag=${2:-r2x}; wl=${3:-C2}; steps=${4:-10}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@"; }
run tools/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.log 2>&1
echo "dist_check rc=$?"; grep -c PARITY gpurun_out/${tag}_dist_check_${N}gpu.log; grep -c MISMATCH gpurun_out/${tag}_dist_check_${N}gpu.log
nvidia-smi nvlink -gt d > gpurun_out/${tag}_nvlink_before_${N}gpu.txt 2>&1
run bench.py --gpus $N --steps $steps --warmup 3 --workload $wl > gpurun_out/${tag}_bench_${wl}_${N}gpu.json 2> gpurun_out/${tag}_bench_${wl}_${N}gpu.err
echo "bench rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/${tag}_nvlink_after_${N}gpu.txt 2>&1
python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/${tag}_bench_${wl}_${N}gpu.json").read().strip().split("\n") if x.startswith("{")][-1])
    print("N", l["n_gpus"], "ms_per_step", round(l["ms_per_step"], 3), "stage_ms", l["stage_ms"], "e2e ms", round(l["e2e"]["ms_per_step"], 3), "digest ok", l["result"]["digest"]["matches_reference_golden"])
    for k, v in list(l["kernels"].items())[:12]:
        print(" ", k, round(v["ms_per_launch"], 4), v["launches_per_step"])
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -3 gpurun_out/${tag}_bench_${wl}_${N}gpu.err
