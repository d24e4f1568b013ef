#!/bin/sh
# developer helper (runs ON the GPU box): ncu --set full with source counters for the kernels matching a regex, one C2 build
# usage: r2_gpu_ncu.sh <tag> <kernel regex> [launch-skip] [launch-count]
tag=${1:-r2x}; re=${2:-k_contract}; skip=${3:-0}; cnt=${4:-3}
CMD="python tools/prof_run.py C2 1"
ncu --set full --clock-control none --import-source on -k regex:"$re" --launch-skip $skip --launch-count $cnt -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/${tag}_prof.ncu-rep
