#!/bin/sh
# developer helper (runs ON the GPU box): ncu launch list of one C2 build + one full capture of the three largest kernels
tag=${1:-r2x}
CMD="python tools/prof_run.py C2 1"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -3 gpurun_out/${tag}_plain.log
# launch list: only this library's kernels (torch's generator kernels come first and are filtered out by name)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 120 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
# prof_run does 2 warm-up builds + 1 profiled + 1 more: skip the first build's instances, then take one of each kernel
ncu --set full --clock-control none --import-source on -k regex:'k_partition|k_count_buckets|k_contract<' --launch-skip 3 --launch-count 3 -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu2.log 2>&1
tail -2 gpurun_out/${tag}_ncu2.log
ls -la gpurun_out/${tag}_prof.ncu-rep
