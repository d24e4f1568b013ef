#!/bin/sh
# developer helper (runs ON the GPU box): for one C2 build — the ncu launch list (durations), the DRAM bytes of every launch
# (the whole-path traffic figure of bench.py's roofline), and one full capture with source counters of the three largest kernels
tag=${1:-r2x}
CMD="python tools/prof_run.py C2 1"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -3 gpurun_out/${tag}_plain.log
# prof_run does 2 warm-up builds + 1 profiled + 1 more (torch's generator kernels come first and are filtered out by name):
# the list holds all four builds, the consumers take the third
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/${tag}_dram.csv $CMD > gpurun_out/${tag}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(k_partition|k_count_buckets|k_contract)$' --launch-skip 5 --launch-count 3 -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu2.log 2>&1
tail -2 gpurun_out/${tag}_ncu2.log
ls -la gpurun_out/${tag}_prof.ncu-rep
