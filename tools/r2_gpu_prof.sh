#!/bin/sh
# developer helper (runs ON the GPU box): ncu launch list + one full capture of the two count-stage kernels (C2)
tag=${1:-r2x}
CMD="python tools/prof_run.py C2 1"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -3 gpurun_out/${tag}_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|tagpu' -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_partition|k_count_buckets|k_contract' -s 6 -c 3 -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu2.log 2>&1
tail -3 gpurun_out/${tag}_ncu2.log
ls -la gpurun_out/${tag}_prof.ncu-rep
