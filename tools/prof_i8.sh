#!/bin/bash
# dev tool: ncu --set full of fused_hess_i8_kernel on the north-star shape at 1e7 rows (after the plain run exited 0)
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --rows 1e7"
$B > gpurun_out/plain_i8.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_i8.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_hess_i8 -s 4 -c 1 -o gpurun_out/r2_fused_hess_i8 $B > gpurun_out/ncu_i8.log 2>&1
tail -2 gpurun_out/ncu_i8.log | cut -c1-200
python tools/summarize_ncu.py gpurun_out/r2_fused_hess_i8.ncu-rep > gpurun_out/r2_fused_hess_i8_kernel_ncu.csv
ncu -i gpurun_out/r2_fused_hess_i8.ncu-rep --page source --csv > gpurun_out/r2_fused_hess_i8_source.csv 2>/dev/null; gzip -f gpurun_out/r2_fused_hess_i8_source.csv
rm -f gpurun_out/r2_fused_hess_i8.ncu-rep
cut -c1-220 gpurun_out/plain_i8.log | tail -1
