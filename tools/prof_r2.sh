#!/bin/bash
# Round-2 ncu captures (run under gpurun, ONE GPU).  Every profiled command first runs plain and must exit 0.
set -u
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
run_plain() { "$@" > gpurun_out/plain.log 2>&1 || { echo "plain run failed: $*"; tail -5 gpurun_out/plain.log; return 1; }; }

# 1. north star shape at 1e7 rows: launch list + full captures of the INT8-tensor-core Hessian kernel, of the DMMA kernel (the
#    bench's hess_i8 = 0 leg) and of the epilogue
run_plain $B --rows 1e7 && {
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fused_|finish_|banded_|wide_|comm_|linpred|rowterms|xtr_" -c 400 --csv --log-file gpurun_out/r2_launches_bench_rows1e7.csv $B --rows 1e7 > gpurun_out/ncu1.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_hess_i8 -s 4 -c 1 -o gpurun_out/r2_fused_hess_i8_kernel $B --rows 1e7 > gpurun_out/ncu2a.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_hess_kernel -s 1 -c 1 -o gpurun_out/r2_fused_hess $B --rows 1e7 > gpurun_out/ncu2.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:finish_fused -s 4 -c 1 -o gpurun_out/r2_finish $B --rows 1e7 > gpurun_out/ncu2b.log 2>&1
}
# 2. config 3 (banded spline evaluator), full size
run_plain $B --config c3 && {
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fused_|finish_|banded_|wide_|comm_|linpred|rowterms|xtr_" -c 400 --csv --log-file gpurun_out/r2_launches_bench_c3.csv $B --config c3 > gpurun_out/ncu3.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spline_banded -s 3 -c 1 -o gpurun_out/r2_spline_banded_bulk $B --config c3 > gpurun_out/ncu4.log 2>&1
}
# 3. config 4 at 1e7 rows: the three Hessian launches of one evaluation (MODE 3, MODE 1, MODE 1)
run_plain $B --config c4 --rows 1e7 && {
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fused_|finish_|banded_|wide_|comm_|linpred|rowterms|xtr_" -c 400 --csv --log-file gpurun_out/r2_launches_bench_c4_rows1e7.csv $B --config c4 --rows 1e7 > gpurun_out/ncu5.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_hess -s 6 -c 3 -o gpurun_out/r2_fused_hess_c4 $B --config c4 --rows 1e7 > gpurun_out/ncu6.log 2>&1
}
# 4. config 5 shard (2.5e6 x 1024): the wide Hessian kernel
run_plain $B --config c5 --rows 2.5e6 && {
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_hessian -s 2 -c 1 -o gpurun_out/r2_wide_hessian_kernel $B --config c5 --rows 2.5e6 > gpurun_out/ncu7.log 2>&1
}
# gpurun brings back at most 64 MiB: keep the metric extracts (and the raw pages), drop the reports
for r in gpurun_out/r2_*.ncu-rep; do
  python tools/summarize_ncu.py $r > ${r%.ncu-rep}_ncu.csv
  ncu -i $r --page source --csv > ${r%.ncu-rep}_source.csv 2>/dev/null
  gzip -f ${r%.ncu-rep}_source.csv
  rm -f $r
done
ls -la gpurun_out/ | tail -20
