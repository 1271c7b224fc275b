#!/bin/bash
# refresh of the one-GPU evidence after a change to the north-star kernel: GPU tests, north star + config 2 bench lines, ncu of the kernel
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for C in northstar c2; do
  timeout 900 python bench.py --config $C > gpurun_out/r2_bench_${C}_n1.json 2> gpurun_out/r2_bench_${C}_n1.err
  echo "== $C N=1 rc=$?"; tail -c 300 gpurun_out/r2_bench_${C}_n1.err; cut -c1-260 gpurun_out/r2_bench_${C}_n1.json
done
bash tools/prof_i8.sh
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --rows 1e7"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fused_|finish_|colmax|comm_" -c 400 --csv --log-file gpurun_out/r2_launches_bench_rows1e7.csv $B > gpurun_out/ncu1.log 2>&1
