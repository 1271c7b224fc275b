#!/bin/bash
# one-GPU evidence after block (1,1) of config 4 moved to the INT8 tensor cores: GPU tests, config-4 bench line, ncu of one evaluation's three Hessian launches
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --config c4 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err
echo "== c4 N=1 rc=$?"; tail -c 300 gpurun_out/r2_bench_c4_n1.err; cut -c1-260 gpurun_out/r2_bench_c4_n1.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --config c4 --rows 1e7"
$B > gpurun_out/plain_c4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_hess -s 6 -c 3 -o gpurun_out/r2_fused_hess_c4 $B > gpurun_out/ncu6.log 2>&1
python tools/summarize_ncu.py gpurun_out/r2_fused_hess_c4.ncu-rep > gpurun_out/r2_fused_hess_c4_ncu.csv; rm -f gpurun_out/r2_fused_hess_c4.ncu-rep
grep -E "gpu__time_duration|dram__bytes_read.sum," gpurun_out/r2_fused_hess_c4_ncu.csv | cut -c1-160
