#!/bin/bash
# Round-2 one-GPU evidence run (inside gpurun): GPU tests, bench records of every config at N=1, the
# reference arm, compute-sanitizer logs, then the ncu captures of tools/prof_r2.sh.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for C in northstar c2 c3 c4; do
  timeout 900 python bench.py --config $C > gpurun_out/r2_bench_${C}_n1.json 2> gpurun_out/r2_bench_${C}_n1.err
  echo "== $C N=1 rc=$?"; tail -c 300 gpurun_out/r2_bench_${C}_n1.err; cut -c1-260 gpurun_out/r2_bench_${C}_n1.json
done
# config 5 does not fit one GPU with headroom (163.8 GB): the 8-GPU shard (2.5e6 rows) on one GPU
timeout 900 python bench.py --config c5 --rows 2.5e6 --steps 5 > gpurun_out/r2_bench_c5shard_n1.json 2> gpurun_out/r2_bench_c5shard_n1.err
echo "== c5 shard N=1 rc=$?"; tail -c 300 gpurun_out/r2_bench_c5shard_n1.err; cut -c1-260 gpurun_out/r2_bench_c5shard_n1.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err
echo "== reference rc=$?"; cut -c1-260 gpurun_out/r2_bench_reference_n1.json
echo "== sanitizer"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_case.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2_sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitizer_case.py > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r2_sanitizer_racecheck.txt
echo "== ncu"
bash tools/prof_r2.sh
