#!/bin/bash
# Round-2 multi-GPU evidence run (inside `gpurun --gpus 8`): the 2-rank NCCL tests, then bench records of the
# sharded configs at N = 8, 4, 2 (tools/run_n.sh writes gpurun_out/r2_bench_<config>_n<N>.json).
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== pytest multi + repeat"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_repeat.py tests/test_gpu_comm.py -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu_multi.log
for C in northstar c4 c5 c3; do bash tools/run_n.sh 8 $C; done
for N in 4 2; do for C in northstar c4; do bash tools/run_n.sh $N $C; done; done
