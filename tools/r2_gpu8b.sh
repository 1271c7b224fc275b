#!/bin/bash
# multi-GPU evidence after the INT8 Hessian path became the north star's default: 2-rank NCCL tests, north star at N = 8, 4, 2
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_comm.py -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu_multi.log
for N in 8 4 2; do bash tools/run_n.sh $N northstar; done
