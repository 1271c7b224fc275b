#!/bin/bash
# usage: run_n.sh N config extra-args...   (inside gpurun)
N=$1; C=$2; shift 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N --config $C "$@" > gpurun_out/r2_bench_${C}_n${N}.json 2> gpurun_out/r2_bench_${C}_n${N}.err
echo "== $C N=$N rc=$?"; tail -c 300 gpurun_out/r2_bench_${C}_n${N}.err; cut -c1-200 gpurun_out/r2_bench_${C}_n${N}.json
