python bench.py --steps 2 --warmup 3 --rows 1e7 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --rows 1e7 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_hess -s 3 -c 1 -o gpurun_out/bench_hess_r1 python bench.py --steps 2 --warmup 3 --rows 1e7 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_eval -s 2 -c 1 -o gpurun_out/bench_grad_r1 python bench.py --steps 2 --warmup 3 --rows 1e7 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log | cut -c1-200
