python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_default.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_default.log 2>&1
tail -1 gpurun_out/ncu_launch_default.log | cut -c1-200
