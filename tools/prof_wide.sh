#!/bin/bash
# dev tool: ncu --set full of wide_hessian_kernel on a config-#5-like shape (after the plain run exited 0)
timeout 200 python tools/time_config.py c5 1e6 1024 > gpurun_out/plain_wide.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_hessian -s 2 -c 1 -o gpurun_out/wide_r1 python tools/time_config.py c5 1e6 1024 > gpurun_out/ncu_wide.log 2>&1
tail -2 gpurun_out/ncu_wide.log | cut -c1-200
cat gpurun_out/plain_wide.log | cut -c1-300
