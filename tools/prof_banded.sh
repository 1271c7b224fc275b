#!/bin/bash
# dev tool: ncu --set full of spline_banded_eval_kernel on config #3 (after the plain run exited 0)
timeout 200 python tools/time_spline.py 5e7 > gpurun_out/plain_banded.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spline_banded -s 3 -c 1 -o gpurun_out/banded_r1 python tools/time_spline.py 5e7 > gpurun_out/ncu_banded.log 2>&1
tail -2 gpurun_out/ncu_banded.log | cut -c1-200
