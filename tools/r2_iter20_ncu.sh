#!/bin/bash
# ncu passes of the kernels this iteration changed (each after the same command exited 0 alone).
mkdir -p gpurun_out
python tools/run_env_steps.py 65536 12 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'raster_warp|mask_select' -s 18 -c 2 -f \
  -o gpurun_out/prof_r2d_env python tools/run_env_steps.py 65536 12 1 > gpurun_out/i20_ncu_env.log 2>&1
tail -2 gpurun_out/i20_ncu_env.log
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --eager > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/i20_launches_c4.csv python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --eager \
  > gpurun_out/i20_ncu_launches.log 2>&1
tail -c 300 gpurun_out/i20_ncu_launches.log
