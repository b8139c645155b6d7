#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/run_env_steps.py 16384 6 > gpurun_out/i5_env.log 2>&1; tail -2 gpurun_out/i5_env.log
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 8 -c 2 -f -o gpurun_out/prof_i5_env_raster python tools/run_env_steps.py 16384 6 > gpurun_out/i5_ncu.log 2>&1; tail -2 gpurun_out/i5_ncu.log
