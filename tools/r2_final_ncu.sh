#!/bin/bash
# ncu pass of the environment-step kernels of the final build (after the same command exited 0).
mkdir -p gpurun_out
python tools/run_env_steps.py 65536 12 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'raster_warp|mask_select|pack_rewards|place_poses|maxplus_stream' -s 55 -c 5 -f \
  -o gpurun_out/prof_r2e_env python tools/run_env_steps.py 65536 12 1 > gpurun_out/f3_ncu_env.log 2>&1
tail -2 gpurun_out/f3_ncu_env.log; ls -la gpurun_out/prof_r2e_env.ncu-rep
