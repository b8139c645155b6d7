#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/exp_e2e.py > gpurun_out/i7_e2e.log 2>&1; cat gpurun_out/i7_e2e.log
python tools/run_env_steps.py 16384 6 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'mask_select|pack_rewards|maxplus_stream' -s 9 -c 3 -f -o gpurun_out/prof_i7_env python tools/run_env_steps.py 16384 6 > gpurun_out/i7_ncu_env.log 2>&1; tail -2 gpurun_out/i7_ncu_env.log
python tools/run_step.py 16 8 > gpurun_out/i7_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'mask_select' -c 1 -s 20 \
  -f -o gpurun_out/prof_i7_ms_c2 python tools/run_step.py 16 8 > gpurun_out/i7_ncu_step.log 2>&1
tail -2 gpurun_out/i7_ncu_step.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/i7_u8_launches.csv python tools/exp_e2e_u8_once.py > gpurun_out/i7_u8_once.log 2>&1
