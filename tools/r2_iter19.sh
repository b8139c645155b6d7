#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/run_env_steps.py 65536 6 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'pack_rewards' -s 4 -c 1 -f -o gpurun_out/prof_i19_pack python tools/run_env_steps.py 65536 6 1 > gpurun_out/i19_ncu.log 2>&1; tail -2 gpurun_out/i19_ncu.log
