#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 16 --csv --log-file gpurun_out/i6_launches.csv python tools/run_env_steps.py 65536 8 > gpurun_out/i6.log 2>&1
tail -2 gpurun_out/i6.log
