#!/bin/bash
# ncu launch lists (gpu__time_duration) of the default bench command and of the c4 step, final build.
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/f4_launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra \
  > gpurun_out/f4_ncu_c2.log 2>&1
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --eager > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
  --log-file gpurun_out/f4_launches_c4.csv python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --eager \
  > gpurun_out/f4_ncu_c4.log 2>&1
wc -l gpurun_out/f4_launches_c2.csv gpurun_out/f4_launches_c4.csv
