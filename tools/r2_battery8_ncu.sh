#!/bin/bash
# Round-2 final battery, part 2: ncu passes.  Every pass runs after the same command has exited 0
# without ncu.  usage: r2_battery8_ncu.sh step|rest   (two calls: the reports of one call must
# stay under 64 MiB)
set -x
mkdir -p gpurun_out
if [ "$1" = step ]; then
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/b8_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra \
  > gpurun_out/b8_ncu_launches.log 2>&1
python tools/run_step.py 16 8 > gpurun_out/b8_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'maxplus_stream|mask_select' -c 2 -s 40 \
  -f -o gpurun_out/prof_r2c_step python tools/run_step.py 16 8 > gpurun_out/b8_ncu_step.log 2>&1
python tools/microbench.py 400 2,7,14,6,8 > /dev/null 2>&1 && \
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.sum,smsp__cycles_active.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,gpu__time_duration.sum \
  --clock-control none -k regex:addmax_kernel --csv --log-file gpurun_out/b8_micro_pipes.csv \
  python tools/microbench.py 400 2,7,14,6,8 > gpurun_out/b8_ncu_micro.log 2>&1
tail -2 gpurun_out/b8_ncu_micro.log
else
SRL_RASTER_MODE=0 python tools/bench_raster.py 4096 10 5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -c 1 -s 4 \
  -f -o gpurun_out/prof_r2c_raster python tools/bench_raster.py 4096 10 2 > gpurun_out/b8_ncu_raster.log 2>&1
SRL_SIAM_MODE=2 python tools/bench_siam.py 148 16 > /dev/null 2>&1 && \
SRL_SIAM_MODE=2 ncu --set full --clock-control none --import-source on -k regex:siam_tc_kernel -c 1 -s 2 \
  -f -o gpurun_out/prof_r2c_siam_tc python tools/bench_siam.py 148 16 > gpurun_out/b8_ncu_siam.log 2>&1
python tools/run_env_steps.py 16384 6 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'raster_kernel|mask_select|pack_rewards|gather_rows' -s 12 -c 4 -f -o gpurun_out/prof_r2c_env python tools/run_env_steps.py 16384 6 > gpurun_out/b8_ncu_env.log 2>&1
python tools/bench_misc.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'maxplus_u8' -c 1 -s 3 -f -o gpurun_out/prof_r2c_u8 python tools/bench_misc.py > gpurun_out/b8_ncu_u8.log 2>&1
ls -l gpurun_out
fi
