#!/bin/bash
# Round-2 final measurement battery (one GPU).  Every ncu pass runs after the same command has
# exited 0 without ncu.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/b8_pytest.log 2>&1; tail -3 gpurun_out/b8_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b8_smoke.log 2>&1; tail -1 gpurun_out/b8_smoke.log
rm -f gpurun_out/b8_siam.log gpurun_out/b8_raster.log
SRL_SIAM_MODE=2 timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/b8_siam.log 2>&1
SRL_SIAM_MODE=2 SRL_SIAM_TC=tf32 timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/b8_siam.log 2>&1
SRL_SIAM_MODE=2 timeout 120 python tools/bench_siam.py 148 32 >> gpurun_out/b8_siam.log 2>&1
SRL_SIAM_MODE=0 timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/b8_siam.log 2>&1
grep -v oracle gpurun_out/b8_siam.log
for m in 0 1; do
  SRL_RASTER_MODE=$m timeout 300 python tools/bench_raster.py 4096 10 20 >> gpurun_out/b8_raster.log 2>&1
done
timeout 300 python tools/bench_raster.py 16384 10 10 >> gpurun_out/b8_raster.log 2>&1
timeout 300 python tools/bench_raster.py 4096 3 20 >> gpurun_out/b8_raster.log 2>&1
cat gpurun_out/b8_raster.log
timeout 300 python tools/bench_misc.py > gpurun_out/b8_misc.log 2>&1; cat gpurun_out/b8_misc.log
timeout 300 python tools/exp_e2e.py > gpurun_out/b8_e2e_chunks.log 2>&1; cat gpurun_out/b8_e2e_chunks.log
( time timeout 900 python bench.py ) > gpurun_out/b8_bench_default.json 2> gpurun_out/b8_bench_default.err; tail -c 400 gpurun_out/b8_bench_default.err
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/b8_bench_c2.json 2> gpurun_out/b8_bench_c2.err; tail -c 300 gpurun_out/b8_bench_c2.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/b8_bench_c2_ref.json 2> gpurun_out/b8_bench_c2_ref.err
timeout 900 python bench.py --workload c4 > gpurun_out/b8_bench_c4.json 2> gpurun_out/b8_bench_c4.err; tail -c 300 gpurun_out/b8_bench_c4.err
timeout 900 python bench.py --workload c4 --eager --no-cpu-baseline > gpurun_out/b8_bench_c4_eager.json 2> gpurun_out/b8_bench_c4_eager.err
timeout 600 python bench.py --workload c4 --impl reference --steps 2 --warmup 1 > gpurun_out/b8_bench_c4_ref.json 2> gpurun_out/b8_bench_c4_ref.err
timeout 1200 python bench.py --workload c5 > gpurun_out/b8_bench_c5.json 2> gpurun_out/b8_bench_c5.err; tail -c 300 gpurun_out/b8_bench_c5.err
timeout 600 python bench.py --workload c5 --impl reference --steps 2 --warmup 1 > gpurun_out/b8_bench_c5_ref.json 2> gpurun_out/b8_bench_c5_ref.err
head -c 700 gpurun_out/b8_bench_c2.json; echo; head -c 400 gpurun_out/b8_bench_c4.json; echo; head -c 400 gpurun_out/b8_bench_c5.json; echo
python tools/microbench.py 2000 0,2,7,3,6,8,14,17,18 fma > gpurun_out/b8_micro.log 2>&1; cat gpurun_out/b8_micro.log
# ---- ncu: launch list of the benched command, then full captures of the top kernels ---- #
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/b8_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra \
  > gpurun_out/b8_ncu_launches.log 2>&1
python tools/run_step.py 16 8 > gpurun_out/b8_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'maxplus_stream|mask_select' -c 2 -s 40 \
  -f -o gpurun_out/prof_r2c_step python tools/run_step.py 16 8 > gpurun_out/b8_ncu_step.log 2>&1
SRL_RASTER_MODE=0 python tools/bench_raster.py 4096 10 5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -c 1 -s 4 \
  -f -o gpurun_out/prof_r2c_raster python tools/bench_raster.py 4096 10 2 > gpurun_out/b8_ncu_raster.log 2>&1
SRL_SIAM_MODE=2 python tools/bench_siam.py 148 16 > /dev/null 2>&1 && \
SRL_SIAM_MODE=2 ncu --set full --clock-control none --import-source on -k regex:siam_tc_kernel -c 1 -s 2 \
  -f -o gpurun_out/prof_r2c_siam_tc python tools/bench_siam.py 148 16 > gpurun_out/b8_ncu_siam.log 2>&1
python tools/run_env_steps.py 16384 6 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'raster_kernel|mask_select|pack_rewards|maxplus_stream|gather_rows|place_poses' -s 18 -c 6 -f -o gpurun_out/prof_r2c_env python tools/run_env_steps.py 16384 6 > gpurun_out/b8_ncu_env.log 2>&1
python tools/bench_misc.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'maxplus_u8' -c 1 -s 3 -f -o gpurun_out/prof_r2c_u8 python tools/bench_misc.py > gpurun_out/b8_ncu_u8.log 2>&1
python tools/microbench.py 400 2,7,14,6,8 > /dev/null 2>&1 && \
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.sum,smsp__cycles_active.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,gpu__time_duration.sum \
  --clock-control none -k regex:addmax_kernel --csv --log-file gpurun_out/b8_micro_pipes.csv \
  python tools/microbench.py 400 2,7,14,6,8 > gpurun_out/b8_ncu_micro.log 2>&1
tail -2 gpurun_out/b8_ncu_micro.log
