#!/bin/bash
# Round-end measurement battery (run under gpurun, one GPU).  Every ncu pass runs
# after the same command has exited 0 without ncu.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py --steps 200 --warmup 10 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
  > gpurun_out/final_ncu_launches.log 2>&1
python tools/run_step.py 5 > gpurun_out/final_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'maxplus_stream|mask_select' -c 2 -s 6 \
  -o gpurun_out/prof_final_step python tools/run_step.py 2 > gpurun_out/final_ncu_step.log 2>&1
python tools/bench_raster.py 4096 3 5 > gpurun_out/final_raster.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:raster -c 1 -s 4 \
  -o gpurun_out/prof_final_raster python tools/bench_raster.py 4096 3 2 > gpurun_out/final_ncu_raster.log 2>&1
python tools/bench_siam.py 148 16 > gpurun_out/final_siam.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:siam -c 1 -s 3 \
  -o gpurun_out/prof_final_siam python tools/bench_siam.py 148 16 > gpurun_out/final_ncu_siam.log 2>&1
python tools/bench_siam.py 32 16 >> gpurun_out/final_siam.log 2>&1
python tools/bench_siam.py 148 64 64 16 >> gpurun_out/final_siam.log 2>&1
python tools/microbench.py 2000 0,2,7,3,6,8,14 fma > gpurun_out/final_micro.log 2>&1
python tools/bench_configs.py > gpurun_out/final_configs.log 2>&1
python tools/bench_misc.py > gpurun_out/final_misc.log 2>&1
cat gpurun_out/final_configs.log gpurun_out/final_micro.log
