#!/bin/bash
# Round-2 battery 2 (one GPU): parity, rasteriser A/B after the prefetch rework, the
# three bench workloads, both arms, ncu capture of the rasteriser.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/b2_pytest.log 2>&1; tail -3 gpurun_out/b2_pytest.log
for m in 0 1; do
  SRL_RASTER_MODE=$m timeout 300 python tools/bench_raster.py 4096 10 20 >> gpurun_out/b2_raster.log 2>&1
  SRL_RASTER_MODE=$m timeout 300 python tools/bench_raster.py 4096 3 20 >> gpurun_out/b2_raster.log 2>&1
done
timeout 300 python tools/bench_env.py 4096 >> gpurun_out/b2_env.log 2>&1
timeout 300 python tools/bench_env.py 16384 >> gpurun_out/b2_env.log 2>&1
cat gpurun_out/b2_raster.log gpurun_out/b2_env.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/b2_bench_c2.json 2> gpurun_out/b2_bench_c2.err; tail -c 400 gpurun_out/b2_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b2_bench_c2_ref.json 2> gpurun_out/b2_bench_c2_ref.err
timeout 600 python bench.py --workload c4 --envs 8192 --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/b2_bench_c4_small.json 2> gpurun_out/b2_bench_c4_small.err; tail -c 600 gpurun_out/b2_bench_c4_small.err
timeout 900 python bench.py --workload c4 > gpurun_out/b2_bench_c4.json 2> gpurun_out/b2_bench_c4.err; tail -c 600 gpurun_out/b2_bench_c4.err
timeout 600 python bench.py --workload c5 --envs 592 --steps 3 --no-cpu-baseline > gpurun_out/b2_bench_c5_small.json 2> gpurun_out/b2_bench_c5_small.err; tail -c 600 gpurun_out/b2_bench_c5_small.err
timeout 1200 python bench.py --workload c5 > gpurun_out/b2_bench_c5.json 2> gpurun_out/b2_bench_c5.err; tail -c 600 gpurun_out/b2_bench_c5.err
head -c 1500 gpurun_out/b2_bench_c2.json; echo; head -c 1500 gpurun_out/b2_bench_c4.json; echo; head -c 1500 gpurun_out/b2_bench_c5.json; echo
SRL_RASTER_MODE=0 python tools/bench_raster.py 4096 10 5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -c 1 -s 4 \
  -o gpurun_out/prof_r2b_raster python tools/bench_raster.py 4096 10 2 > gpurun_out/b2_ncu_raster.log 2>&1
tail -2 gpurun_out/b2_ncu_raster.log
