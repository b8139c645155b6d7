#!/bin/bash
# Round-2 battery 1 (one GPU): parity of everything new, A/B of the rasterisers,
# device-side env step, short bench.  Each test file under its own timeout.
set -x
mkdir -p gpurun_out
for f in tests/test_gpu_maxplus.py tests/test_gpu_raster.py tests/test_gpu_select.py tests/test_gpu_features.py tests/test_gpu_nets.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --maxfail=25 -p no:cacheprovider > gpurun_out/b1_$n.log 2>&1
  tail -3 gpurun_out/b1_$n.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b1_smoke.log 2>&1; tail -1 gpurun_out/b1_smoke.log
for m in 0 1; do
  SRL_RASTER_MODE=$m timeout 300 python tools/bench_raster.py 4096 10 20 >> gpurun_out/b1_raster.log 2>&1
  SRL_RASTER_MODE=$m timeout 300 python tools/bench_raster.py 4096 3 20 >> gpurun_out/b1_raster.log 2>&1
  SRL_RASTER_MODE=$m timeout 300 python tools/bench_env.py 4096 >> gpurun_out/b1_env.log 2>&1
done
cat gpurun_out/b1_raster.log gpurun_out/b1_env.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/b1_bench.json 2> gpurun_out/b1_bench.err; tail -c 600 gpurun_out/b1_bench.err
SRL_RASTER_MODE=0 python tools/bench_raster.py 4096 10 5 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -c 1 -s 4 \
  -o gpurun_out/prof_r2_raster python tools/bench_raster.py 4096 10 2 > gpurun_out/b1_ncu_raster.log 2>&1
tail -3 gpurun_out/b1_ncu_raster.log
