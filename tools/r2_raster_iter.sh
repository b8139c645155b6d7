#!/bin/bash
# Raster iteration: bitwise parity tests, then throughput at config 3, then (optionally) ncu.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py -q -m gpu -x -p no:cacheprovider > gpurun_out/ri_pytest.log 2>&1; tail -5 gpurun_out/ri_pytest.log
rm -f gpurun_out/ri_raster.log
for c in 8 7; do
SRL_RASTER_CTAS=$c timeout 300 python tools/bench_raster.py 4096 10 20 >> gpurun_out/ri_raster.log 2>&1
SRL_RASTER_CTAS=$c timeout 300 python tools/bench_raster.py 4096 3 20 >> gpurun_out/ri_raster.log 2>&1
SRL_RASTER_CTAS=$c timeout 300 python tools/bench_raster.py 16384 10 10 >> gpurun_out/ri_raster.log 2>&1
done
cat gpurun_out/ri_raster.log
if [ "$1" = "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -c 1 -s 4 \
  -f -o gpurun_out/prof_ri_raster python tools/bench_raster.py 4096 10 2 > gpurun_out/ri_ncu.log 2>&1
tail -2 gpurun_out/ri_ncu.log
fi
