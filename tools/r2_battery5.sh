#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nets.py -q -m gpu --maxfail=30 -p no:cacheprovider -x > gpurun_out/b5_nets.log 2>&1; echo "nets rc=$?"; tail -5 gpurun_out/b5_nets.log
rm -f gpurun_out/b5_siam.log
for m in 1 0; do
  SRL_SIAM_MODE=$m timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/b5_siam.log 2>&1
  SRL_SIAM_MODE=$m timeout 120 python tools/bench_siam.py 296 16 >> gpurun_out/b5_siam.log 2>&1
  SRL_SIAM_MODE=$m timeout 120 python tools/bench_siam.py 32 16 >> gpurun_out/b5_siam.log 2>&1
done
grep -v oracle gpurun_out/b5_siam.log
SRL_SIAM_MODE=1 python tools/bench_siam.py 148 16 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:siam_tc_kernel -c 1 -s 2 \
  -o gpurun_out/prof_r2_siam_tc python tools/bench_siam.py 148 16 > gpurun_out/b5_ncu_siam.log 2>&1
tail -2 gpurun_out/b5_ncu_siam.log
