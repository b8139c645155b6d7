#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/i3_pytest.log 2>&1; tail -5 gpurun_out/i3_pytest.log
SRL_SIAM_MODE=2 timeout 120 python tools/bench_siam.py 148 16 > gpurun_out/i3_siam.log 2>&1
SRL_SIAM_MODE=2 SRL_SIAM_TC=tf32 timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/i3_siam.log 2>&1
SRL_SIAM_MODE=2 timeout 120 python tools/bench_siam.py 148 32 >> gpurun_out/i3_siam.log 2>&1
grep -v oracle gpurun_out/i3_siam.log
SRL_SIAM_MODE=2 ncu --set full --clock-control none --import-source on -k regex:siam_tc_kernel -c 1 -s 2 -f -o gpurun_out/prof_i3_siam python tools/bench_siam.py 148 16 > gpurun_out/i3_ncu_siam.log 2>&1
