#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/i2_pytest.log 2>&1; tail -5 gpurun_out/i2_pytest.log
SRL_SIAM_MODE=2 timeout 120 python tools/bench_siam.py 148 16 > gpurun_out/i2_siam.log 2>&1; grep -v oracle gpurun_out/i2_siam.log
SRL_SIAM_MODE=2 ncu --set full --clock-control none --import-source on -k regex:siam_tc_kernel -c 1 -s 2 -f -o gpurun_out/prof_i2_siam python tools/bench_siam.py 148 16 > gpurun_out/i2_ncu_siam.log 2>&1
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/i2_bench_c4.json 2> gpurun_out/i2_bench_c4.err; head -c 600 gpurun_out/i2_bench_c4.json
