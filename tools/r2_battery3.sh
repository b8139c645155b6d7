#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/b3_pytest.log 2>&1; tail -3 gpurun_out/b3_pytest.log
timeout 300 python tools/bench_env.py 4096 > gpurun_out/b3_env.log 2>&1
timeout 300 python tools/bench_env.py 16384 >> gpurun_out/b3_env.log 2>&1
cat gpurun_out/b3_env.log
timeout 900 python bench.py --workload c4 > gpurun_out/b3_bench_c4.json 2> gpurun_out/b3_bench_c4.err; tail -c 600 gpurun_out/b3_bench_c4.err
head -c 600 gpurun_out/b3_bench_c4.json
