#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nets.py -q -m gpu --maxfail=30 -p no:cacheprovider -x > gpurun_out/b4_nets.log 2>&1; echo "nets rc=$?"; tail -25 gpurun_out/b4_nets.log
for m in 1 0; do
  SRL_SIAM_MODE=$m timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/b4_siam.log 2>&1
  SRL_SIAM_MODE=$m timeout 120 python tools/bench_siam.py 32 16 >> gpurun_out/b4_siam.log 2>&1
done
cat gpurun_out/b4_siam.log
timeout 900 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider --deselect tests/test_gpu_nets.py > gpurun_out/b4_pytest.log 2>&1; tail -3 gpurun_out/b4_pytest.log
timeout 300 python tools/bench_env.py 16384 > gpurun_out/b4_env.log 2>&1; cat gpurun_out/b4_env.log
