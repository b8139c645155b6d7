#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_maxplus.py tests/test_gpu_select.py -q -m gpu -x -p no:cacheprovider > gpurun_out/i10_pytest.log 2>&1; tail -3 gpurun_out/i10_pytest.log
python tools/bench_misc.py 2>&1 | grep maxplus | tee gpurun_out/i10_misc.log
timeout 300 python tools/exp_e2e.py 2>&1 | grep uint8 | tee gpurun_out/i10_e2e.log
ncu --set full --clock-control none --import-source on -k regex:'maxplus_u8' -c 1 -s 3 -f -o gpurun_out/prof_i10_u8 python tools/bench_misc.py > gpurun_out/i10_ncu_u8.log 2>&1; tail -1 gpurun_out/i10_ncu_u8.log
