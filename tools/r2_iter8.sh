#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_select.py tests/test_gpu_raster.py -q -m gpu -x -p no:cacheprovider > gpurun_out/i8_pytest.log 2>&1; tail -5 gpurun_out/i8_pytest.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/i8_bench_c4.json 2> gpurun_out/i8_bench_c4.err; tail -c 400 gpurun_out/i8_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/i8_bench_c4.json')); print(d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'])"
python tools/bench_misc.py > gpurun_out/i8_misc.log 2>&1; head -20 gpurun_out/i8_misc.log
ncu --set full --clock-control none --import-source on -k regex:'maxplus_u8' -c 1 -s 3 -f -o gpurun_out/prof_i8_u8 python tools/bench_misc.py > gpurun_out/i8_ncu_u8.log 2>&1; tail -2 gpurun_out/i8_ncu_u8.log
