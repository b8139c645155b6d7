#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_capi_symbols.py -q -m gpu -x -p no:cacheprovider > gpurun_out/i14_pytest.log 2>&1; tail -5 gpurun_out/i14_pytest.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/i14_bench_c4.json 2> gpurun_out/i14.err; tail -c 300 gpurun_out/i14.err
python -c "
import json; d=json.load(open('gpurun_out/i14_bench_c4.json')); print(d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d['e2e']['value'])"
python tools/run_env_steps.py 16384 6 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'raster_kernel' -s 3 -c 1 -f -o gpurun_out/prof_i14_wall python tools/run_env_steps.py 16384 6 > gpurun_out/i14_ncu.log 2>&1; tail -2 gpurun_out/i14_ncu.log
