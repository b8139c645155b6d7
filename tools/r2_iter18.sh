#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_capi_symbols.py -q -m gpu -x -p no:cacheprovider > gpurun_out/i18_pytest.log 2>&1; tail -8 gpurun_out/i18_pytest.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/i18_bench_c4.json 2> gpurun_out/i18.err; tail -c 300 gpurun_out/i18.err
python -c "
import json; d=json.load(open('gpurun_out/i18_bench_c4.json')); print(d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['e2e']['value'])"
