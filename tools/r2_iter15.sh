#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_select.py tests/test_gpu_features.py -q -m gpu -x -p no:cacheprovider > gpurun_out/i15_pytest.log 2>&1; tail -3 gpurun_out/i15_pytest.log
python tools/bench_mask_select.py 2>&1 | tee gpurun_out/i15_ms_c2.log
python tools/run_step.py 40 8 2>&1 | tee -a gpurun_out/i15_ms_c2.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/i15_bench_c4.json 2> gpurun_out/i15.err; tail -c 300 gpurun_out/i15.err
python -c "
import json; d=json.load(open('gpurun_out/i15_bench_c4.json')); print(d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'], d['roofline']['frac'], d['e2e']['value'])"
