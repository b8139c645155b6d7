#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_select.py tests/test_gpu_raster.py -q -m gpu -x -p no:cacheprovider > gpurun_out/i9_pytest.log 2>&1; tail -5 gpurun_out/i9_pytest.log
for m in 0 1 2; do SRL_MS_COLUMNS=$m python tools/bench_mask_select_c4.py 65536; done 2>&1 | tee gpurun_out/i9_ms.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/i9_bench_c4.json 2> gpurun_out/i9_bench_c4.err; tail -c 400 gpurun_out/i9_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/i9_bench_c4.json')); print(d['value'], d['ms_per_step'], d['roofline']['breakdown_ms'])"
SRL_MS_COLUMNS=1 ncu --set full --clock-control none --import-source on -k regex:'mask_select' -c 1 -s 4 -f -o gpurun_out/prof_i9_ms python tools/bench_mask_select_c4.py 16384 > gpurun_out/i9_ncu.log 2>&1; tail -2 gpurun_out/i9_ncu.log
