#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_select.py -q -m gpu -x -p no:cacheprovider -k "host_pipeline" > gpurun_out/i11_pytest.log 2>&1; tail -3 gpurun_out/i11_pytest.log
for kt in 9 17 25; do SRL_U8_KT=$kt python tools/bench_misc.py 2>&1 | grep "maxplus_u8" | sed "s/^/KT=$kt /"; done | tee gpurun_out/i11_kt.log
timeout 300 python tools/exp_e2e.py 2>&1 | tee gpurun_out/i11_e2e.log
