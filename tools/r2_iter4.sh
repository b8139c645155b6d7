#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/i4_pytest.log 2>&1; tail -5 gpurun_out/i4_pytest.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/i4_bench_c4.json 2> gpurun_out/i4_bench_c4.err; tail -c 300 gpurun_out/i4_bench_c4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/i4_bench_c4.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], json.dumps(d['roofline']['breakdown_ms']))
PY
