#!/bin/bash
# Last pass on the final build: GPU tests, smoke, the default bench line and the config-4 lines.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/b8_pytest.log 2>&1; tail -3 gpurun_out/b8_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b8_smoke.log 2>&1; tail -1 gpurun_out/b8_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/b8_bench_default.json 2> gpurun_out/b8_bench_default.err; tail -c 200 gpurun_out/b8_bench_default.err
timeout 900 python bench.py --workload c4 > gpurun_out/b8_bench_c4.json 2> gpurun_out/b8_bench_c4.err; tail -c 300 gpurun_out/b8_bench_c4.err
timeout 900 python bench.py --workload c4 --eager --no-cpu-baseline > gpurun_out/b8_bench_c4_eager.json 2> gpurun_out/b8_bench_c4_eager.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b8_bench_default.json')); print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_uint8']['value'], d['roofline']['frac'])
for n in ('c4', 'c4_eager'):
  d=json.load(open('gpurun_out/b8_bench_%s.json' % n)); print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['breakdown_ms'])
PY
python tools/run_env_steps.py 16384 6 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'raster_kernel|mask_select|pack_rewards|gather_rows' -s 12 -c 4 -f -o gpurun_out/prof_r2c_env python tools/run_env_steps.py 16384 6 1 > gpurun_out/b8_ncu_env.log 2>&1; tail -1 gpurun_out/b8_ncu_env.log
