#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/b8_bench_default.json 2> gpurun_out/b8_bench_default.err; tail -c 300 gpurun_out/b8_bench_default.err
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/b8_bench_c2.json 2> gpurun_out/b8_bench_c2.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/b8_bench_c2_ref.json 2> gpurun_out/b8_bench_c2_ref.err; tail -c 200 gpurun_out/b8_bench_c2_ref.err
timeout 900 python bench.py --workload c4 > gpurun_out/b8_bench_c4.json 2> gpurun_out/b8_bench_c4.err
timeout 600 python bench.py --workload c4 --impl reference --steps 2 --warmup 1 > gpurun_out/b8_bench_c4_ref.json 2> gpurun_out/b8_bench_c4_ref.err
timeout 1200 python bench.py --workload c5 > gpurun_out/b8_bench_c5.json 2> gpurun_out/b8_bench_c5.err
timeout 600 python bench.py --workload c5 --impl reference --steps 2 --warmup 1 > gpurun_out/b8_bench_c5_ref.json 2> gpurun_out/b8_bench_c5_ref.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b8_bench_default.json'))
print(d['value'], d['e2e']['value'], d['cpu_baseline'])
print(json.dumps(d['extra']['raster'])[:600])
print(json.dumps(d['extra']['siam_correlation'])[:900])
for w in ('c2','c4','c5'):
  print(w, json.load(open('gpurun_out/b8_bench_%s_ref.json'%w))['value'])
PY
