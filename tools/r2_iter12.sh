#!/bin/bash
set -x
mkdir -p gpurun_out
for e in 8192 65536; do
  timeout 600 python bench.py --workload c4 --envs $e --no-cpu-baseline > gpurun_out/i12_c4_graph_$e.json 2> gpurun_out/i12.err; tail -c 300 gpurun_out/i12.err
  timeout 600 python bench.py --workload c4 --envs $e --no-cpu-baseline --eager > gpurun_out/i12_c4_eager_$e.json 2> gpurun_out/i12.err; tail -c 300 gpurun_out/i12.err
done
python - <<'PY'
import json
for e in (8192, 65536):
  for m in ('graph', 'eager'):
    d = json.load(open('gpurun_out/i12_c4_%s_%d.json' % (m, e)))
    print(e, m, '%.3e env obs/s  %.3f ms  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline']['breakdown_ms'])
PY
