#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/i17.json 2> gpurun_out/i17.err
  python -c "
import json; d=json.load(open('gpurun_out/i17.json')); b=d['roofline']['breakdown_ms']; print('$*', '%.3f ms' % d['ms_per_step'], ' '.join('%.3f' % v for v in b.values()))"
}
run A=1
run SRL_RASTER_CTAS=8
run SRL_MP_NSLOT=4
run SRL_MP_NSLOT=6
