#!/bin/bash
# A/B of the warp-per-image raster kernel at config 4 (tests first).
mkdir -p gpurun_out
timeout 250 python -m pytest tests/test_gpu_raster.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for c in 7 8; do
  SRL_RASTER_WARP_CTAS=$c timeout 280 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/wc${c}_c4.json 2> gpurun_out/wc${c}_c4.err
  python -c "
import json
d=json.load(open('gpurun_out/wc${c}_c4.json'))
print($c, d['ms_per_step'], d['value'], json.dumps(d['roofline']['breakdown_ms']))
"
done
