#!/bin/bash
# Round-2 closing battery (one GPU) after the warp-per-image raster / mask_select / wide-tile /
# pack_rewards changes: tests, smoke, the three bench workloads with both arms, kernel benches.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/b9_pytest.log 2>&1; tail -3 gpurun_out/b9_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b9_smoke.log 2>&1; tail -1 gpurun_out/b9_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/b9_bench_default.json 2> gpurun_out/b9_bench_default.err; tail -c 300 gpurun_out/b9_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/b9_bench_c2_ref.json 2> gpurun_out/b9_bench_c2_ref.err
timeout 900 python bench.py --workload c4 > gpurun_out/b9_bench_c4.json 2> gpurun_out/b9_bench_c4.err; tail -c 300 gpurun_out/b9_bench_c4.err
timeout 900 python bench.py --workload c4 --eager --no-cpu-baseline > gpurun_out/b9_bench_c4_eager.json 2> gpurun_out/b9_bench_c4_eager.err
timeout 600 python bench.py --workload c4 --impl reference --steps 2 --warmup 1 > gpurun_out/b9_bench_c4_ref.json 2> gpurun_out/b9_bench_c4_ref.err
timeout 1200 python bench.py --workload c5 > gpurun_out/b9_bench_c5.json 2> gpurun_out/b9_bench_c5.err; tail -c 300 gpurun_out/b9_bench_c5.err
timeout 600 python bench.py --workload c5 --impl reference --steps 2 --warmup 1 > gpurun_out/b9_bench_c5_ref.json 2> gpurun_out/b9_bench_c5_ref.err
head -c 600 gpurun_out/b9_bench_default.json; echo; head -c 400 gpurun_out/b9_bench_c4.json; echo; head -c 400 gpurun_out/b9_bench_c5.json; echo
timeout 300 python tools/bench_raster.py 4096 10 20 > gpurun_out/b9_raster.log 2>&1; cat gpurun_out/b9_raster.log
timeout 300 python tools/bench_policy_c4.py > gpurun_out/b9_policy_c4.log 2>&1; cat gpurun_out/b9_policy_c4.log
timeout 300 python tools/bench_misc.py > gpurun_out/b9_misc.log 2>&1; cat gpurun_out/b9_misc.log
