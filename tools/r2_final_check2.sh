#!/bin/bash
# Last pass on the final build: tests, smoke, the default bench line and the c4 line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/f2_pytest.log 2>&1; tail -3 gpurun_out/f2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2_smoke.log 2>&1; tail -1 gpurun_out/f2_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/f2_bench_default.json 2> gpurun_out/f2_bench_default.err; tail -4 gpurun_out/f2_bench_default.err
timeout 900 python bench.py --workload c4 > gpurun_out/f2_bench_c4.json 2> gpurun_out/f2_bench_c4.err; tail -c 300 gpurun_out/f2_bench_c4.err
timeout 900 python bench.py --workload c4 --eager --no-cpu-baseline > gpurun_out/f2_bench_c4_eager.json 2> gpurun_out/f2_bench_c4_eager.err
head -c 300 gpurun_out/f2_bench_default.json; echo; head -c 300 gpurun_out/f2_bench_c4.json; echo
