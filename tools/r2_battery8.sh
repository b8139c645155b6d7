#!/bin/bash
# Round-2 final measurement battery, part 1 (one GPU): tests, smoke, bench lines, logs.
# (The ncu passes are tools/r2_battery8_ncu.sh: a call may bring back 64 MiB at most.)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > gpurun_out/b8_pytest.log 2>&1; tail -3 gpurun_out/b8_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b8_smoke.log 2>&1; tail -1 gpurun_out/b8_smoke.log
rm -f gpurun_out/b8_siam.log gpurun_out/b8_raster.log
SRL_SIAM_MODE=2 timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/b8_siam.log 2>&1
SRL_SIAM_MODE=2 SRL_SIAM_TC=tf32 timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/b8_siam.log 2>&1
SRL_SIAM_MODE=2 timeout 120 python tools/bench_siam.py 148 32 >> gpurun_out/b8_siam.log 2>&1
SRL_SIAM_MODE=0 timeout 120 python tools/bench_siam.py 148 16 >> gpurun_out/b8_siam.log 2>&1
grep -v oracle gpurun_out/b8_siam.log
for m in 0 1; do
  SRL_RASTER_MODE=$m timeout 300 python tools/bench_raster.py 4096 10 20 >> gpurun_out/b8_raster.log 2>&1
done
timeout 300 python tools/bench_raster.py 16384 10 10 >> gpurun_out/b8_raster.log 2>&1
timeout 300 python tools/bench_raster.py 4096 3 20 >> gpurun_out/b8_raster.log 2>&1
cat gpurun_out/b8_raster.log
timeout 300 python tools/bench_misc.py > gpurun_out/b8_misc.log 2>&1; cat gpurun_out/b8_misc.log
timeout 300 python tools/exp_e2e.py > gpurun_out/b8_e2e_chunks.log 2>&1; cat gpurun_out/b8_e2e_chunks.log
( time timeout 900 python bench.py ) > gpurun_out/b8_bench_default.json 2> gpurun_out/b8_bench_default.err; tail -c 400 gpurun_out/b8_bench_default.err
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/b8_bench_c2.json 2> gpurun_out/b8_bench_c2.err; tail -c 300 gpurun_out/b8_bench_c2.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/b8_bench_c2_ref.json 2> gpurun_out/b8_bench_c2_ref.err
timeout 900 python bench.py --workload c4 > gpurun_out/b8_bench_c4.json 2> gpurun_out/b8_bench_c4.err; tail -c 300 gpurun_out/b8_bench_c4.err
timeout 900 python bench.py --workload c4 --eager --no-cpu-baseline > gpurun_out/b8_bench_c4_eager.json 2> gpurun_out/b8_bench_c4_eager.err
timeout 600 python bench.py --workload c4 --impl reference --steps 2 --warmup 1 > gpurun_out/b8_bench_c4_ref.json 2> gpurun_out/b8_bench_c4_ref.err
timeout 1200 python bench.py --workload c5 > gpurun_out/b8_bench_c5.json 2> gpurun_out/b8_bench_c5.err; tail -c 300 gpurun_out/b8_bench_c5.err
timeout 600 python bench.py --workload c5 --impl reference --steps 2 --warmup 1 > gpurun_out/b8_bench_c5_ref.json 2> gpurun_out/b8_bench_c5_ref.err
head -c 700 gpurun_out/b8_bench_c2.json; echo; head -c 400 gpurun_out/b8_bench_c4.json; echo; head -c 400 gpurun_out/b8_bench_c5.json; echo
python tools/microbench.py 2000 0,2,7,3,6,8,14,17,18 fma > gpurun_out/b8_micro.log 2>&1; cat gpurun_out/b8_micro.log
