#!/bin/bash
# Copy the final battery's outputs (gpurun_out/b8_*, prof_r2c_*) into profiles/ (tracked).
set -x
G=gpurun_out; P=profiles
cp $G/b8_bench_c2.json $P/r2_bench_c2_n1.json
cp $G/b8_bench_default.json $P/r2_bench_default_n1.json
cp $G/b8_bench_c2_ref.json $P/r2_bench_c2_reference_arm.json
cp $G/b8_bench_c4.json $P/r2_bench_c4_n1.json
cp $G/b8_bench_c4_eager.json $P/r2_bench_c4_eager_n1.json
cp $G/b8_bench_c4_ref.json $P/r2_bench_c4_reference_arm.json
cp $G/b8_bench_c5.json $P/r2_bench_c5_n1.json
cp $G/b8_bench_c5_ref.json $P/r2_bench_c5_reference_arm.json
cp $G/b8_micro.log $P/r2_microbench.log
cp $G/b8_raster.log $P/r2_raster_after.log
cp $G/b8_siam.log $P/r2_siam_correlation.log
cp $G/b8_misc.log $P/r2_secondary_kernels.log
cp $G/b8_e2e_chunks.log $P/r2_e2e_chunks.log
( tail -3 $G/b8_pytest.log; tail -1 $G/b8_smoke.log ) > $P/r2_gpu_tests.log
[ -f $G/b8_launches.csv ] && cp $G/b8_launches.csv $P/r2_launches_bench.csv
[ -f $G/b8_micro_pipes.csv ] && cp $G/b8_micro_pipes.csv $P/r2_microbench_pipes.csv
reps=$(ls $G/prof_r2c_*.ncu-rep 2>/dev/null)
if [ -n "$reps" ]; then
  python tools/ncu_summary.py $P/r2_ncu_full_summary.json $reps
  src() {  # report kernel-regex file.cu out [launch index]
    [ -f $1 ] || return
    ( python tools/ncu_src.py $1 $2; echo; echo "per source line:"; python tools/ncu_lines.py $1 $2 $3 30 ${5:-0} ) > $P/$4 2>&1
  }
  src $G/prof_r2c_step.ncu-rep maxplus_stream maxplus.cu r2_maxplus_source_summary.txt
  src $G/prof_r2c_step.ncu-rep mask_select select.cu r2_mask_select_source_summary.txt
  src $G/prof_r2c_raster.ncu-rep raster_kernel raster.cu r2_raster_after_source_summary.txt
  src $G/prof_r2c_siam_tc.ncu-rep siam_tc siam_tc.cu r2_siam_tc_source_summary.txt
  src $G/prof_r2c_u8.ncu-rep maxplus_u8 maxplus_u8.cu r2_maxplus_u8_source_summary.txt
  src $G/prof_r2c_env.ncu-rep raster_kernel raster.cu r2_env_wall_raster_source_summary.txt
  src $G/prof_r2c_env.ncu-rep mask_select select.cu r2_env_mask_select_source_summary.txt
  src $G/prof_r2c_env.ncu-rep pack_rewards envstep.cu r2_env_pack_rewards_source_summary.txt
fi
python tools/sass_excerpt.py raster.cu raster_kernel FMUL2 56 > $P/r2_sass_raster_own_lane_shading.txt
python tools/sass_excerpt.py maxplus.cu 'maxplus_stream_kernel<\(int\)17, \(int\)16>' VIMNMX3 64 > $P/r2_sass_maxplus_float_sweep.txt
python tools/sass_excerpt.py maxplus.cu 'maxplus_stream_kernel<\(int\)17, \(int\)16>' VIADDMNMX 64 > $P/r2_sass_maxplus_fixed_point_sweep.txt
python tools/sass_excerpt.py maxplus_u8.cu 'maxplus_u8_tile_kernel<\(int\)16, \(int\)9>' VIADDMNMX 64 > $P/r2_sass_maxplus_u8_sweep.txt
python tools/sass_excerpt.py siam_tc.cu 'siam_tc_kernel<\(bool\)1>' UTCHMMA 48 > $P/r2_sass_siam_tc_mma_loop.txt
python tools/sass_excerpt.py select.cu 'mask_select_packed_kernel<float, float.*2149581832' POPC 48 > $P/r2_sass_mask_select_overlap.txt
ls -la $P | tail -50
