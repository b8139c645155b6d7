// extern "C" surface of libstackrl_b200.so (include/stackrl_b200.h) and the
// error plumbing shared by the kernels' host-side dispatchers.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {
thread_local char g_error[512] = "";
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(SRL_E_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return SRL_OK;
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return 0;
  return n;
}

}  // namespace srl

using srl::fail;

extern "C" {

int srl_version(void) { return 100; }  // 0.1.0

const char* srl_last_error(void) { return srl::g_error; }

int srl_device_sm_count(int* host_out) {
  SRL_REQUIRE(host_out != nullptr, SRL_E_INVALID, "srl_device_sm_count: null");
  const int n = srl::sm_count();
  SRL_REQUIRE(n > 0, SRL_E_CUDA, "srl_device_sm_count: no CUDA device");
  *host_out = n;
  return SRL_OK;
}

int srl_maxplus_f32(const float* walls, const float* rocks, const float* level,
                    float* out, int E, int R, int H, int W, int h,
                    float threshold, srl_stream_t stream) {
  int variant = 1;
  if (const char* s = getenv("SRL_MAXPLUS_VARIANT")) variant = atoi(s);
  return srl::maxplus_f32(walls, rocks, level, out, E, R, H, W, h, threshold,
                          variant, SRL_NO_QUANTUM, (cudaStream_t)stream);
}

int srl_maxplus_f32_q(const float* walls, const float* rocks, const float* level,
                      float* out, int E, int R, int H, int W, int h, float threshold,
                      int quantum_log2, srl_stream_t stream) {
  int variant = 1;
  if (const char* s = getenv("SRL_MAXPLUS_VARIANT")) variant = atoi(s);
  return srl::maxplus_f32(walls, rocks, level, out, E, R, H, W, h, threshold,
                          variant, quantum_log2, (cudaStream_t)stream);
}

int srl_maxplus_u8(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
                   double* out, int E, int R, int H, int W, int h, srl_stream_t stream) {
  return srl::maxplus_u8(walls, rocks, level, out, E, R, H, W, h, (cudaStream_t)stream);
}

int srl_drop_height_f32(const float* walls, const float* rocks, const int32_t* picks,
                        float* out, int E, int R, int H, int W, int h, float threshold,
                        srl_stream_t stream) {
  return srl::drop_height_f32(walls, rocks, picks, out, E, R, H, W, h, threshold,
                              (cudaStream_t)stream);
}

int srl_goal_overlap_f32(const float* walls, const float* goals, const float* rocks,
                         int32_t* counts, int E, int R, int H, int W, int h,
                         srl_stream_t stream) {
  return srl::goal_overlap_f32(walls, goals, rocks, counts, E, R, H, W, h,
                               (cudaStream_t)stream);
}

int srl_goal_overlap_u8(const uint8_t* walls, const uint8_t* goals, const uint8_t* rocks,
                        int32_t* counts, int E, int R, int H, int W, int h,
                        srl_stream_t stream) {
  return srl::goal_overlap_u8(walls, goals, rocks, counts, E, R, H, W, h,
                              (cudaStream_t)stream);
}

int srl_select_f32(const float* values, const int32_t* counts, int64_t* actions,
                   double* shown, int64_t* best, int E, int R, int Ph, int Pw,
                   int minorder, double overlap_threshold, srl_stream_t stream) {
  return srl::select_f32(values, counts, actions, shown, best, E, R, Ph, Pw, minorder,
                         overlap_threshold, (cudaStream_t)stream);
}

int srl_select_f64(const double* values, const int32_t* counts, int64_t* actions,
                   double* shown, int64_t* best, int E, int R, int Ph, int Pw,
                   int minorder, double overlap_threshold, srl_stream_t stream) {
  return srl::select_f64(values, counts, actions, shown, best, E, R, Ph, Pw, minorder,
                         overlap_threshold, (cudaStream_t)stream);
}

int srl_difference_weights(const float* rocks, const float* level, double* weights,
                           int E, int R, int h, int weights_exponent,
                           srl_stream_t stream) {
  return srl::difference_weights(rocks, level, weights, E, R, h, weights_exponent,
                                 (cudaStream_t)stream);
}

int srl_difference_f32(const float* walls, const float* rocks, const float* level,
                       const double* weights, double* out, float* top, int E, int R,
                       int H, int W, int h, int difference_exponent,
                       srl_stream_t stream) {
  return srl::difference_f32(walls, rocks, level, weights, out, top, E, R, H, W, h,
                             difference_exponent, (cudaStream_t)stream);
}

int srl_difference_weights_u8(const uint8_t* rocks, double* weights, int E, int R, int h,
                              int weights_exponent, srl_stream_t stream) {
  return srl::difference_weights_u8(rocks, weights, E, R, h, weights_exponent,
                                    (cudaStream_t)stream);
}

int srl_difference_u8(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
                      const double* weights, double* out, double* top, int E, int R, int H,
                      int W, int h, int difference_exponent, srl_stream_t stream) {
  return srl::difference_u8(walls, rocks, level, weights, out, top, E, R, H, W, h,
                            difference_exponent, (cudaStream_t)stream);
}

int srl_corrcoef_localized_f32(const float* walls, const float* rocks, const float* level,
                               void* work, double* out, int E, int R, int H, int W, int h,
                               srl_stream_t stream) {
  return srl::corrcoef_localized_f32(walls, rocks, level, work, out, E, R, H, W, h,
                                     (cudaStream_t)stream);
}

int srl_corrcoef_localized_u8(const uint8_t* walls, const uint8_t* rocks,
                              const uint8_t* level, void* work, double* out, int E, int R,
                              int H, int W, int h, srl_stream_t stream) {
  return srl::corrcoef_localized_u8(walls, rocks, level, work, out, E, R, H, W, h,
                                    (cudaStream_t)stream);
}

int srl_raster(const float* verts, const int32_t* tris, const srl_raster_instance* insts,
               const srl_raster_job* jobs, float* out, int njobs, int rows, int cols,
               int mode, double far_plane, srl_stream_t stream) {
  return srl::raster(verts, tris, insts, jobs, nullptr, nullptr, 0, out, njobs, rows, cols,
                     mode, far_plane, 0, (cudaStream_t)stream);
}

int srl_raster_ex(const float* verts, const int32_t* tris, const srl_raster_instance* insts,
                  const srl_raster_job* jobs, const int32_t* inst_counts, float* out,
                  int njobs, int rows, int cols, int mode, double far_plane,
                  int max_cached_verts, srl_stream_t stream) {
  return srl::raster(verts, tris, insts, jobs, inst_counts, nullptr, 0, out, njobs, rows, cols,
                     mode, far_plane, max_cached_verts, (cudaStream_t)stream);
}

int srl_raster_incremental(const float* verts, const int32_t* tris,
                           const srl_raster_instance* insts, const srl_raster_job* jobs,
                           const int32_t* inst_counts, float* depth_state, int only_last,
                           float* out, int njobs, int rows, int cols, int mode,
                           double far_plane, int max_cached_verts, srl_stream_t stream) {
  SRL_REQUIRE(depth_state != nullptr || njobs == 0, SRL_E_INVALID,
              "raster_incremental: depth_state is null");
  return srl::raster(verts, tris, insts, jobs, inst_counts, depth_state, only_last, out, njobs,
                     rows, cols, mode, far_plane, max_cached_verts, (cudaStream_t)stream);
}

int srl_reward_sums_f32(const float* walls, const float* goals, const float* goal_z,
                        float* inter, float* uni, float* vol, int E, int H, int W,
                        srl_stream_t stream) {
  return srl::reward_sums_f32(walls, goals, goal_z, inter, uni, vol, E, H, W,
                              (cudaStream_t)stream);
}

int srl_pack_obs(const float* walls, const float* goals, const float* rocks,
                 void* wall_goal, void* rock, int E, int R, int H, int W, int h,
                 int dtype_code, float scale, int repeat_wall, srl_stream_t stream) {
  return srl::pack_obs(walls, goals, rocks, wall_goal, rock, E, R, H, W, h, dtype_code,
                       scale, repeat_wall, (cudaStream_t)stream);
}

int srl_score_f32(const float* walls, const float* goals, const float* rocks,
                  const float* level, float* values, int64_t* actions, int64_t* best, int E,
                  int R, int H, int W, int h, int level_mode, int minorder,
                  double overlap_threshold, srl_stream_t stream) {
  return srl::score_f32(walls, goals, rocks, level, values, actions, best, E, R, H, W, h,
                        level_mode, minorder, overlap_threshold, (cudaStream_t)stream);
}

int srl_mask_select_f32(const float* values, const float* walls, const float* goals,
                        const float* rocks, int64_t* actions, double* shown, int64_t* best,
                        int E, int R, int H, int W, int h, int minorder,
                        double overlap_threshold, srl_stream_t stream) {
  return srl::mask_select_f32(values, walls, goals, rocks, actions, shown, best, E, R, H, W, h,
                              minorder, overlap_threshold, (cudaStream_t)stream);
}

int srl_mask_select_f64(const double* values, const float* walls, const float* goals,
                        const float* rocks, int64_t* actions, double* shown, int64_t* best,
                        int E, int R, int H, int W, int h, int minorder,
                        double overlap_threshold, srl_stream_t stream) {
  return srl::mask_select_f64(values, walls, goals, rocks, actions, shown, best, E, R, H, W, h,
                              minorder, overlap_threshold, (cudaStream_t)stream);
}

int srl_mask_select_f64_u8(const double* values, const uint8_t* walls, const uint8_t* goals,
                           const uint8_t* rocks, int64_t* actions, double* shown,
                           int64_t* best, int E, int R, int H, int W, int h, int minorder,
                           double overlap_threshold, srl_stream_t stream) {
  return srl::mask_select_f64_u8(values, walls, goals, rocks, actions, shown, best, E, R, H, W,
                                 h, minorder, overlap_threshold, (cudaStream_t)stream);
}

int srl_correlate_f32(const float* walls, const float* rocks, const float* level,
                      float* corr, float* coef, int E, int R, int H, int W, int h,
                      srl_stream_t stream) {
  return srl::correlate_f32(walls, rocks, level, corr, coef, E, R, H, W, h,
                            (cudaStream_t)stream);
}

int srl_siam_correlation_f32(const float* x, const float* w, float* out, int B, int H, int W,
                             int C, int h, int wd, srl_stream_t stream) {
  return srl::siam_correlation_f32(x, w, out, B, H, W, C, h, wd, (cudaStream_t)stream);
}

int srl_place_poses_f32(const float* walls, const float* rocks, const int64_t* views,
                        const int64_t* flat, const double* orientations, double* poses,
                        int32_t* status, int E, int R, int H, int W, int h,
                        int action_stride, double pixel_h, double pixel_w, double object_x,
                        double object_y, double object_z, float threshold,
                        srl_stream_t stream) {
  return srl::place_poses_f32(walls, rocks, views, flat, orientations, poses, status, E, R, H,
                              W, h, action_stride, pixel_h, pixel_w, object_x, object_y,
                              object_z, threshold, (cudaStream_t)stream);
}

int srl_contact_precheck_f32(const float* walls, const float* rocks, const int64_t* views,
                             const int64_t* flat, int32_t* contacts, int32_t* octants,
                             uint8_t* supported, int E, int R, int H, int W, int h,
                             int action_stride, float threshold, float eps,
                             srl_stream_t stream) {
  return srl::contact_precheck_f32(walls, rocks, views, flat, contacts, octants, supported, E, R,
                                   H, W, h, action_stride, threshold, eps,
                                   (cudaStream_t)stream);
}

int srl_env_reset(const srl_env_state* host_state, const int32_t* env_ids, int n,
                  srl_stream_t stream) {
  return srl::env_reset(host_state, env_ids, n, (cudaStream_t)stream);
}

int srl_env_advance(const srl_env_state* host_state, const double* rest, const double* placed,
                    srl_stream_t stream) {
  return srl::env_advance(host_state, rest, placed, (cudaStream_t)stream);
}

int srl_env_set_poses(const srl_env_state* host_state, const double* poses, int n_given,
                      srl_stream_t stream) {
  return srl::env_set_poses(host_state, poses, n_given, (cudaStream_t)stream);
}

int srl_env_draw(const srl_env_state* host_state, int32_t* order, int32_t* rects,
                 const int32_t* env_ids, int n, int n_meshes, int H, int W, int object_h,
                 int object_w, int goal_mode, int goal_size, int goal_size_h, int goal_size_w,
                 uint64_t seed, uint64_t episode, srl_stream_t stream) {
  return srl::env_draw(host_state, order, rects, env_ids, n, n_meshes, H, W, object_h, object_w,
                       goal_mode, goal_size, goal_size_h, goal_size_w, seed, episode,
                       (cudaStream_t)stream);
}

int srl_fill_goals_f32(const int32_t* rects, const float* goal_z, const int32_t* env_ids,
                       float* goals, int n, int H, int W, srl_stream_t stream) {
  return srl::fill_goals_f32(rects, goal_z, env_ids, goals, n, H, W, (cudaStream_t)stream);
}

int srl_goal_level_f32(const float* goals, float* level, int E, int HW, srl_stream_t stream) {
  return srl::goal_level_f32(goals, level, E, HW, (cudaStream_t)stream);
}

int srl_goal_level_u8(const uint8_t* goals, uint8_t* level, int E, int HW,
                      srl_stream_t stream) {
  return srl::goal_level_u8(goals, level, E, HW, (cudaStream_t)stream);
}

int srl_rewards_f32(const srl_env_state* host_state, const float* walls, const float* goals,
                    const float* goal_z, const int32_t* rects, float* reward, double* value,
                    int H, int W, int metric, double scale, double pixel_h, double pixel_w,
                    double pmax, double pexp, double oexp, srl_stream_t stream) {
  return srl::rewards_f32(host_state, walls, goals, goal_z, rects, reward, value, H, W, metric,
                          scale, pixel_h, pixel_w, pmax, pexp, oexp, (cudaStream_t)stream);
}

int srl_pack_rewards_f32(const srl_env_state* host_state, const float* walls,
                         const float* goals, const float* rocks, const float* goal_z,
                         const int32_t* rects, void* wall_goal, void* rock, float* reward,
                         double* value, int R, int H, int W, int h, int dtype_code,
                         float obs_scale, int repeat_wall, int metric, double scale,
                         double pixel_h, double pixel_w, double pmax, double pexp, double oexp,
                         srl_stream_t stream) {
  return srl::pack_rewards_f32(host_state, walls, goals, rocks, goal_z, rects, wall_goal, rock,
                               reward, value, R, H, W, h, dtype_code, obs_scale, repeat_wall,
                               metric, scale, pixel_h, pixel_w, pmax, pexp, oexp,
                               (cudaStream_t)stream);
}

int srl_quantise_planes_u8(const float* walls, const float* goals, const float* rocks,
                           uint8_t* walls8, uint8_t* goals8, uint8_t* rocks8, int E, int R,
                           int H, int W, int h, float scale, srl_stream_t stream) {
  return srl::quantise_planes_u8(walls, goals, rocks, walls8, goals8, rocks8, E, R, H, W, h,
                                 scale, (cudaStream_t)stream);
}

int srl_raster_incremental_rows(const float* verts, const int32_t* tris,
                                const srl_raster_instance* insts, const srl_raster_job* jobs,
                                const int32_t* inst_counts, float* depth_state, int only_last,
                                float* out, int32_t* rows_out, int njobs, int rows, int cols,
                                int mode, double far_plane, int max_cached_verts,
                                srl_stream_t stream) {
  SRL_REQUIRE(depth_state != nullptr || njobs == 0, SRL_E_INVALID,
              "raster_incremental_rows: depth_state is null");
  SRL_REQUIRE(rows_out != nullptr || njobs == 0, SRL_E_INVALID,
              "raster_incremental_rows: rows_out is null");
  return srl::raster(verts, tris, insts, jobs, inst_counts, depth_state, only_last, out, njobs,
                     rows, cols, mode, far_plane, max_cached_verts, (cudaStream_t)stream,
                     rows_out);
}

int srl_pack_rewards_rows_f32(const srl_env_state* host_state, const float* walls,
                              const float* goals, const float* rocks, const float* goal_z,
                              const int32_t* rects, const int32_t* rows, uint8_t* full,
                              void* wall_goal, void* rock, float* reward, double* value, int R,
                              int H, int W, int h, int dtype_code, float obs_scale,
                              int repeat_wall, int metric, double scale, double pixel_h,
                              double pixel_w, double pmax, double pexp, double oexp,
                              srl_stream_t stream) {
  SRL_REQUIRE(rows != nullptr || host_state == nullptr || host_state->E == 0, SRL_E_INVALID,
              "pack_rewards_rows: rows is null");
  return srl::pack_rewards_f32(host_state, walls, goals, rocks, goal_z, rects, wall_goal, rock,
                               reward, value, R, H, W, h, dtype_code, obs_scale, repeat_wall,
                               metric, scale, pixel_h, pixel_w, pmax, pexp, oexp,
                               (cudaStream_t)stream, rows, full);
}

int srl_gather_rows_f32(const float* table, const int32_t* index, float* out, int rows_out,
                        int row_floats, int table_rows, srl_stream_t stream) {
  return srl::gather_rows_f32(table, index, out, rows_out, row_floats, table_rows,
                              (cudaStream_t)stream);
}

int srl_siam_correlation_grad_f32(const float* x, const float* w, const float* grad_out,
                                  float* grad_x, float* grad_w, int B, int H, int W, int C,
                                  int h, int wd, srl_stream_t stream) {
  return srl::siam_correlation_grad_f32(x, w, grad_out, grad_x, grad_w, B, H, W, C, h, wd,
                                        (cudaStream_t)stream);
}

int srl_microbench_addmax(int variant, int iters, double* host_cells_per_s) {
  return srl::microbench_addmax(variant, iters, host_cells_per_s);
}

int srl_microbench_fma(int variant, int iters, double* host_fma_per_s) {
  return srl::microbench_fma(variant, iters, host_fma_per_s);
}

}  // extern "C"
