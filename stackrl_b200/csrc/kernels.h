// Host-side dispatchers implemented by the .cu files, called from capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "stackrl_b200.h"

namespace srl {

int maxplus_f32(const float* walls, const float* rocks, const float* level,
                float* out, int E, int R, int H, int W, int h, float threshold,
                int variant, int quantum_log2, cudaStream_t stream);

int goal_overlap_f32(const float* walls, const float* goals, const float* rocks,
                     int32_t* counts, int E, int R, int H, int W, int h,
                     cudaStream_t stream);
int goal_overlap_u8(const uint8_t* walls, const uint8_t* goals, const uint8_t* rocks,
                    int32_t* counts, int E, int R, int H, int W, int h,
                    cudaStream_t stream);

int select_f32(const float* values, const int32_t* counts, int64_t* actions,
               double* shown, int64_t* best, int E, int R, int Ph, int Pw, int minorder,
               double overlap_threshold, cudaStream_t stream);
int select_f64(const double* values, const int32_t* counts, int64_t* actions,
               double* shown, int64_t* best, int E, int R, int Ph, int Pw, int minorder,
               double overlap_threshold, cudaStream_t stream);

int mask_select_f32(const float* values, const float* walls, const float* goals,
                    const float* rocks, int64_t* actions, double* shown, int64_t* best,
                    int E, int R, int H, int W, int h, int minorder, double overlap_threshold,
                    cudaStream_t stream);
int mask_select_f64(const double* values, const float* walls, const float* goals,
                    const float* rocks, int64_t* actions, double* shown, int64_t* best,
                    int E, int R, int H, int W, int h, int minorder, double overlap_threshold,
                    cudaStream_t stream);
int mask_select_f64_u8(const double* values, const uint8_t* walls, const uint8_t* goals,
                       const uint8_t* rocks, int64_t* actions, double* shown, int64_t* best,
                       int E, int R, int H, int W, int h, int minorder,
                       double overlap_threshold, cudaStream_t stream);

int drop_height_f32(const float* walls, const float* rocks, const int32_t* picks,
                    float* out, int E, int R, int H, int W, int h, float threshold,
                    cudaStream_t stream);

int maxplus_u8(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
               double* out, int E, int R, int H, int W, int h, cudaStream_t stream);

int difference_weights(const float* rocks, const float* level, double* weights, int E,
                       int R, int h, int weights_exponent, cudaStream_t stream);
int difference_f32(const float* walls, const float* rocks, const float* level,
                   const double* weights, double* out, float* top, int E, int R, int H,
                   int W, int h, int difference_exponent, cudaStream_t stream);

int difference_weights_u8(const uint8_t* rocks, double* weights, int E, int R, int h,
                          int weights_exponent, cudaStream_t stream);
int difference_u8(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
                  const double* weights, double* out, double* top, int E, int R, int H,
                  int W, int h, int difference_exponent, cudaStream_t stream);

int corrcoef_localized_f32(const float* walls, const float* rocks, const float* level,
                           void* work, double* out, int E, int R, int H, int W, int h,
                           cudaStream_t stream);
int corrcoef_localized_u8(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
                          void* work, double* out, int E, int R, int H, int W, int h,
                          cudaStream_t stream);

int raster(const float* verts, const int32_t* tris, const srl_raster_instance* insts,
           const srl_raster_job* jobs, const int32_t* inst_counts, float* depth_state,
           int only_last, float* out, int njobs, int rows, int cols, int mode,
           double far_plane, int vert_cap_hint, cudaStream_t stream,
           int32_t* rows_out = nullptr);

int pack_obs(const float* walls, const float* goals, const float* rocks, void* wall_goal,
             void* rock, int E, int R, int H, int W, int h, int dtype_code, float scale,
             int repeat_wall, cudaStream_t stream);
int reward_sums_f32(const float* walls, const float* goals, const float* goal_z, float* inter,
                    float* uni, float* vol, int E, int H, int W, cudaStream_t stream);

int score_f32(const float* walls, const float* goals, const float* rocks, const float* level,
              float* values, int64_t* actions, int64_t* best, int E, int R, int H, int W,
              int h, int level_mode, int minorder, double overlap_threshold,
              cudaStream_t stream);

int correlate_f32(const float* walls, const float* rocks, const float* level, float* corr,
                  float* coef, int E, int R, int H, int W, int h, cudaStream_t stream);

int siam_correlation_f32(const float* x, const float* w, float* out, int B, int H, int W,
                         int C, int h, int wd, cudaStream_t stream);

int place_poses_f32(const float* walls, const float* rocks, const int64_t* views,
                    const int64_t* flat, const double* orientations, double* poses,
                    int32_t* status, int E, int R, int H, int W, int h, int action_stride,
                    double pixel_h, double pixel_w, double object_x, double object_y,
                    double object_z, float threshold, cudaStream_t stream);
int contact_precheck_f32(const float* walls, const float* rocks, const int64_t* views,
                         const int64_t* flat, int32_t* contacts, int32_t* octants,
                         uint8_t* supported, int E, int R, int H, int W, int h,
                         int action_stride, float threshold, float eps, cudaStream_t stream);
int env_advance(const srl_env_state* st, const double* rest, const double* placed,
                cudaStream_t stream);
int env_reset(const srl_env_state* st, const int32_t* env_ids, int n, cudaStream_t stream);
int env_set_poses(const srl_env_state* st, const double* poses, int n_given,
                  cudaStream_t stream);
int env_draw(const srl_env_state* st, int32_t* order, int32_t* rects, const int32_t* env_ids,
             int n, int n_meshes, int H, int W, int object_h, int object_w, int goal_mode,
             int goal_size, int goal_size_h, int goal_size_w, unsigned long long seed,
             unsigned long long episode, cudaStream_t stream);
int fill_goals_f32(const int32_t* rects, const float* goal_z, const int32_t* env_ids,
                   float* goals, int n, int H, int W, cudaStream_t stream);
int goal_level_f32(const float* goals, float* level, int E, int HW, cudaStream_t stream);
int goal_level_u8(const uint8_t* goals, uint8_t* level, int E, int HW, cudaStream_t stream);
int rewards_f32(const srl_env_state* st, const float* walls, const float* goals,
                const float* goal_z, const int32_t* rects, float* reward, double* value,
                int H, int W, int metric, double scale, double pixel_h, double pixel_w,
                double pmax, double pexp, double oexp, cudaStream_t stream);
int pack_rewards_f32(const srl_env_state* st, const float* walls, const float* goals,
                     const float* rocks, const float* goal_z, const int32_t* rects,
                     void* wall_goal, void* rock, float* reward, double* value, int R, int H,
                     int W, int h, int dtype_code, float obs_scale, int repeat_wall, int metric,
                     double scale, double pixel_h, double pixel_w, double pmax, double pexp,
                     double oexp, cudaStream_t stream, const int32_t* rows = nullptr,
                     uint8_t* full = nullptr);
int gather_rows_f32(const float* table, const int32_t* index, float* out, int rows_out,
                    int row_floats, int table_rows, cudaStream_t stream);
int quantise_planes_u8(const float* walls, const float* goals, const float* rocks,
                       uint8_t* walls8, uint8_t* goals8, uint8_t* rocks8, int E, int R, int H,
                       int W, int h, float scale, cudaStream_t stream);

int siam_correlation_tc(const float* x, const float* f, float* out, int B, int H, int W, int C,
                        int h, int wd, cudaStream_t stream);

int siam_correlation_grad_f32(const float* x, const float* f, const float* g, float* grad_x,
                              float* grad_f, int B, int H, int W, int C, int h, int wd,
                              cudaStream_t stream);

int microbench_addmax(int variant, int iters, double* host_cells_per_s);

int microbench_fma(int variant, int iters, double* host_fma_per_s);

}  // namespace srl
