/*
 * stackrl_b200 -- C ABI of the B200-native observation + placement-scoring path.
 *
 * The reference (menezesandre/stackrl) is pure Python and has no FFI of its own;
 * its extension points are Python constructor injection and duck typing
 * (SURVEY.md section 8b).  Each entry point below replaces the arithmetic of one
 * reference function and is what a reference-side ctypes binding would call
 * (INTEGRATION.md shows the stubs).  Citations are relative to the reference
 * tree root.
 *
 * Conventions (all entry points)
 *   - return 0 on success, a negative SRL_E_* code on failure; never throws.
 *     srl_last_error() returns a thread-local description of the last failure.
 *   - every pointer is a DEVICE pointer unless the parameter is named host_*;
 *     the caller owns all buffers (the library neither allocates nor frees
 *     device memory) and has made the right device current.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream) and the call returns without synchronising.
 *   - layouts are dense row-major; "planar" maps are [count, rows, cols].
 *   - P = (H-h+1)*(W-h+1) candidate positions per map, row-major (the reference
 *     only defines square rocks, baselines.py:39 -- SURVEY quirk Q3).
 */
#ifndef STACKRL_B200_H_
#define STACKRL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRL_OK            0
#define SRL_E_INVALID    -1   /* bad argument (null pointer, size, alignment) */
#define SRL_E_UNSUPPORTED -2  /* shape outside what the kernels are built for */
#define SRL_E_CUDA       -3   /* a CUDA runtime call failed (see srl_last_error) */

typedef void* srl_stream_t;   /* cudaStream_t */

#if defined(__GNUC__)
#define SRL_API __attribute__((visibility("default")))
#else
#define SRL_API
#endif

/* Library ABI version (major*10000 + minor*100 + patch). */
SRL_API int srl_version(void);
/* Thread-local message for the last non-zero return on this thread. */
SRL_API const char* srl_last_error(void);
/* Number of SMs of the current device (148 on B200). */
SRL_API int srl_device_sm_count(int* host_out);

/* ---- a5: baselines.get_inputs + baselines.height (baselines.py:21-43) -------
 * out[e,r,i,j] = max_{u,v}( n > threshold ? o[e,i+u,j+v] + n : 0 ),
 *   o = walls[e]/level[e],  n = rocks[e,r,u,v]/level[e]   (IEEE float32 div, add)
 * level == NULL skips the normalisation (o = wall, n = rock), which with
 * threshold = 1e-4 is the Observer.pose mask (observer.py:405-409).
 * A masked cell contributes 0 (not -inf): baselines.py:37-41, quirk Q2.
 *   walls [E,H,W] f32, rocks [E,R,h,h] f32, level [E] f32 or NULL,
 *   out [E,R,H-h+1,W-h+1] f32.  Bit-exact with the reference's float32 path. */
SRL_API int srl_maxplus_f32(const float* walls, const float* rocks, const float* level,
                    float* out, int E, int R, int H, int W, int h,
                    float threshold, srl_stream_t stream);

/* Same maps, with a hint: the caller expects every wall and rock value to be a
 * non-negative multiple of 2^quantum_log2 below 2^(quantum_log2 + 14) -- true for
 * heightmaps that come from the reference's float32 depth->elevation formulas
 * (observer.py:259-260, 274-275: multiples of ulp(1000) = 2^-14 m).  Environments
 * for which that holds (checked value by value on the device) and whose level is a
 * power of two are swept in exact 16-bit fixed point (VIADDMNMX.S16x2, two cells
 * per instruction); all others take the float32 path.  Results are bit-identical
 * to srl_maxplus_f32 either way.  SRL_NO_QUANTUM disables the hint. */
#define SRL_NO_QUANTUM 0x7fffffff
SRL_API int srl_maxplus_f32_q(const float* walls, const float* rocks, const float* level,
                              float* out, int E, int R, int H, int W, int h,
                              float threshold, int quantum_log2, srl_stream_t stream);

/* Same for uint8 observations (registered Stack-v0/1/2 dtype): the reference
 * divides uint8 by uint8 -> float64, so this evaluates IEEE float64
 * a/g + b/g per cell.  level [E] u8 (goal.max()), out f64. */
SRL_API int srl_maxplus_u8(const uint8_t* walls, const uint8_t* rocks,
                   const uint8_t* level, double* out, int E, int R, int H,
                   int W, int h, srl_stream_t stream);

/* ---- a4: Observer.pose drop height (observer.py:401-409) --------------------
 * z[e] = max( (walls[e, i:i+h, j:j+h] + rocks[e, r])[rocks[e, r] > 1e-4] ),
 * (r, i, j) = picks[e] (int32 triples).  out [E] f32. */
SRL_API int srl_drop_height_f32(const float* walls, const float* rocks,
                        const int32_t* picks, float* out, int E, int R, int H,
                        int W, int h, float threshold, srl_stream_t stream);

/* ---- a7: baselines.goal_overlap counts (baselines.py:152-155) ---------------
 * counts[e,r,i,j] = sum_{u,v} (walls[e,i+u,j+v] < goals[e,i+u,j+v]) *
 *                             (rocks[e,r,u,v] > 0)        (exact int32)
 * The >= threshold*max compare is part of srl_select. */
SRL_API int srl_goal_overlap_f32(const float* walls, const float* goals,
                         const float* rocks, int32_t* counts, int E, int R,
                         int H, int W, int h, srl_stream_t stream);
SRL_API int srl_goal_overlap_u8(const uint8_t* walls, const uint8_t* goals,
                        const uint8_t* rocks, int32_t* counts, int E, int R,
                        int H, int W, int h, srl_stream_t stream);

/* ---- a9/a10: Baseline.call + PyGreedy batchwise (baselines.py:201-217,
 *      agents/policies.py:57-91) ----------------------------------------------
 * Per map (e,r): mask = counts >= overlap_threshold*max(counts) (float64
 * compare, quirk Q7); minima = mask & (minimum_filter(values, 1+2*minorder,
 * zero padded) == values); action = first-index argmin over minima if any, else
 * over mask; shown = -where(mask, values, max(values[mask]) + 0.001).
 * counts == NULL is the goal=False branch: action = argmin(values), shown =
 * -values.  Then per env: best_r = first argmax_r shown[e,r,action[e,r]].
 *   values [E,R,P] (f32 or f64), counts [E,R,P] i32 or NULL,
 *   actions [E,R] i64, shown [E,R,P] f64 or NULL (the reference's value map is
 *   float64 even for float32 scores: `max + 0.001` is a float64 add),
 *   best [E,2] i64 = (r*, action[e,r*]) or NULL. */
SRL_API int srl_select_f32(const float* values, const int32_t* counts, int64_t* actions,
                   double* shown, int64_t* best, int E, int R, int Ph, int Pw,
                   int minorder, double overlap_threshold, srl_stream_t stream);
SRL_API int srl_select_f64(const double* values, const int32_t* counts, int64_t* actions,
                   double* shown, int64_t* best, int E, int R, int Ph, int Pw,
                   int minorder, double overlap_threshold, srl_stream_t stream);

/* ---- a7+a9+a10 in one launch: goal_overlap feeding Baseline.call ---------------
 * Same results as srl_goal_overlap_* followed by srl_select_*, but the overlap
 * counts are produced and consumed in shared memory (they never reach HBM).
 * values come from srl_maxplus_f32 (f32), srl_maxplus_u8 / srl_difference_f32
 * (f64); walls/goals/rocks are the RAW observation planes (baselines.py:153-154). */
SRL_API int srl_mask_select_f32(const float* values, const float* walls, const float* goals,
                                const float* rocks, int64_t* actions, double* shown,
                                int64_t* best, int E, int R, int H, int W, int h,
                                int minorder, double overlap_threshold, srl_stream_t stream);
SRL_API int srl_mask_select_f64(const double* values, const float* walls, const float* goals,
                                const float* rocks, int64_t* actions, double* shown,
                                int64_t* best, int E, int R, int H, int W, int h,
                                int minorder, double overlap_threshold, srl_stream_t stream);
SRL_API int srl_mask_select_f64_u8(const double* values, const uint8_t* walls,
                                   const uint8_t* goals, const uint8_t* rocks,
                                   int64_t* actions, double* shown, int64_t* best, int E,
                                   int R, int H, int W, int h, int minorder,
                                   double overlap_threshold, srl_stream_t stream);

/* ---- a5+a7+a9+a10 fused: Baseline('height', batched, batchwise).__call__ -----
 * (baselines.py:201-217 over :21-43 and :152-156; agents/policies.py:57-91).
 * One launch computes the max-plus score maps, the goal-overlap mask, the masked
 * local-minimum arg-min per view and the batch-wise pick; overlap counts stay in
 * shared memory.  Results are identical to srl_maxplus_f32 + srl_goal_overlap_f32
 * + srl_select_f32.
 *   level_mode 0: no normalisation; 1: level[e]; 2: max(goals[e]) computed
 *   in-kernel (get_inputs, baselines.py:23; goal heights must be >= 0).
 *   goals == NULL is goal=False (plain arg-min).  values [E,R,P] f32 may be NULL
 *   (score maps not written); actions [E,R] i64; best [E,2] i64 or NULL.
 * Returns SRL_E_UNSUPPORTED for shapes outside the fused kernel (rows not
 * multiples of 16 B, more than 288 (view,row,strip) items per environment, or
 * maps that do not fit shared memory): use the separate entry points then. */
SRL_API int srl_score_f32(const float* walls, const float* goals, const float* rocks,
                          const float* level, float* values, int64_t* actions,
                          int64_t* best, int E, int R, int H, int W, int h,
                          int level_mode, int minorder, double overlap_threshold,
                          srl_stream_t stream);

/* ---- a6: baselines.difference (baselines.py:45-77), exponents (2, 2|0) -------
 * f = sum_{u,v} w[u,v] * |h0 - (o+n)|^p  in numpy's order: float32 lift and
 * residual, float64 weights/product, pairwise summation over the contiguous
 * [h,h] block.  p in {1, 2}.  weights [E,R,h,h] f64 come from
 * srl_difference_weights.  out [E,R,P] f64; top (h0) [E,R,P] f32 or NULL (when
 * given, h0 comes from the max-plus kernel -- same bits, much faster than the
 * in-kernel scalar pass used otherwise). */
SRL_API int srl_difference_weights(const float* rocks, const float* level,
                           double* weights, int E, int R, int h,
                           int weights_exponent, srl_stream_t stream);
SRL_API int srl_difference_f32(const float* walls, const float* rocks,
                       const float* level, const double* weights, double* out,
                       float* top, int E, int R, int H, int W, int h,
                       int difference_exponent, srl_stream_t stream);
/* Same for uint8 observations: every step of baselines.py:64-69 is float64 then
 * (uint8 / uint8 -> float64 in get_inputs); level [E] u8, top [E,R,P] f64 or NULL. */
SRL_API int srl_difference_weights_u8(const uint8_t* rocks, double* weights, int E, int R,
                                      int h, int weights_exponent, srl_stream_t stream);
SRL_API int srl_difference_u8(const uint8_t* walls, const uint8_t* rocks,
                              const uint8_t* level, const double* weights, double* out,
                              double* top, int E, int R, int H, int W, int h,
                              int difference_exponent, srl_stream_t stream);

/* ---- a8: baselines.correlate / baselines.corrcoef (baselines.py:141-143, 79-85) --
 * corr = correlate2d(o, n, 'valid') / n.sum(); coef = TM_CCOEFF_NORMED(o, n), with
 * o, n the normalised maps.  The reference values come from scipy / OpenCV library
 * code (summation order outside the reference tree): matched to a tolerance, not
 * bit for bit (float64 window sums, rounded once).  Either output may be NULL. */
SRL_API int srl_correlate_f32(const float* walls, const float* rocks, const float* level,
                              float* corr, float* coef, int E, int R, int H, int W, int h,
                              srl_stream_t stream);

/* ---- the Siamese correlation layer of the DQN ------------------
 * Replaces stackrl.nets.correlation (nets/layers.py:21-38; called from
 * nets/models.py:89 and :182), i.e. per sample
 *   tf.nn.conv2d(x[b][None], w[b][..., None], strides=1, padding='VALID'):
 *   out[b,i,j] = sum_{u,v,c} x[b,i+u,j+v,c] * w[b,u,v,c]
 * x [B,H,W,C], w [B,h,wd,C] float32 channels-last (the reference's tensor layout),
 * out [B,H-h+1,W-wd+1] (the reference's trailing unit channel is a view).  Shapes with
 * C % 8 == 0 and h <= 32 that fit shared memory run on the tensor cores (tcgen05.mma,
 * kind::tf32, every operand split hi + lo: three products, float32 accumulation in
 * TMEM per filter row, float32 sum over the filter rows); everything else on the FP32
 * FMA pipe.  Inputs must be finite.
 * Matched to a tolerance (1e-5 of the largest output): TensorFlow's own summation
 * order is outside the reference tree. */
SRL_API int srl_siam_correlation_f32(const float* x, const float* w, float* out, int B,
                                     int H, int W, int C, int h, int wd,
                                     srl_stream_t stream);

/* The two vector-Jacobian products of the layer (the DQN trains through it,
 * nets/models.py:89, 182):  grad_w[b,u,v,c] = sum_{i,j} grad_out[b,i,j] x[b,i+u,j+v,c],
 * grad_x[b,r,s,c] = sum_{u,v} grad_out[b,r-u,s-v] w[b,u,v,c].  grad_out [B,H-h+1,W-wd+1];
 * either output may be NULL (then the matching input may be NULL too).  float32. */
SRL_API int srl_siam_correlation_grad_f32(const float* x, const float* w,
                                          const float* grad_out, float* grad_x, float* grad_w,
                                          int B, int H, int W, int C, int h, int wd,
                                          srl_stream_t stream);

/* corrcoef(localized=True) (baselines.py:87-114): the masked variant, every sum in
 * numpy's pairwise order and in the observation's arithmetic type (float32, or
 * float64 for uint8): bit-exact.  work: caller-owned scratch of
 * E*R*(h*h + 2) elements of that type (4 or 8 bytes each); out [E,R,P] f64. */
SRL_API int srl_corrcoef_localized_f32(const float* walls, const float* rocks,
                                       const float* level, void* work, double* out, int E,
                                       int R, int H, int W, int h, srl_stream_t stream);
SRL_API int srl_corrcoef_localized_u8(const uint8_t* walls, const uint8_t* rocks,
                                      const uint8_t* level, void* work, double* out, int E,
                                      int R, int H, int W, int h, srl_stream_t stream);

/* ---- a2/a3: Observer.__call__ rasterisation + depth->elevation
 *      (observer.py:252-260, 267-277; pybullet.getCameraImage) ----------------
 * One job = one image: a camera (column-major GL view and projection matrices,
 * as computeViewMatrix / computeProjectionMatrix define them, in float64) and a
 * contiguous range of instances; one instance = a mesh of the bank (vertex and
 * triangle ranges, triangle indices local to the mesh) placed by a rotation
 * matrix (row-major) and a translation.  The reference delegates this step to
 * Bullet's TinyRenderer, which is outside its tree: the exact sampling rules
 * are the ones of oracle/csrc/oracle.c (DESIGN.md "raster").  Depth follows the
 * GL convention the reference inverts (observer.py:259-260); `mode` selects the
 * fused conversion, evaluated in float32 in numpy's operation order:
 *   SRL_RASTER_DEPTH  out = depth in [0,1] (background 1)
 *   SRL_RASTER_WALL   out = far - far*(far-zr)/(far - zr*d)            (:259-260)
 *   SRL_RASTER_ROCK   out = far + zr/2 - (far^2-(zr/2)^2)/(far + zr*(1/2-d)),
 *                           columns mirrored                            (:274-277)
 *   verts [nverts,3] f32; tris [ntris,3] i32; insts, jobs: device arrays of the
 *   structs below; out [njobs, rows, cols] f32. */
#define SRL_RASTER_DEPTH 0
#define SRL_RASTER_WALL  1
#define SRL_RASTER_ROCK  2
typedef struct srl_raster_instance {
  double rot[9];         /* row-major rotation, mesh frame -> world */
  double pos[3];
  int32_t vert_begin, vert_count, tri_begin, tri_count;
} srl_raster_instance;
typedef struct srl_raster_job {
  double view[16];       /* column-major */
  double proj[16];       /* column-major */
  int32_t inst_begin, inst_count;
  double zrange;         /* zr above: max_z (wall) or object_z (rock) */
} srl_raster_job;
SRL_API int srl_raster(const float* verts, const int32_t* tris,
                       const srl_raster_instance* insts, const srl_raster_job* jobs,
                       float* out, int njobs, int rows, int cols, int mode,
                       double far_plane, srl_stream_t stream);
/* Same, for scenes that live on the device: inst_counts [njobs] i32 (or NULL)
 * overrides jobs[k].inst_count -- the number of placed rocks of each environment
 * changes every step while the job table stays put -- and max_cached_verts (0:
 * default 2048) sizes the per-image screen-space vertex cache in shared memory: the
 * vertex count of the largest mesh (rock images) or of one image's instances (wall
 * images); anything larger still renders, its corners projected per triangle. */
SRL_API int srl_raster_ex(const float* verts, const int32_t* tris,
                          const srl_raster_instance* insts, const srl_raster_job* jobs,
                          const int32_t* inst_counts, float* out, int njobs, int rows,
                          int cols, int mode, double far_plane, int max_cached_verts,
                          srl_stream_t stream);

/* Incremental form for scenes that only GROW between observations (a placed rock
 * stays where it is: the static backend, or a settle step that moved nothing else).
 * depth_state [njobs, rows, cols] f32 is the GL depth image kept between calls.
 * only_last == 0: draw every instance (like srl_raster_ex) and leave the depth image
 * in depth_state; only_last != 0: start from depth_state and draw ONLY the last
 * instance of each job; only_last == 2 additionally promises that `out` still holds
 * the image this function wrote for the current depth_state: only the pixels the new
 * instance changes are then read and written (in place).  The depth image is a minimum over triangles, so the result
 * is bit-identical to re-drawing all instances (the reference re-renders the whole
 * scene every step, observer.py:252-257, because pybullet's renderer has no such
 * mode); the wall image of a 30-rock episode costs one rock per step instead of 15 on
 * average.  With only_last == 2 and 0 < max_cached_verts <= 128 (the true size of the largest
 * mesh: the reference's rocks have at most 68 vertices) one warp draws each image into a
 * window of it -- same bits; a larger mesh met at run time is still drawn correctly. */
SRL_API int srl_raster_incremental(const float* verts, const int32_t* tris,
                                   const srl_raster_instance* insts,
                                   const srl_raster_job* jobs, const int32_t* inst_counts,
                                   float* depth_state, int only_last, float* out, int njobs,
                                   int rows, int cols, int mode, double far_plane,
                                   int max_cached_verts, srl_stream_t stream);

/* srl_raster_incremental that also reports, per job, the image rows it may have changed:
 * rows_out [njobs,2] int32 = (first row, past-last row); (0, rows) for a whole-scene draw,
 * the span of the appended instance's projected vertices (one row of margin) for
 * only_last == 2.  srl_pack_rewards_rows_f32 uses it to rewrite only those rows of a
 * persistent packed observation. */
SRL_API int srl_raster_incremental_rows(const float* verts, const int32_t* tris,
                                        const srl_raster_instance* insts,
                                        const srl_raster_job* jobs,
                                        const int32_t* inst_counts, float* depth_state,
                                        int only_last, float* out, int32_t* rows_out,
                                        int njobs, int rows, int cols, int mode,
                                        double far_plane, int max_cached_verts,
                                        srl_stream_t stream);

/* ---- a11: Rewarder._intersection/_union (rewarder.py:297-307) ---------------
 * inter[e] = sum(min(walls[e][goal != 0], goal_z[e])), uni[e] = sum(max(walls[e],
 * goals[e])), vol[e] = sum(goals[e]).  Accumulated in float64 and rounded once;
 * numpy's float32 pairwise order is not reproduced (tolerance 1e-6 relative,
 * see DESIGN.md). */
SRL_API int srl_reward_sums_f32(const float* walls, const float* goals,
                        const float* goal_z, float* inter, float* uni,
                        float* vol, int E, int H, int W, srl_stream_t stream);

/* ---- a14: StackEnv.observation/_return (env.py:171-180, 226-231, 472-480) ---
 * wall_goal[e,(r),i,j,0:2] = cast(walls[e,i,j], goals[e,i,j]); rock[e,r,u,v,0] =
 * cast(rocks[e,r,u,v]).  dtype_code: 0 float32 (copy), 1 uint8: trunc(x*255/
 * scale) evaluated in float32 like numpy (quirk Q11).  repeat_wall != 0 writes
 * the TestStackEnv layout ([E,R,H,W,2]), else [E,H,W,2]. */
SRL_API int srl_pack_obs(const float* walls, const float* goals, const float* rocks,
                 void* wall_goal, void* rock, int E, int R, int H, int W, int h,
                 int dtype_code, float scale, int repeat_wall,
                 srl_stream_t stream);

/* ---- a4 for a batch: Observer.pose (observer.py:392-421) ------------------------
 * For every environment: (row, col) = divmod(flat[e], W-h+1) (env.py:240-241),
 * z = max((walls[e, row:row+h, col:col+h] + rocks[e, r])[rocks[e, r] > threshold])
 * in float32 with r = views[e] (0 when views == NULL), and
 *   poses[e] = (row*pixel_h + object_x/2, col*pixel_w + object_y/2,
 *               z - object_z/2 [float32], orientations[r][0..3])      float64 [E,7]
 * views/flat are int64 device arrays with `action_stride` elements between
 * consecutive environments (1 for plain [E] arrays; 2 when they are the two columns
 * of the `best` [E,2] table the selection kernels write); orientations [R,4] float64
 * (Observer._object_orientations).
 * The reference asserts action_space.contains(action) (env.py:237): an action out
 * of range reads nothing, yields a NaN pose and sets status[e] = 1 (status [E] i32,
 * may be NULL; the kernel never clears it, so one read after many steps tells
 * whether any of them was invalid). */
SRL_API int srl_place_poses_f32(const float* walls, const float* rocks, const int64_t* views,
                                const int64_t* flat, const double* orientations,
                                double* poses, int32_t* status, int E, int R, int H, int W,
                                int h, int action_stride, double pixel_h, double pixel_w,
                                double object_x, double object_y, double object_z,
                                float threshold, srl_stream_t stream);

/* ---- contact pre-check of a placement from the heightmaps ----------
 * The reference learns whether a dropped rock rests by stepping physics until it
 * has >= 3 contact points (Simulator._drop, simulator.py:337-341).  The maps answer a
 * cheaper version of the question before any physics: with h0 = max((walls[e, window]
 * + rocks[e, r])[rocks > threshold]) (a4 / a5), the cells with h0 - (wall + rock) <=
 * eps -- difference()'s residual field (baselines.py:64-72) thresholded -- are where
 * the rock touches.  contacts [E] i32: their number; octants [E] i32: bit k set when a
 * touching cell lies in octant k (45 degree sectors, counter-clockwise from +row)
 * around the map centre (= the centre of mass in x, y: the object camera looks at the
 * inertial frame, observer.py:148-164); supported [E] u8: contacts >= 3 and no four
 * consecutive empty octants (no empty half-plane through the centre).  Actions as in
 * srl_place_poses_f32; an invalid action gives (0, 0, 0). */
SRL_API int srl_contact_precheck_f32(const float* walls, const float* rocks,
                                     const int64_t* views, const int64_t* flat,
                                     int32_t* contacts, int32_t* octants, uint8_t* supported,
                                     int E, int R, int H, int W, int h, int action_stride,
                                     float threshold, float eps, srl_stream_t stream);

/* ---- a15: episode bookkeeping of StackEnv.step / reset on the device --------------
 * (env.py:233-293 around the physics call; Simulator.positions /
 * distances_from_place, simulator.py:86-127, for the discounted rewards).
 * HOST struct of DEVICE pointers describing E environments:
 *   mesh_ranges [M,4] i32 (vert_begin, vert_count, tri_begin, tri_count), mesh_coms
 *   [M,3] f64 (inertial origin of each URDF), spawn_rows [M] instances at the spawn
 *   pose; instances [E*capacity] + counts [E]: the placed rocks srl_raster draws into
 *   the wall image; order [E,length] i32: mesh of every step of the episode (the
 *   env's episode list, env.py:268-272, in pop order); cursor/current/n_placed [E]
 *   i32; hist_rest / hist_placed [E,length,7] f64 (x,y,z,qx,qy,qz,qw) and hist_mesh
 *   [E,length] i32: rest pose, placement pose and mesh of every placed rock; done [E]
 *   u8; rock_instances [E]: the spawned rock srl_raster draws into the rock images;
 *   memory [E,4] f64: Rewarder._memory (IoU, OR, DIoU, DOR). */
typedef struct srl_env_state {
  int32_t E, capacity, length, reserved;
  const int32_t* mesh_ranges;
  const double* mesh_coms;
  const srl_raster_instance* spawn_rows;
  srl_raster_instance* instances;
  int32_t* counts;
  const int32_t* order;
  int32_t* cursor;
  int32_t* current;
  double* hist_rest;
  double* hist_placed;
  int32_t* hist_mesh;
  int32_t* n_placed;
  uint8_t* done;
  srl_raster_instance* rock_instances;
  double* memory;
} srl_env_state;
/* Start episodes for env_ids [n] i32 (NULL: environments 0..n-1): empty instance
 * table, first rock of `order` spawned, reward memory cleared (env.py:266-293). */
SRL_API int srl_env_reset(const srl_env_state* host_state, const int32_t* env_ids, int n,
                          srl_stream_t stream);
/* One step for every environment that is not done: the spawned rock comes to rest at
 * rest[e] (x,y,z,qx,qy,qz,qw; inertial frame, like resetBasePositionAndOrientation),
 * its instance (rotation from the quaternion, position - R*com) is appended, the
 * history rows are written (placed == NULL: placed = rest), and the next rock of the
 * episode is spawned or the environment is marked done (env.py:245-249). */
SRL_API int srl_env_advance(const srl_env_state* host_state, const double* rest,
                            const double* placed, srl_stream_t stream);
/* Rewrite the rest poses of the first n_given placed rocks of every environment
 * (poses [E,n_given,7]): a physics step that moved earlier rocks. */
SRL_API int srl_env_set_poses(const srl_env_state* host_state, const double* poses,
                              int n_given, srl_stream_t stream);

/* Episode draws on the device, for batches too large for 2E host RandomState streams:
 * order[e, 0..length) = the episode's rock list (env.py:268-272: without replacement
 * when n_meshes >= length) and rects[e] = a goal rectangle of Rewarder._reset_goal
 * (rewarder.py:211-250; goal_mode 0: goal_size_ratio None, 1: scalar -> goal_size =
 * int(ratio*H*W), 2: tuple -> goal_size_h/w = int(ratio_k * H|W)), from a counter-based
 * generator keyed by (seed, episode, e).  Same distributions as the reference, NOT its
 * numpy draw sequence (the per-environment RandomState streams stay on the host:
 * stackrl_b200/episodes.py).  order: [E, length] i32 (normally host_state->order),
 * rects [E,4] i32, both indexed by environment. */
SRL_API int srl_env_draw(const srl_env_state* host_state, int32_t* order, int32_t* rects,
                         const int32_t* env_ids, int n, int n_meshes, int H, int W,
                         int object_h, int object_w, int goal_mode, int goal_size,
                         int goal_size_h, int goal_size_w, uint64_t seed, uint64_t episode,
                         srl_stream_t stream);

/* ---- a13: the goal map of Rewarder._reset_goal (rewarder.py:252-258) ----------------
 * goals[e, u0:u1, v0:v1] = goal_z[e], 0 elsewhere, for e = env_ids[k] (NULL: e = k),
 * rects [n,4] i32 = (u0, v0, u1, v1) = Rewarder._goal_lims.  goal_z is indexed by e. */
SRL_API int srl_fill_goals_f32(const int32_t* rects, const float* goal_z,
                               const int32_t* env_ids, float* goals, int n, int H, int W,
                               srl_stream_t stream);
/* get_inputs' goal.max() (baselines.py:23) per environment: level [E]. */
SRL_API int srl_goal_level_f32(const float* goals, float* level, int E, int HW,
                               srl_stream_t stream);
SRL_API int srl_goal_level_u8(const uint8_t* goals, uint8_t* level, int E, int HW,
                              srl_stream_t stream);

/* ---- a11/a12: Rewarder.call (rewarder.py:162-179, 261-307) ---------------------------
 * metric 0 IoU, 1 OR, 2 DIoU, 3 DOR, 4 all (Rewarder.metrics order).  reward[e] =
 * (value - memory[e][metric]) * scale and memory is updated; metric 4 writes reward
 * [E,4].  value [E,4] f64 (raw metric values) may be NULL.  IoU / OR: float64
 * accumulation of the float32 maps rounded once (1e-6 relative, see
 * srl_reward_sums_f32), then the reference's float32 division.  DIoU / DOR: the
 * sequential float64 sum of rewarder.py:261-295 over hist_rest / hist_placed with
 * Python's float floor division for xy_to_pixel; pexp / oexp < 0 mean "no discount". */
SRL_API int srl_rewards_f32(const srl_env_state* host_state, const float* walls,
                            const float* goals, const float* goal_z, const int32_t* rects,
                            float* reward, double* value, int H, int W, int metric,
                            double scale, double pixel_h, double pixel_w, double pmax,
                            double pexp, double oexp, srl_stream_t stream);

/* a14 + a11/a12 in one pass over the wall and goal maps: srl_pack_obs and
 * srl_rewards_f32 of the same step (what StackEnv.step returns, env.py:255-264) --
 * the observation is packed while the reward sums are taken, so the maps are read
 * once.  Arguments as in those two; obs_scale is srl_pack_obs's `scale`.  `goals` may be
 * NULL when `rects` is given: the goal map of environment e then IS the rectangle
 * rects[e] at height goal_z[e] (Rewarder._reset_goal, rewarder.py:252-258 -- what
 * srl_fill_goals_f32 writes) and is not read from memory. */
SRL_API int srl_pack_rewards_f32(const srl_env_state* host_state, const float* walls,
                                 const float* goals, const float* rocks, const float* goal_z,
                                 const int32_t* rects, void* wall_goal, void* rock,
                                 float* reward, double* value, int R, int H, int W, int h,
                                 int dtype_code, float obs_scale, int repeat_wall, int metric,
                                 double scale, double pixel_h, double pixel_w, double pmax,
                                 double pexp, double oexp, srl_stream_t stream);

/* srl_pack_rewards_f32 into PERSISTENT observation buffers: wall_goal still holds what
 * the previous call wrote for the same environments, and only the wall-image rows
 * rows[e] = (first, past-last) that changed since (srl_raster_incremental_rows) are
 * rewritten -- a placed rock touches ~18 of 64 rows, so a step writes a quarter of the
 * packed observation.  full [E] (uint8, may be NULL = always everything): non-zero
 * entries rewrite every row of that environment (new episode / goal, whole-scene redraw,
 * first use of the buffer) and are cleared by the call.  The rock images and the rewards
 * are always written in full; the result is the same bytes as srl_pack_rewards_f32. */
SRL_API int srl_pack_rewards_rows_f32(const srl_env_state* host_state, const float* walls,
                                      const float* goals, const float* rocks,
                                      const float* goal_z, const int32_t* rects,
                                      const int32_t* rows, uint8_t* full, void* wall_goal,
                                      void* rock, float* reward, double* value, int R, int H,
                                      int W, int h, int dtype_code, float obs_scale,
                                      int repeat_wall, int metric, double scale,
                                      double pixel_h, double pixel_w, double pmax, double pexp,
                                      double oexp, srl_stream_t stream);

/* StackEnv._return's uint8 cast (env.py:171-178) on the planar maps (the form the
 * scoring kernels take): q = trunc(x*255/scale) in float32. */
SRL_API int srl_quantise_planes_u8(const float* walls, const float* goals,
                                   const float* rocks, uint8_t* walls8, uint8_t* goals8,
                                   uint8_t* rocks8, int E, int R, int H, int W, int h,
                                   float scale, srl_stream_t stream);

/* Rock images by mesh id: out[e, :] = table[index[e], :] (rows of row_floats float32).
 * The image Observer.__call__ renders of the spawned rock (observer.py:262-293) depends
 * only on the mesh and the fixed spawn pose / orientation list, so a batched environment
 * rasterises every mesh of its bank once (srl_raster) and a step fetches the images of the
 * rocks spawned by srl_env_advance (state->current) with this call.  An index outside
 * [0, table_rows) fills the row with NaN. */
SRL_API int srl_gather_rows_f32(const float* table, const int32_t* index, float* out,
                                int rows_out, int row_floats, int table_rows,
                                srl_stream_t stream);

/* ---- measurement helpers (not part of the reference's surface) --------------
 * Issue-rate micro-benchmark used to fix the FP32 roofline of the max-plus
 * kernel: runs `iters` dependent rounds of (add, max) cells on every lane of a
 * full-chip grid and reports cells/s.  variant 0: FADD+FMNMX, 1: FADD+FMNMX3,
 * 2: FADD2+FMNMX3 (the kernel's mix).  Synchronises the device. */
SRL_API int srl_microbench_addmax(int variant, int iters, double* host_cells_per_s);
/* FP32 FMA issue rate (0 FFMA, 1 FFMA2 with a shared multiplicand like the
 * correlation kernel, 2 FFMA2 with distinct operands): FMAs per second. */
SRL_API int srl_microbench_fma(int variant, int iters, double* host_fma_per_s);

#ifdef __cplusplus
}
#endif
#endif  /* STACKRL_B200_H_ */
