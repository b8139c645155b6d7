// Observation packing and heightmap rewards.
//
// pack_obs:   StackEnv.observation / _return (stackrl/envs/stack/env.py:171-180,
//             226-231) and TestStackEnv.observation (:472-480): interleave wall
//             and goal into [.., H, W, 2], add the channel axis to the rock map,
//             cast to the env dtype (uint8: x*255/scale in float32, truncated).
// reward_sums: Rewarder._intersection / _union (rewarder.py:297-307) and the
//             goal volume of :257, one warp-shuffle + shared-memory reduction
//             per environment.
#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

__device__ __forceinline__ uint8_t to_u8(float x, float scale) {
  // numpy: np.array(x*(2**8-1)/max(max_z, omd), dtype='uint8'), float32 ops,
  // C truncation toward zero.
  const float q = __fdiv_rn(__fmul_rn(x, 255.f), scale);
  return (uint8_t)(int)q;
}

template <bool U8>
__global__ void __launch_bounds__(256)
pack_obs_kernel(const float* __restrict__ walls, const float* __restrict__ goals,
                const float* __restrict__ rocks, void* __restrict__ wall_goal,
                void* __restrict__ rock, int E, int R, int H, int W, int h, float scale,
                int repeat_wall) {
  const size_t HW = (size_t)H * W;
  const size_t views = repeat_wall ? (size_t)R : 1;
  const size_t n_wall = (size_t)E * views * HW;          // (wall, goal) pairs to write
  const size_t n_rock = (size_t)E * R * h * h;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_wall; k += stride) {
    const size_t e = k / (views * HW), px = k % HW;
    const float w = walls[e * HW + px], g = goals[e * HW + px];
    if (U8) {
      reinterpret_cast<uchar2*>(wall_goal)[k] = make_uchar2(to_u8(w, scale), to_u8(g, scale));
    } else {
      reinterpret_cast<float2*>(wall_goal)[k] = make_float2(w, g);
    }
  }
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_rock; k += stride) {
    if (U8) reinterpret_cast<uint8_t*>(rock)[k] = to_u8(rocks[k], scale);
    else reinterpret_cast<float*>(rock)[k] = rocks[k];
  }
}

__global__ void __launch_bounds__(256)
reward_sums_kernel(const float* __restrict__ walls, const float* __restrict__ goals,
                   const float* __restrict__ goal_z, float* __restrict__ inter,
                   float* __restrict__ uni, float* __restrict__ vol, int HW) {
  __shared__ double s[3][8];
  const int e = blockIdx.x;
  const float* w = walls + (size_t)e * HW;
  const float* g = goals + (size_t)e * HW;
  const float gz = goal_z[e];
  double a = 0., b = 0., c = 0.;
  for (int k = threadIdx.x; k < HW; k += blockDim.x) {
    const float wv = w[k], gv = g[k];
    if (gv != 0.f) a += (double)fminf(wv, gz);      // rewarder.py:298-301
    b += (double)fmaxf(wv, gv);                     // rewarder.py:304-307
    c += (double)gv;                                // rewarder.py:257
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s[0][warp] = a;
    s[1][warp] = b;
    s[2][warp] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0., tb = 0., tc = 0.;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      ta += s[0][k];
      tb += s[1][k];
      tc += s[2][k];
    }
    inter[e] = (float)ta;
    uni[e] = (float)tb;
    if (vol) vol[e] = (float)tc;
  }
}

}  // namespace

int pack_obs(const float* walls, const float* goals, const float* rocks, void* wall_goal,
             void* rock, int E, int R, int H, int W, int h, int dtype_code, float scale,
             int repeat_wall, cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && H >= 1 && W >= 1 && h >= 1, SRL_E_INVALID,
              "pack_obs: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  SRL_REQUIRE(dtype_code == 0 || dtype_code == 1, SRL_E_UNSUPPORTED,
              "pack_obs: dtype code %d (0 float32, 1 uint8)", dtype_code);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && goals && rocks && wall_goal && rock, SRL_E_INVALID,
              "pack_obs: null pointer");
  SRL_REQUIRE(dtype_code == 0 || scale > 0.f, SRL_E_INVALID, "pack_obs: scale must be > 0");
  const int sms = sm_count();
  const int blocks = sms > 0 ? sms * 8 : 1184;
  if (dtype_code == 1)
    pack_obs_kernel<true><<<blocks, 256, 0, stream>>>(walls, goals, rocks, wall_goal, rock, E,
                                                      R, H, W, h, scale, repeat_wall);
  else
    pack_obs_kernel<false><<<blocks, 256, 0, stream>>>(walls, goals, rocks, wall_goal, rock, E,
                                                       R, H, W, h, scale, repeat_wall);
  return check_launch("pack_obs_kernel");
}

int reward_sums_f32(const float* walls, const float* goals, const float* goal_z, float* inter,
                    float* uni, float* vol, int E, int H, int W, cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && H >= 1 && W >= 1, SRL_E_INVALID, "reward_sums: bad shape");
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && goals && goal_z && inter && uni, SRL_E_INVALID,
              "reward_sums: null pointer");
  reward_sums_kernel<<<E, 256, 0, stream>>>(walls, goals, goal_z, inter, uni, vol, H * W);
  return check_launch("reward_sums_kernel");
}

}  // namespace srl
