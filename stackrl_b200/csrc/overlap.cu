// goal_overlap counts (reference: stackrl/baselines.py:152-155):
//
//   counts[e,r,i,j] = sum_{u,v} (wall[e,i+u,j+v] < goal[e,i+u,j+v]) * (rock[e,r,u,v] > 0)
//
// scipy.signal.correlate2d on two 0/1 integer images is an integer sliding sum;
// here both images are bit-packed (one warp ballot per 32 pixels), the h-bit
// window of every (wall row, column offset) is extracted once with a funnel
// shift and shared by all R rotations, and a count is h AND+POPC steps.
// Exact int32; the `>= threshold * max` compare lives in srl_select.
#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

template <typename In>
__global__ void __launch_bounds__(256)
goal_overlap_kernel(const In* __restrict__ walls, const In* __restrict__ goals,
                    const In* __restrict__ rocks, int32_t* __restrict__ counts, int R,
                    int H, int W, int h) {
  extern __shared__ uint32_t smem_u[];
  const int e = blockIdx.x;
  const int Ph = H - h + 1, Pw = W - h + 1;
  const int nW = (W + 31) / 32 + 1;              // +1 zero word for the funnel shift
  uint32_t* below = smem_u;                      // [H][nW]
  uint32_t* foot = below + H * nW;               // [R][h]
  uint32_t* win = foot + R * h;                  // [H][Pw]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;

  // Stage the two comparisons as bytes first: every global load of the block is
  // independent and coalesced (several in flight per thread), the bit-packing
  // below then only touches shared memory.
  uint8_t* flag = reinterpret_cast<uint8_t*>(win + H * Pw);      // [H*W + R*h*h]
  const In* wall = walls + (size_t)e * H * W;
  const In* goal = goals + (size_t)e * H * W;
  const In* rock = rocks + (size_t)e * R * h * h;
#pragma unroll 4
  for (int k = tid; k < H * W; k += blockDim.x) flag[k] = wall[k] < goal[k];
  uint8_t* rflag = flag + H * W;
#pragma unroll 4
  for (int k = tid; k < R * h * h; k += blockDim.x) rflag[k] = rock[k] > In(0);
  __syncthreads();
  for (int k = warp; k < H * nW; k += nwarps) {
    const int row = k / nW, col = (k % nW) * 32 + lane;
    const uint32_t bits = __ballot_sync(0xffffffffu, col < W && flag[row * W + col]);
    if (lane == 0) below[k] = bits;
  }
  for (int k = warp; k < R * h; k += nwarps) {
    const uint32_t bits = __ballot_sync(0xffffffffu, lane < h && rflag[k * h + lane]);
    if (lane == 0) foot[k] = bits;
  }
  __syncthreads();
  for (int k = tid; k < H * Pw; k += blockDim.x) {
    const int row = k / Pw, j = k % Pw;
    const uint32_t lo = below[row * nW + (j >> 5)], hi = below[row * nW + (j >> 5) + 1];
    win[k] = __funnelshift_r(lo, hi, j & 31);    // bits above h are cut by `foot`
  }
  __syncthreads();
  int32_t* out = counts + (size_t)e * R * Ph * Pw;
  for (int k = tid; k < R * Ph * Pw; k += blockDim.x) {
    const int j = k % Pw;
    const int ri = k / Pw;
    const int i = ri % Ph, r = ri / Ph;
    int c = 0;
    for (int u = 0; u < h; ++u) c += __popc(win[(i + u) * Pw + j] & foot[r * h + u]);
    out[k] = c;
  }
}

template <typename In>
int launch(const In* walls, const In* goals, const In* rocks, int32_t* counts, int E,
           int R, int H, int W, int h, cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h, SRL_E_INVALID,
              "goal_overlap: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && goals && rocks && counts, SRL_E_INVALID,
              "goal_overlap: null pointer");
  SRL_REQUIRE(h <= 32, SRL_E_UNSUPPORTED,
              "goal_overlap: rock side %d > 32 (bit-packed rows hold 32 pixels)", h);
  const int nW = (W + 31) / 32 + 1;
  const size_t smem = 4 * ((size_t)H * nW + (size_t)R * h + (size_t)H * (W - h + 1)) +
                      (size_t)H * W + (size_t)R * h * h;
  SRL_REQUIRE(smem <= 220 * 1024, SRL_E_UNSUPPORTED,
              "goal_overlap: %dx%d wall with %d rotations exceeds shared memory", H, W, R);
  auto k = goal_overlap_kernel<In>;
  SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<E, 256, smem, stream>>>(walls, goals, rocks, counts, R, H, W, h);
  return check_launch("goal_overlap_kernel");
}

}  // namespace

int goal_overlap_f32(const float* walls, const float* goals, const float* rocks,
                     int32_t* counts, int E, int R, int H, int W, int h,
                     cudaStream_t stream) {
  return launch<float>(walls, goals, rocks, counts, E, R, H, W, h, stream);
}

int goal_overlap_u8(const uint8_t* walls, const uint8_t* goals, const uint8_t* rocks,
                    int32_t* counts, int E, int R, int H, int W, int h,
                    cudaStream_t stream) {
  return launch<uint8_t>(walls, goals, rocks, counts, E, R, H, W, h, stream);
}

}  // namespace srl
