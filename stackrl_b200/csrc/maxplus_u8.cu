// Max-plus map for uint8 observations (registered Stack-v0/1/2 dtype).
//
// Reference: stackrl/baselines.py:21-43 on uint8 arrays: get_inputs divides
// uint8 by uint8, which numpy evaluates in float64, so the reference computes
//     f[i,j] = max_{u,v}( b > 0 ? fl64(a/g) + fl64(b/g) : 0 )
// with a = wall[i+u,j+v], b = rock[u,v], g = goal.max().  fl(a/g)+fl(b/g) is
// NOT fl((a+b)/g) bitwise, so the kernel evaluates exactly that:
// a 256-entry table of IEEE float64 quotients x/g per environment, one DADD per
// cell, max over cells.  Output float64, bit-exact.
//
// One CTA per (environment, rotation, band of output rows) so that a single
// observation (the drop-in `height(obs)` call) still spreads over the chip.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

constexpr int kT = 4;          // outputs per thread along a row
constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads)
maxplus_u8_kernel(const uint8_t* __restrict__ walls, const uint8_t* __restrict__ rocks,
                  const uint8_t* __restrict__ level, double* __restrict__ out, int R,
                  int H, int W, int h, int band, int nbands) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Ph = H - h + 1, Pw = W - h + 1;
  int b = blockIdx.x;
  const int bandi = b % nbands; b /= nbands;
  const int r = b % R;
  const int e = b / R;
  const int i0 = bandi * band;
  const int rows_out = min(band, Ph - i0);
  const int rows_in = rows_out + h - 1;
  const int Wp = W + kT;                           // row stride (zero padded)

  double* lut = reinterpret_cast<double*>(smem_raw);         // [256]
  double* wall_d = lut + 256;                                // [rows_in][Wp]
  double* rock_d = wall_d + (size_t)(band + h - 1) * Wp;     // [h][h], -inf = masked
  __shared__ int s_masked;

  const int tid = threadIdx.x;
  const double g = (double)level[e];
  for (int k = tid; k < 256; k += kThreads) lut[k] = __ddiv_rn((double)k, g);
  if (tid == 0) s_masked = 0;
  __syncthreads();
  const uint8_t* wall = walls + ((size_t)e * H + i0) * W;
  for (int k = tid; k < rows_in * Wp; k += kThreads) {
    const int row = k / Wp, c = k % Wp;
    wall_d[k] = c < W ? lut[wall[row * W + c]] : 0.;
  }
  const uint8_t* rock = rocks + ((size_t)e * R + r) * h * h;
  bool dead = false;
  for (int k = tid; k < h * h; k += kThreads) {
    const uint8_t x = rock[k];
    // n > 0 with n = x/g  <=>  x > 0 (baselines.py:32)
    rock_d[k] = x > 0 ? lut[x] : -CUDART_INF;
    dead = dead || x == 0;
  }
  if (dead) s_masked = 1;
  __syncthreads();
  const bool floor0 = s_masked != 0;

  const int strips = (Pw + kT - 1) / kT;
  double* o = out + (((size_t)e * R + r) * Ph + i0) * Pw;
  for (int item = tid; item < rows_out * strips; item += kThreads) {
    const int i = item / strips, j0 = (item % strips) * kT;
    double acc[kT];
#pragma unroll
    for (int t = 0; t < kT; ++t) acc[t] = -CUDART_INF;
    for (int u = 0; u < h; ++u) {
      const double* wrow = wall_d + (i + u) * Wp + j0;
      const double* nrow = rock_d + u * h;
      double win[kT];
#pragma unroll
      for (int t = 0; t < kT - 1; ++t) win[t] = wrow[t];
      for (int v = 0; v < h; ++v) {
        win[kT - 1] = wrow[v + kT - 1];
        const double n = nrow[v];
#pragma unroll
        for (int t = 0; t < kT; ++t) acc[t] = fmax(acc[t], __dadd_rn(win[t], n));
#pragma unroll
        for (int t = 0; t < kT - 1; ++t) win[t] = win[t + 1];
      }
    }
#pragma unroll
    for (int t = 0; t < kT; ++t)
      if (j0 + t < Pw) o[i * Pw + j0 + t] = floor0 ? fmax(acc[t], 0.) : acc[t];
  }
}

}  // namespace

int maxplus_u8(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
               double* out, int E, int R, int H, int W, int h, cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h, SRL_E_INVALID,
              "maxplus_u8: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && rocks && level && out, SRL_E_INVALID, "maxplus_u8: null pointer");
  const int Ph = H - h + 1;
  const int sms = sm_count();
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "maxplus_u8: no CUDA device");
  // Bands: enough CTAs to fill the chip twice, and a band must fit shared memory.
  auto smem_for = [&](int band) {
    return (size_t)8 * (256 + (size_t)(band + h - 1) * (W + kT) + (size_t)h * h);
  };
  int band = Ph;
  while (band > 1 && ((size_t)E * R * ((Ph + band - 1) / band) < (size_t)2 * sms ||
                      smem_for(band) > 100 * 1024))
    band = (band + 1) / 2;
  SRL_REQUIRE(smem_for(band) <= 220 * 1024, SRL_E_UNSUPPORTED,
              "maxplus_u8: wall rows of %d columns with a %d-row rock exceed shared memory",
              W, h);
  const int nbands = (Ph + band - 1) / band;
  const size_t smem = smem_for(band);
  SRL_CUDA(cudaFuncSetAttribute(maxplus_u8_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  maxplus_u8_kernel<<<E * R * nbands, kThreads, smem, stream>>>(walls, rocks, level, out, R,
                                                             H, W, h, band, nbands);
  return check_launch("maxplus_u8_kernel");
}

}  // namespace srl
