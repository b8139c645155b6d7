// baselines.correlate and baselines.corrcoef (reference: stackrl/baselines.py:141-143,
// 79-85): VALID cross-correlation of the normalised wall with the normalised rock,
//   correlate = correlate2d(o, n) / n.sum()                       (scipy, float32)
//   corrcoef  = matchTemplate(o, n, TM_CCOEFF_NORMED)              (OpenCV, float32)
// Both reference results come out of third-party library code (scipy's direct
// sum, OpenCV's DFT / integral-image path) whose internal summation order is not
// part of the reference tree, so these two maps are matched to a TOLERANCE
// (1e-5 relative / 2e-5 absolute in the tests), not bit for bit.  corrcoef: every window
// sum is accumulated in float64 and rounded once (variance and covariance cancel).
// correlate alone: float32 FMAs over a rock row (scipy itself sums in float32), rows added
// in float64.
#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

__device__ __forceinline__ float div_level(float x, float level) {
  return x == 0.f ? __fmul_rn(x, level) : __fdiv_rn(x, level);
}

template <bool COEF>
__global__ void __launch_bounds__(128)
correlate_kernel(const float* __restrict__ walls, const float* __restrict__ rocks,
                 const float* __restrict__ level, float* __restrict__ corr,
                 float* __restrict__ coef, int R, int H, int W, int h, int band, int nbands) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_sum[4], s_sq[4];
  const int Ph = H - h + 1, Pw = W - h + 1;
  int b = blockIdx.x;
  const int bandi = b % nbands; b /= nbands;
  const int r = b % R;
  const int e = b / R;
  const int i0 = bandi * band;
  const int rows_out = min(band, Ph - i0);
  const int rows_in = rows_out + h - 1;
  float* rock_s = reinterpret_cast<float*>(smem_raw);     // [h*h]
  float* wall_s = rock_s + h * h;                         // [rows_in][Ws]
  const int Ws = W | 1;      // odd row stride: consecutive rows fall into consecutive banks
  const int tid = threadIdx.x;
  const bool scaled = level != nullptr;
  const float g = scaled ? level[e] : 1.f;
  const float* wall = walls + ((size_t)e * H + i0) * W;
  for (int k = tid; k < rows_in * W; k += blockDim.x) {
    const int row = k / W, col = k - row * W;
    wall_s[row * Ws + col] = scaled ? div_level(wall[k], g) : wall[k];
  }
  const float* rock = rocks + ((size_t)e * R + r) * h * h;
  double sn = 0., snn = 0.;
  for (int k = tid; k < h * h; k += blockDim.x) {
    const float n = scaled ? div_level(rock[k], g) : rock[k];
    rock_s[k] = n;
    sn += (double)n;
    snn += (double)n * (double)n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sn += __shfl_xor_sync(0xffffffffu, sn, o);
    snn += __shfl_xor_sync(0xffffffffu, snn, o);
  }
  if ((tid & 31) == 0) {
    s_sum[tid >> 5] = sn;
    s_sq[tid >> 5] = snn;
  }
  __syncthreads();
  sn = s_sum[0] + s_sum[1] + s_sum[2] + s_sum[3];
  snn = s_sq[0] + s_sq[1] + s_sq[2] + s_sq[3];
  const double N = (double)h * h;
  const double n_var = snn - sn * sn / N;

  if constexpr (!COEF) {
    // `correlate` alone: the window sums of o and o^2 are not needed, and scipy's own
    // correlate2d accumulates in float32.  Four adjacent outputs per thread share every rock
    // value; a rock row is accumulated with float32 FMAs (h terms), rows are added in float64.
    // (The last group of a row reads up to 3 floats past its window: inside wall_s, which is
    // allocated 4 floats longer; those outputs are not stored.)
    const int groups = (Pw + 3) / 4;
    for (int item = tid; item < rows_out * groups; item += blockDim.x) {
      // consecutive lanes = consecutive rows of one column group (conflict-free wall loads)
      const int grp = item / rows_out, i = item - grp * rows_out, j0 = grp * 4;
      double s0 = 0., s1 = 0., s2 = 0., s3 = 0.;
      for (int u = 0; u < h; ++u) {
        const float* wr = wall_s + (i + u) * Ws + j0;
        const float* rr = rock_s + u * h;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        float w0 = wr[0], w1 = wr[1], w2 = wr[2];
        int v = 0;
        for (; v + 4 <= h; v += 4) {          // register window slides without moves
          const float w3 = wr[v + 3], w4 = wr[v + 4], w5 = wr[v + 5], w6 = wr[v + 6];
          const float n0 = rr[v], n1 = rr[v + 1], n2 = rr[v + 2], n3 = rr[v + 3];
          a0 = fmaf(w0, n0, a0); a1 = fmaf(w1, n0, a1); a2 = fmaf(w2, n0, a2); a3 = fmaf(w3, n0, a3);
          a0 = fmaf(w1, n1, a0); a1 = fmaf(w2, n1, a1); a2 = fmaf(w3, n1, a2); a3 = fmaf(w4, n1, a3);
          a0 = fmaf(w2, n2, a0); a1 = fmaf(w3, n2, a1); a2 = fmaf(w4, n2, a2); a3 = fmaf(w5, n2, a3);
          a0 = fmaf(w3, n3, a0); a1 = fmaf(w4, n3, a1); a2 = fmaf(w5, n3, a2); a3 = fmaf(w6, n3, a3);
          w0 = w4; w1 = w5; w2 = w6;
        }
        for (; v < h; ++v) {
          const float w3 = wr[v + 3], n = rr[v];
          a0 = fmaf(w0, n, a0); a1 = fmaf(w1, n, a1); a2 = fmaf(w2, n, a2); a3 = fmaf(w3, n, a3);
          w0 = w1; w1 = w2; w2 = w3;
        }
        s0 += (double)a0; s1 += (double)a1; s2 += (double)a2; s3 += (double)a3;
      }
      const size_t at = (((size_t)e * R + r) * Ph + i0 + i) * Pw + j0;
      const float d = (float)sn;
      const double sv[4] = {s0, s1, s2, s3};
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (j0 + t < Pw) corr[at + t] = (float)sv[t] / d;
    }
    return;
  }
  for (int item = tid; item < rows_out * Pw; item += blockDim.x) {
    const int i = item / Pw, j = item % Pw;
    const float* win = wall_s + i * Ws + j;
    double son = 0., so = 0., soo = 0.;
    for (int u = 0; u < h; ++u)
      for (int v = 0; v < h; ++v) {
        const double o = (double)win[u * Ws + v];
        son += o * (double)rock_s[u * h + v];
        so += o;
        soo += o * o;
      }
    const size_t at = (((size_t)e * R + r) * Ph + i0 + i) * Pw + j;
    if (corr) corr[at] = (float)son / (float)sn;
    if (coef) {
      // OpenCV's degenerate-window rule (templmatch.cpp): |num| < den -> num/den,
      // |num| < 1.125 den -> +-1, else 0.
      const double num = son - so * sn / N;
      const double o_var = soo - so * so / N;
      const double den = sqrt(fmax(n_var, 0.) * fmax(o_var, 0.));
      double c;
      if (fabs(num) < den) c = num / den;
      else if (fabs(num) < den * 1.125) c = num > 0 ? 1. : -1.;
      else c = 0.;
      coef[at] = (float)c;
    }
  }
}

}  // namespace

int correlate_f32(const float* walls, const float* rocks, const float* level, float* corr,
                  float* coef, int E, int R, int H, int W, int h, cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h, SRL_E_INVALID,
              "correlate: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && rocks && (corr || coef), SRL_E_INVALID, "correlate: null pointer");
  const int Ph = H - h + 1;
  const int sms = sm_count();
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "correlate: no CUDA device");
  auto smem_for = [&](int band) {
    return (size_t)4 * h * h + (size_t)4 * (band + h - 1) * (W | 1) + 16;      // + the 4-float overrun
  };
  int band = Ph;
  while (band > 1 && ((size_t)E * R * ((Ph + band - 1) / band) < (size_t)2 * sms ||
                      smem_for(band) > 100 * 1024))
    band = (band + 1) / 2;
  SRL_REQUIRE(smem_for(band) <= 220 * 1024, SRL_E_UNSUPPORTED,
              "correlate: %d-column wall rows with a %d-row rock exceed shared memory", W, h);
  const int nbands = (Ph + band - 1) / band;
  const size_t smem = smem_for(band);
  if (coef) {
    SRL_CUDA(cudaFuncSetAttribute(correlate_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    correlate_kernel<true><<<E * R * nbands, 128, smem, stream>>>(walls, rocks, level, corr, coef,
                                                                 R, H, W, h, band, nbands);
  } else {
    SRL_CUDA(cudaFuncSetAttribute(correlate_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    correlate_kernel<false><<<E * R * nbands, 128, smem, stream>>>(walls, rocks, level, corr,
                                                                  coef, R, H, W, h, band, nbands);
  }
  return check_launch("correlate_kernel");
}

}  // namespace srl
