// baselines.difference (reference: stackrl/baselines.py:45-77):
//
//   lift[u,v] = o[i+u,j+v] + n[u,v]                       (float32)
//   h0        = max( n > 0 ? lift : 0 )                    (float32, = height())
//   f[i,j]    = np.sum( w * |h0 - lift| ** p )             (float64)
//   w         = where(n > 0, ((u-h/2)^2 + (v-h/2)^2) ** (q/2), 0) / sum   (float64)
//
// To be bit-identical the kernel reproduces numpy's evaluation order: the
// residual and its square are float32 ops, the product with w is a float64
// multiply, and np.sum over the contiguous [h,h] block is numpy's PAIRWISE
// summation (blocks of <= 128 elements with 8 interleaved accumulators, combined
// as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), blocks joined by a binary tree that
// splits at n/2 rounded down to a multiple of 8).  The block/tree structure is
// compiled on the host into a tiny stack program (PwProg) every thread runs.
// Supported exponents: p in {1, 2}, q in {0, 2} (the reference defaults are 2, 2);
// other values go through libm pow in numpy and are not reproducible bitwise.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

constexpr int kMaxOps = 96;
struct PwProg {
  int n_ops;
  short start[kMaxOps];
  short len[kMaxOps];   // > 0: leaf [start, start+len); 0: add the two top entries
};

void build_prog(int start, int n, PwProg& prog) {
  if (n <= 128) {
    prog.start[prog.n_ops] = (short)start;
    prog.len[prog.n_ops++] = (short)n;
    return;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  build_prog(start, n2, prog);
  build_prog(start + n2, n - n2, prog);
  prog.start[prog.n_ops] = 0;
  prog.len[prog.n_ops++] = 0;
}

// numpy's pairwise sum of term(0..n-1) following `prog`; term(k) must be called
// with k increasing by one (callers keep running row/column counters).
template <typename F>
__device__ __forceinline__ double pairwise_sum(const PwProg& prog, F term) {
  double stack[8];
  int sp = 0;
  for (int op = 0; op < prog.n_ops; ++op) {
    const int len = prog.len[op];
    if (len == 0) {
      --sp;
      stack[sp - 1] = __dadd_rn(stack[sp - 1], stack[sp]);
      continue;
    }
    double res;
    if (len < 8) {
      res = 0.;
      for (int k = 0; k < len; ++k) res = __dadd_rn(res, term());
    } else {
      double r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = term();
      const int body = len - len % 8;
      for (int i = 8; i < body; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], term());
      }
      res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                      __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
      for (int i = body; i < len; ++i) res = __dadd_rn(res, term());
    }
    stack[sp++] = res;
  }
  return stack[0];
}

__device__ __forceinline__ float div_level(float x, float level) {
  return x == 0.f ? __fmul_rn(x, level) : __fdiv_rn(x, level);
}

// One thread per rock: radial weights, pairwise-summed and normalised.
__global__ void __launch_bounds__(128)
difference_weights_kernel(const float* __restrict__ rocks, const float* __restrict__ level,
                          double* __restrict__ weights, int nrocks, int R, int h,
                          int weights_exponent, const PwProg prog) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrocks) return;
  const float* rock = rocks + (size_t)idx * h * h;
  double* w = weights + (size_t)idx * h * h;
  const bool scaled = level != nullptr;
  const float g = scaled ? level[idx / R] : 1.f;
  const double half = (double)h / 2.;
  int u = 0, v = 0;
  auto term = [&]() {
    const float n = scaled ? div_level(rock[u * h + v], g) : rock[u * h + v];
    double x = 0.;
    if (n > 0.f) {
      if (weights_exponent > 0) {
        const double du = (double)u - half, dv = (double)v - half;
        x = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));
      } else {
        x = 1.;
      }
    }
    w[u * h + v] = x;
    if (++v == h) { v = 0; ++u; }
    return x;
  };
  const double total = pairwise_sum(prog, term);
  for (int k = 0; k < h * h; ++k) w[k] = __ddiv_rn(w[k], total);
}

__global__ void __launch_bounds__(128)
difference_kernel(const float* __restrict__ walls, const float* __restrict__ rocks,
                  const float* __restrict__ level, const double* __restrict__ weights,
                  double* __restrict__ out, float* __restrict__ top, int R, int H, int W,
                  int h, int band, int nbands, int difference_exponent,
                  const PwProg prog) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Ph = H - h + 1, Pw = W - h + 1;
  int b = blockIdx.x;
  const int bandi = b % nbands; b /= nbands;
  const int r = b % R;
  const int e = b / R;
  const int i0 = bandi * band;
  const int rows_out = min(band, Ph - i0);
  const int rows_in = rows_out + h - 1;

  double* w_s = reinterpret_cast<double*>(smem_raw);              // [h*h]
  float* rock_s = reinterpret_cast<float*>(w_s + h * h);          // [h*h] normalised
  float* wall_s = rock_s + h * h;                                 // [rows_in][W]
  __shared__ int s_masked;

  const int tid = threadIdx.x;
  const bool scaled = level != nullptr;
  const float g = scaled ? level[e] : 1.f;
  if (tid == 0) s_masked = 0;
  __syncthreads();
  const float* wall = walls + ((size_t)e * H + i0) * W;
  for (int k = tid; k < rows_in * W; k += blockDim.x)
    wall_s[k] = scaled ? div_level(wall[k], g) : wall[k];
  const float* rock = rocks + ((size_t)e * R + r) * h * h;
  const double* wgt = weights + ((size_t)e * R + r) * h * h;
  bool dead = false;
  for (int k = tid; k < h * h; k += blockDim.x) {
    const float n = scaled ? div_level(rock[k], g) : rock[k];
    rock_s[k] = n;
    w_s[k] = wgt[k];
    dead = dead || !(n > 0.f);
  }
  if (dead) s_masked = 1;
  __syncthreads();
  const bool floor0 = s_masked != 0;

  for (int item = tid; item < rows_out * Pw; item += blockDim.x) {
    const int i = item / Pw, j = item % Pw;
    const float* win = wall_s + i * W + j;
    // pass 1: h0, the max-plus value (baselines.py:68).
    float h0 = kNegInf;
    for (int u = 0; u < h; ++u)
      for (int v = 0; v < h; ++v) {
        const float n = rock_s[u * h + v];
        if (n > 0.f) h0 = fmaxf(h0, __fadd_rn(win[u * W + v], n));
      }
    if (floor0) h0 = fmaxf(h0, 0.f);
    // pass 2: weighted residual, numpy's pairwise order (baselines.py:69).
    int u = 0, v = 0;
    auto term = [&]() {
      const float lift = __fadd_rn(win[u * W + v], rock_s[u * h + v]);
      const float res = fabsf(__fsub_rn(h0, lift));
      const float pw = difference_exponent == 2 ? __fmul_rn(res, res) : res;
      const double x = __dmul_rn(w_s[u * h + v], (double)pw);
      if (++v == h) { v = 0; ++u; }
      return x;
    };
    const double f = pairwise_sum(prog, term);
    const size_t o = (((size_t)e * R + r) * Ph + i0 + i) * Pw + j;
    out[o] = f;
    if (top) top[o] = h0;
  }
}

}  // namespace

int difference_weights(const float* rocks, const float* level, double* weights, int E,
                       int R, int h, int weights_exponent, cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1, SRL_E_INVALID,
              "difference_weights: bad shape E=%d R=%d h=%d", E, R, h);
  SRL_REQUIRE(weights_exponent == 0 || weights_exponent == 2, SRL_E_UNSUPPORTED,
              "difference_weights: exponent %d (only 0 and 2 are bit-reproducible)",
              weights_exponent);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(rocks && weights, SRL_E_INVALID, "difference_weights: null pointer");
  SRL_REQUIRE(h <= 64, SRL_E_UNSUPPORTED, "difference_weights: rock side %d > 64", h);
  PwProg prog;
  prog.n_ops = 0;
  build_prog(0, h * h, prog);
  const int n = E * R;
  difference_weights_kernel<<<(n + 127) / 128, 128, 0, stream>>>(rocks, level, weights, n, R,
                                                                 h, weights_exponent, prog);
  return check_launch("difference_weights_kernel");
}

int difference_f32(const float* walls, const float* rocks, const float* level,
                   const double* weights, double* out, float* top, int E, int R, int H,
                   int W, int h, int difference_exponent, cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h, SRL_E_INVALID,
              "difference: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  SRL_REQUIRE(difference_exponent == 1 || difference_exponent == 2, SRL_E_UNSUPPORTED,
              "difference: exponent %d (only 1 and 2 are bit-reproducible)",
              difference_exponent);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && rocks && weights && out, SRL_E_INVALID, "difference: null pointer");
  SRL_REQUIRE(h <= 64, SRL_E_UNSUPPORTED, "difference: rock side %d > 64", h);
  const int Ph = H - h + 1;
  const int sms = sm_count();
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "difference: no CUDA device");
  auto smem_for = [&](int band) {
    return (size_t)12 * h * h + (size_t)4 * (band + h - 1) * W;
  };
  int band = Ph;
  while (band > 1 && ((size_t)E * R * ((Ph + band - 1) / band) < (size_t)2 * sms ||
                      smem_for(band) > 100 * 1024))
    band = (band + 1) / 2;
  SRL_REQUIRE(smem_for(band) <= 220 * 1024, SRL_E_UNSUPPORTED,
              "difference: %d-column wall rows with a %d-row rock exceed shared memory", W, h);
  const int nbands = (Ph + band - 1) / band;
  PwProg prog;
  prog.n_ops = 0;
  build_prog(0, h * h, prog);
  const size_t smem = smem_for(band);
  SRL_CUDA(cudaFuncSetAttribute(difference_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  difference_kernel<<<E * R * nbands, 128, smem, stream>>>(
      walls, rocks, level, weights, out, top, R, H, W, h, band, nbands,
      difference_exponent, prog);
  return check_launch("difference_kernel");
}

}  // namespace srl
