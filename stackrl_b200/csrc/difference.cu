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
#include "pairwise.cuh"

namespace srl { constexpr int kNoQuantumHint = SRL_NO_QUANTUM; }

namespace srl {

namespace {

// One thread per rock: radial weights, pairwise-summed and normalised.
template <typename In>
__global__ void __launch_bounds__(128)
difference_weights_kernel(const In* __restrict__ rocks, const In* __restrict__ level,
                          double* __restrict__ weights, int nrocks, int R, int h,
                          int weights_exponent, const PwProg prog) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrocks) return;
  typedef typename Arith<In>::C C;
  const In* rock = rocks + (size_t)idx * h * h;
  double* w = weights + (size_t)idx * h * h;
  const bool scaled = level != nullptr;
  const In g = scaled ? level[idx / R] : In(1);
  const double half = (double)h / 2.;
  int u = 0, v = 0;
  auto term = [&]() {
    const C n = Arith<In>::norm(rock[u * h + v], g, scaled);
    double x = 0.;
    if (n > C(0)) {
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

template <typename In, bool HAVE_TOP>
__global__ void __launch_bounds__(128)
difference_kernel(const In* __restrict__ walls, const In* __restrict__ rocks,
                  const In* __restrict__ level, const double* __restrict__ weights,
                  double* __restrict__ out, typename Arith<In>::C* __restrict__ top, int R,
                  int H, int W,
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

  typedef Arith<In> A;
  typedef typename A::C C;
  double* w_s = reinterpret_cast<double*>(smem_raw);              // [h*h]
  C* rock_s = reinterpret_cast<C*>(w_s + h * h);                  // [h*h] normalised
  C* wall_s = rock_s + h * h;                                     // [rows_in][W]
  __shared__ int s_masked;

  const int tid = threadIdx.x;
  const bool scaled = level != nullptr;
  const In g = scaled ? level[e] : In(1);
  if (tid == 0) s_masked = 0;
  __syncthreads();
  const In* wall = walls + ((size_t)e * H + i0) * W;
  for (int k = tid; k < rows_in * W; k += blockDim.x) wall_s[k] = A::norm(wall[k], g, scaled);
  const In* rock = rocks + ((size_t)e * R + r) * h * h;
  const double* wgt = weights + ((size_t)e * R + r) * h * h;
  bool dead = false;
  for (int k = tid; k < h * h; k += blockDim.x) {
    const C n = A::norm(rock[k], g, scaled);
    rock_s[k] = n;
    w_s[k] = wgt[k];
    dead = dead || !(n > C(0));
  }
  if (dead) s_masked = 1;
  __syncthreads();
  const bool floor0 = s_masked != 0;

  for (int item = tid; item < rows_out * Pw; item += blockDim.x) {
    const int i = item / Pw, j = item % Pw;
    const C* win = wall_s + i * W + j;
    // pass 1: h0, the max-plus value (baselines.py:68) -- already in `top` when the
    // max-plus kernel produced it (same arithmetic, bit for bit).
    const size_t o = (((size_t)e * R + r) * Ph + i0 + i) * Pw + j;
    C h0;
    if constexpr (HAVE_TOP) {
      h0 = top[o];
    } else {
      h0 = C(kNegInf);
      for (int u = 0; u < h; ++u)
        for (int v = 0; v < h; ++v) {
          const C n = rock_s[u * h + v];
          if (n > C(0)) {
            const C lifted = A::add(win[u * W + v], n);
            h0 = lifted > h0 ? lifted : h0;
          }
        }
      if (floor0 && !(h0 > C(0))) h0 = C(0);
    }
    // pass 2: weighted residual, numpy's pairwise order (baselines.py:69).
    int u = 0, v = 0;
    auto term = [&]() {
      const C lift = A::add(win[u * W + v], rock_s[u * h + v]);
      C res = A::sub(h0, lift);
      res = res < C(0) ? -res : res;
      const C pw = difference_exponent == 2 ? A::mul(res, res) : res;
      const double x = __dmul_rn(w_s[u * h + v], (double)pw);
      if (++v == h) { v = 0; ++u; }
      return x;
    };
    const double f = pairwise_sum(prog, term);
    out[o] = f;
  }
}

}  // namespace

static int maxplus_into(const float* walls, const float* rocks, const float* level,
                        float* top, int E, int R, int H, int W, int h, cudaStream_t stream) {
  return maxplus_f32(walls, rocks, level, top, E, R, H, W, h, 0.f, 1, kNoQuantumHint, stream);
}
static int maxplus_into(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
                        double* top, int E, int R, int H, int W, int h, cudaStream_t stream) {
  return maxplus_u8(walls, rocks, level, top, E, R, H, W, h, stream);
}

template <typename In>
static int difference_weights_t(const In* rocks, const In* level, double* weights, int E,
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
  difference_weights_kernel<In><<<(n + 127) / 128, 128, 0, stream>>>(
      rocks, level, weights, n, R, h, weights_exponent, prog);
  return check_launch("difference_weights_kernel");
}

template <typename In>
static int difference_t(const In* walls, const In* rocks, const In* level,
                        const double* weights, double* out, typename Arith<In>::C* top,
                        int E, int R, int H, int W, int h, int difference_exponent,
                        cudaStream_t stream) {
  typedef typename Arith<In>::C C;
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
    return (size_t)(8 + sizeof(C)) * h * h + sizeof(C) * (size_t)(band + h - 1) * W;
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
  if (top != nullptr) {
    // h0 is the height() map: the max-plus kernels compute it much faster than a
    // per-thread scalar pass (and with the same bits).
    int rc = maxplus_into(walls, rocks, level, top, E, R, H, W, h, stream);
    if (rc != SRL_OK) return rc;
    SRL_CUDA(cudaFuncSetAttribute(difference_kernel<In, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    difference_kernel<In, true><<<E * R * nbands, 128, smem, stream>>>(
        walls, rocks, level, weights, out, top, R, H, W, h, band, nbands,
        difference_exponent, prog);
  } else {
    SRL_CUDA(cudaFuncSetAttribute(difference_kernel<In, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    difference_kernel<In, false><<<E * R * nbands, 128, smem, stream>>>(
        walls, rocks, level, weights, out, top, R, H, W, h, band, nbands,
        difference_exponent, prog);
  }
  return check_launch("difference_kernel");
}

int difference_weights(const float* rocks, const float* level, double* weights, int E,
                       int R, int h, int weights_exponent, cudaStream_t stream) {
  return difference_weights_t<float>(rocks, level, weights, E, R, h, weights_exponent, stream);
}
int difference_weights_u8(const uint8_t* rocks, double* weights, int E, int R, int h,
                          int weights_exponent, cudaStream_t stream) {
  // a / g > 0 <=> a > 0: the level is not needed for the uint8 weights
  return difference_weights_t<uint8_t>(rocks, nullptr, weights, E, R, h, weights_exponent,
                                       stream);
}
int difference_f32(const float* walls, const float* rocks, const float* level,
                   const double* weights, double* out, float* top, int E, int R, int H,
                   int W, int h, int difference_exponent, cudaStream_t stream) {
  return difference_t<float>(walls, rocks, level, weights, out, top, E, R, H, W, h,
                             difference_exponent, stream);
}
int difference_u8(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
                  const double* weights, double* out, double* top, int E, int R, int H,
                  int W, int h, int difference_exponent, cudaStream_t stream) {
  return difference_t<uint8_t>(walls, rocks, level, weights, out, top, E, R, H, W, h,
                               difference_exponent, stream);
}

}  // namespace srl
