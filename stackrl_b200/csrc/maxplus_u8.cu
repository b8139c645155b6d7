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
#include <stdlib.h>

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

// --------------------------------------------------------------------------- //
// Integer-key kernel (default).  The float64 value only matters for ONE cell per
// output, the maximum; finding it does not need float64 arithmetic:
//   * T[x] = fl64(x / g) = x/g + d_x with g * d_x * 2^60 =: n_x an INTEGER,
//     |n_x| < 2^15 (T[x] is a multiple of 2^-60 and |d_x| <= ulp(T[x]) / 2).
//   * The exact real value of T[a] + T[b] is ((a + b) + (n_a + n_b) 2^-60) / g:
//     cells are ordered by a + b first (a step of 1/g dwarfs the rounding terms),
//     by n_a + n_b second, and fl64 is monotone, so
//         max_cells fl64(T[a] + T[b]) = fl64 of the cell with the largest integer key
//         key = (a + b) 2^22 + (n_a + n_b + 2^16) 2^5 + (a >> 3)          (< 2^31)
//     which is the SUM of a per-pixel wall key and a per-pixel rock key.
//   * The sweep is therefore one VIADDMNMX.S32 (acc = max(acc, wall + rock)) per
//     cell; the a >> 3 bits name 8 candidate wall values, among which the
//     epilogue finds the pair (a, S - a) with the winning n_a + n_b and returns
//     T[a] + T[S - a] (one float64 add: the reference's own operation).
// Bit-exact with the float64 kernel above by construction (tests compare both).
// --------------------------------------------------------------------------- //
// One LDS.128 the compiler may not re-issue (it otherwise trades the register
// tile for one shared-memory load per cell).
__device__ __forceinline__ int4 lds128_once(const int* p) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(smem_u32(p)));
  return v;
}

constexpr int kKT = 8;            // outputs per thread along a row
constexpr int kKeyThreads = 128;
constexpr int kMaskedKey = -(1 << 30);

template <int HH>                 // rock side at compile time (0: run time)
__global__ void __launch_bounds__(kKeyThreads)
maxplus_u8_key_kernel(const uint8_t* __restrict__ walls, const uint8_t* __restrict__ rocks,
                      const uint8_t* __restrict__ level, double* __restrict__ out, int R,
                      int H, int W, int h_rt, int band, int nbands) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int h = HH > 0 ? HH : h_rt;
  const int Ph = H - h + 1, Pw = W - h + 1;
  int b = blockIdx.x;
  const int bandi = b % nbands; b /= nbands;
  const int r = b % R;
  const int e = b / R;
  const int i0 = bandi * band;
  const int rows_out = min(band, Ph - i0);
  const int rows_in = rows_out + h - 1;
  const int Wp = round_up(W + kKT, 4);                          // row stride, zero keys behind W

  double* lut = reinterpret_cast<double*>(smem_raw);            // [256] T[x]
  int* nx = reinterpret_cast<int*>(lut + 256);                  // [256] n_x
  int* wkey = nx + 256;                                         // [rows_in][Wp]
  int* nkey = wkey + (size_t)(band + h - 1) * Wp;               // [h][h]
  __shared__ int s_masked;

  const int tid = threadIdx.x;
  const double g = (double)level[e];
  for (int k = tid; k < 256; k += kKeyThreads) {
    const double t = __ddiv_rn((double)k, g);
    // g * t - k is a multiple of 2^-60 below 2^-38: the fma result is exact
    const double rho = __fma_rn(g, t, -(double)k);
    lut[k] = t;
    nx[k] = __double2int_rn(rho * 1152921504606846976.0);        // * 2^60, exact
  }
  if (tid == 0) s_masked = 0;
  __syncthreads();
  const uint8_t* wall = walls + ((size_t)e * H + i0) * W;
  for (int k = tid; k < rows_in * Wp; k += kKeyThreads) {
    const int row = k / Wp, c = k - row * Wp;
    int key = 0;
    if (c < W) {
      const int a = wall[row * W + c];
      key = (a << 22) + ((nx[a] + 32768) << 5) + (a >> 3);
    }
    wkey[k] = key;
  }
  const uint8_t* rock = rocks + ((size_t)e * R + r) * h * h;
  bool dead = false;
  for (int k = tid; k < h * h; k += kKeyThreads) {
    const int x = rock[k];
    // n > 0 with n = x/g  <=>  x > 0 (baselines.py:32)
    nkey[k] = x > 0 ? (x << 22) + ((nx[x] + 32768) << 5) : kMaskedKey;
    dead = dead || x == 0;
  }
  if (dead) s_masked = 1;
  __syncthreads();
  const bool floor0 = s_masked != 0;

  const int strips = (Pw + kKT - 1) / kKT;
  double* o = out + (((size_t)e * R + r) * Ph + i0) * Pw;
  for (int item = tid; item < rows_out * strips; item += kKeyThreads) {
    const int i = item / strips, j0 = (item - i * strips) * kKT;
    int acc[kKT];
#pragma unroll
    for (int t = 0; t < kKT; ++t) acc[t] = 2 * kMaskedKey;
    for (int u = 0; u < h; ++u) {
      const int* wrow = wkey + (i + u) * Wp + j0;
      const int* nrow = nkey + u * h;
      if constexpr (HH > 0) {
        // whole row segment in registers, fully unrolled
        int w[HH + kKT], n[HH];
#pragma unroll
        for (int k = 0; k < (HH + kKT) / 4; ++k) {
          const int4 x = lds128_once(wrow + 4 * k);
          w[4 * k] = x.x; w[4 * k + 1] = x.y; w[4 * k + 2] = x.z; w[4 * k + 3] = x.w;
        }
#pragma unroll
        for (int k = 0; k < HH / 4; ++k) {
          const int4 x = lds128_once(nrow + 4 * k);
          n[4 * k] = x.x; n[4 * k + 1] = x.y; n[4 * k + 2] = x.z; n[4 * k + 3] = x.w;
        }
#pragma unroll
        for (int v = 0; v < HH; ++v) {
#pragma unroll
          for (int t = 0; t < kKT; ++t) acc[t] = __viaddmax_s32(w[t + v], n[v], acc[t]);
        }
      } else {
        int win[kKT];
#pragma unroll
        for (int t = 0; t < kKT - 1; ++t) win[t] = wrow[t];
        for (int v = 0; v < h; ++v) {
          win[kKT - 1] = wrow[v + kKT - 1];
          const int n = nrow[v];
#pragma unroll
          for (int t = 0; t < kKT; ++t) acc[t] = __viaddmax_s32(win[t], n, acc[t]);
#pragma unroll
          for (int t = 0; t < kKT - 1; ++t) win[t] = win[t + 1];
        }
      }
    }
#pragma unroll
    for (int t = 0; t < kKT; ++t) {
      if (j0 + t >= Pw) continue;
      double val = 0.;                      // no live cell: np.where(...) is all zeros
      const int key = acc[t];
      if (key >= 0) {
        const int S = key >> 22, D = ((key >> 5) & 0x1ffff) - 65536, a0 = (key & 31) << 3;
        val = -CUDART_INF;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int a = a0 + c, bb = S - a;
          if (bb >= 1 && bb <= 255 && nx[a] + nx[bb] == D) val = __dadd_rn(lut[a], lut[bb]);
        }
        if (floor0) val = fmax(val, 0.);
      }
      o[i * Pw + j0 + t] = val;
    }
  }
}

// Register-tiled variant for the reference rock sizes: one CTA per (environment,
// rotation chunk, band of output rows) so that the quotient table and the wall
// keys are built once for RC rotations; each thread owns KT = 4m+1 consecutive
// outputs of a row (strips pitched KT-1 apart, like maxplus_f32) and keeps the
// HH + KT - 1 wall keys and HH rock keys of one rock row in registers.
template <int HH, int KT>
__global__ void __launch_bounds__(256)
maxplus_u8_tile_kernel(const uint8_t* __restrict__ walls, const uint8_t* __restrict__ rocks,
                       const uint8_t* __restrict__ level, double* __restrict__ out, int R,
                       int H, int W, int band, int nbands, int RC, int rchunks, int Wp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int h = HH, S = KT - 1;
  constexpr int NW4 = (HH + KT - 1 + 3) / 4;                   // LDS.128 per wall row segment
  const int Ph = H - h + 1, Pw = W - h + 1;
  int b = blockIdx.x;
  const int bandi = b % nbands; b /= nbands;
  const int rc = b % rchunks;
  const int e = b / rchunks;
  const int r0 = rc * RC, RCv = min(RC, R - r0);
  const int i0 = bandi * band;
  const int rows_out = min(band, Ph - i0);
  const int rows_in = rows_out + h - 1;

  // Closed-form epilogue.  The reference value of the winning cell is fl64(T[a] + T[b]),
  // and the exact sum is (S 2^60 + D) / (g 2^60) with S = a + b, D = n_a + n_b: it depends
  // on the key's (S, D) alone, not on which pair produced it.  With t1 = fl64(S / g) and
  // its exact remainder r_S = S - g t1 (|r_S| 2^60 < 2^24, an integer),
  //   T[a] + T[b] = t1 + ((r_S 2^60 + D) / g) 2^-60
  // and one rounding of that sum is the reference's one rounding: the quotient is needed
  // to 1 ulp only (a non-tie sum is at least ulp(t1) / 510 away from a rounding boundary)
  // except when it is exactly representable (ties: the quotient is then a power of two),
  // which a reciprocal product with one fma correction reproduces exactly.  Checked for
  // every level, wall and rock value (16.6 M pairs) in numpy; n_x = -r_x 2^60 for x <= 255.
  double* lut = reinterpret_cast<double*>(smem_raw);            // [512] t1(S) = fl64(S / g)
  int* nx = reinterpret_cast<int*>(lut + 512);                  // [512] -r_S 2^60 (n_x for x < 256)
  int* wkey = nx + 512;                                         // [band+h-1][Wp]
  int* nkey = wkey + (size_t)(band + h - 1) * Wp;               // [RC][h][h]
  int* dead_s = nkey + (size_t)RC * h * h;                      // [RC]

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const double g = (double)level[e];
  const double rg = __drcp_rn(g);
  for (int k = tid; k < 511; k += nthreads) {
    const double t = __ddiv_rn((double)k, g);
    const double rho = __fma_rn(g, t, -(double)k);              // exact (see above)
    lut[k] = t;
    nx[k] = __double2int_rn(rho * 1152921504606846976.0);       // * 2^60, exact
  }
  for (int k = tid; k < RC; k += nthreads) dead_s[k] = 0;
  __syncthreads();
  const uint8_t* wall = walls + ((size_t)e * H + i0) * W;
  for (int k = tid; k < rows_in * Wp; k += nthreads) {
    const int row = k / Wp, c = k - row * Wp;
    int key = 0;
    if (c < W) {
      const int a = wall[row * W + c];
      key = (a << 22) + ((nx[a] + 32768) << 5);
    }
    wkey[k] = key;
  }
  const uint8_t* rock = rocks + ((size_t)e * R + r0) * h * h;
  for (int k = tid; k < RCv * h * h; k += nthreads) {
    const int x = rock[k];
    nkey[k] = x > 0 ? (x << 22) + ((nx[x] + 32768) << 5) : kMaskedKey;
    if (x == 0) dead_s[k / (h * h)] = 1;
  }
  __syncthreads();

  const int strips = Pw <= KT ? 1 : (Pw - KT + S - 1) / S + 1;
  const int per_rot = rows_out * strips;
  for (int item = tid; item < RCv * per_rot; item += nthreads) {
    // consecutive lanes = consecutive output rows of one strip: with Wp / 4 odd the eight
    // lanes of a quarter-warp read eight different 4-bank groups (no LDS.128 conflicts)
    const int rr = item / per_rot, rem = item - rr * per_rot;
    const int strip = rem / rows_out, i = rem - strip * rows_out;
    const int j0 = strip * S;
    const int* nbase = nkey + rr * h * h;
    int acc[KT];
#pragma unroll
    for (int t = 0; t < KT; ++t) acc[t] = 2 * kMaskedKey;
#pragma unroll 2
    for (int u = 0; u < h; ++u) {
      const int* wrow = wkey + (i + u) * Wp + j0;
      const int* nrow = nbase + u * h;
      int w[4 * NW4], n[HH];
#pragma unroll
      for (int k = 0; k < NW4; ++k) {
        const int4 x = lds128_once(wrow + 4 * k);
        w[4 * k] = x.x; w[4 * k + 1] = x.y; w[4 * k + 2] = x.z; w[4 * k + 3] = x.w;
      }
#pragma unroll
      for (int k = 0; k < HH / 4; ++k) {
        const int4 x = lds128_once(nrow + 4 * k);
        n[4 * k] = x.x; n[4 * k + 1] = x.y; n[4 * k + 2] = x.z; n[4 * k + 3] = x.w;
      }
#pragma unroll
      for (int v = 0; v < HH; ++v) {
#pragma unroll
        for (int t = 0; t < KT; ++t) acc[t] = __viaddmax_s32(w[t + v], n[v], acc[t]);
      }
    }
    const bool floor0 = dead_s[rr] != 0;
    const int ncols = strip == strips - 1 ? min(KT, Pw - j0) : S;
    double* o = out + ((((size_t)e * R + r0 + rr) * Ph + i0 + i) * Pw + j0);
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      if (t >= ncols) continue;
      double val = 0.;                      // no live cell: np.where(...) is all zeros
      const int key = acc[t];
      if (key >= 0) {
        const int Ssum = key >> 22, D = ((key >> 5) & 0x1ffff) - 65536;
        const double pnum = (double)(D - nx[Ssum]);             // r_S 2^60 + D, exact
        const double q = __dmul_rn(pnum, rg);
        const double q2 = __fma_rn(__fma_rn(-q, g, pnum), rg, q);
        val = __fma_rn(q2, 8.67361737988403547e-19, lut[Ssum]); // * 2^-60 (exact), one rounding
        if (floor0) val = fmax(val, 0.);
      }
      o[t] = val;
    }
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
  // ---- integer-key kernel (default; SRL_U8_MODE=0 selects the float64 one) -------- //
  bool use_keys = true;
  if (const char* m = getenv("SRL_U8_MODE")) use_keys = atoi(m) != 0;
  if (use_keys && (h == 8 || h == 16 || h == 32)) {
    // Register-tiled kernel: KT in {9, 17, 25} by lane waste, rotations chunked to
    // fit shared memory, rows banded until the chip is full.
    const int Pw = W - h + 1;
    int KT = 9;
    double best = 1e30;
    for (int cand : {9, 17, 25}) {
      const int st = Pw <= cand ? 1 : (Pw - cand + cand - 2) / (cand - 1) + 1;
      // lane waste, biased towards narrow tiles: fewer registers per thread means
      // more resident warps, which is what hides the LDS latency of this kernel
      // (measured: 9 beats 17 by 20 % at 32/16, ties with it at 128/32)
      const double cost = (double)st * cand / Pw + 0.01 * cand;
      if (cost < best - 1e-12) { best = cost; KT = cand; }
    }
    // latency-bound calls (one observation): more, shorter threads
    if ((long long)E * R * Ph * ((Pw + 7) / 8) < (long long)sms * 256) KT = 9;
    if (const char* kt = getenv("SRL_U8_KT")) {   // tuning override
      const int v = atoi(kt);
      if (v == 9 || v == 17 || v == 25) KT = v;
    }
    const int strips = Pw <= KT ? 1 : (Pw - KT + KT - 2) / (KT - 1) + 1;
    const int nw4 = (h + KT - 1 + 3) / 4;
    int Wp = (strips - 1) * (KT - 1) + 4 * nw4;
    if (Wp < W) Wp = W;
    Wp = round_up(Wp, 4);
    if ((Wp / 4) % 2 == 0) Wp += 4;       // odd number of 16-byte groups per row (see the kernel)
    auto tsmem = [&](int bd, int rc) {
      return (size_t)512 * 12 + 4 * ((size_t)(bd + h - 1) * Wp + (size_t)rc * h * h + rc);
    };
    int RC = R, tband = Ph;
    while (RC > 1 && tsmem(1, RC) > 96 * 1024) RC = (RC + 1) / 2;
    while (tband > 1 && ((size_t)E * ((R + RC - 1) / RC) * ((Ph + tband - 1) / tband) <
                             (size_t)2 * sms ||
                         tsmem(tband, RC) > 100 * 1024))
      tband = (tband + 1) / 2;
    // Few CTAs left although every band has one row: split the rotations instead.
    while (RC > 1 && (size_t)E * ((R + RC - 1) / RC) * ((Ph + tband - 1) / tband) <
                         (size_t)2 * sms)
      RC = (RC + 1) / 2;
    if (tsmem(tband, RC) <= 200 * 1024) {
      const int tnb = (Ph + tband - 1) / tband, rchunks = (R + RC - 1) / RC;
      const size_t sm = tsmem(tband, RC);
      // block size: whole warps, as few idle lanes over the passes as possible
      const int items = RC * tband * strips;
      int threads = 256;
      double bw = 1e30;
      for (int t = 64; t <= 256; t += 32) {
        const int passes = (items + t - 1) / t;
        const double wst = (double)passes * t / items + 0.02 * passes;
        if (wst < bw - 1e-9) { bw = wst; threads = t; }
      }
#define SRL_U8T(HH, KK)                                                                    \
  if (h == HH && KT == KK) {                                                               \
    SRL_CUDA(cudaFuncSetAttribute(maxplus_u8_tile_kernel<HH, KK>,                          \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));  \
    maxplus_u8_tile_kernel<HH, KK><<<E * rchunks * tnb, threads, sm, stream>>>(        \
        walls, rocks, level, out, R, H, W, tband, tnb, RC, rchunks, Wp);                   \
    return check_launch("maxplus_u8_tile_kernel");                                         \
  }
      SRL_U8T(8, 9) SRL_U8T(8, 17) SRL_U8T(8, 25)
      SRL_U8T(16, 9) SRL_U8T(16, 17) SRL_U8T(16, 25)
      SRL_U8T(32, 9) SRL_U8T(32, 17) SRL_U8T(32, 25)
#undef SRL_U8T
    }
  }
  if (use_keys) {
    const int Wp = round_up(W + kKT, 4);
    auto ksmem = [&](int bd) {
      return (size_t)256 * 12 + 4 * ((size_t)(bd + h - 1) * Wp + (size_t)h * h);
    };
    int kband = Ph;
    while (kband > 1 && ((size_t)E * R * ((Ph + kband - 1) / kband) < (size_t)2 * sms ||
                         ksmem(kband) > 100 * 1024))
      kband = (kband + 1) / 2;
    if (ksmem(kband) <= 220 * 1024) {
      const int knb = (Ph + kband - 1) / kband;
      const size_t sm = ksmem(kband);
#define SRL_U8_LAUNCH(HH)                                                                  \
  do {                                                                                     \
    SRL_CUDA(cudaFuncSetAttribute(maxplus_u8_key_kernel<HH>,                               \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));  \
    maxplus_u8_key_kernel<HH><<<E * R * knb, kKeyThreads, sm, stream>>>(                   \
        walls, rocks, level, out, R, H, W, h, kband, knb);                                 \
  } while (0)
      if (h == 32) SRL_U8_LAUNCH(32);
      else if (h == 16) SRL_U8_LAUNCH(16);
      else if (h == 8) SRL_U8_LAUNCH(8);
      else SRL_U8_LAUNCH(0);
#undef SRL_U8_LAUNCH
      return check_launch("maxplus_u8_key_kernel");
    }
  }
  const int nbands = (Ph + band - 1) / band;
  const size_t smem = smem_for(band);
  SRL_CUDA(cudaFuncSetAttribute(maxplus_u8_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  maxplus_u8_kernel<<<E * R * nbands, kThreads, smem, stream>>>(walls, rocks, level, out, R,
                                                             H, W, h, band, nbands);
  return check_launch("maxplus_u8_kernel");
}

}  // namespace srl
