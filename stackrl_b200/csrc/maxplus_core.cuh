// Shared pieces of the max-plus kernels (maxplus.cu) and the fused scoring kernel
// (score.cu): fast index division, kernel parameters, the register-tiled
// (add, max) sweep, and the rock normalise/mask helper.  See maxplus.cu for the
// design notes.
#pragma once

#include <stdlib.h>

#include "common.cuh"

namespace srl {

// q = n / d for n * d < 2^32 (all index spaces here are < 2^16 x 2^16).
struct FastDiv {
  uint32_t mul, d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (uint32_t)d;
  f.mul = d == 1 ? 0u : (uint32_t)(((1ull << 32) + (uint32_t)d - 1) / (uint32_t)d);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv f) {
  return f.d == 1 ? n : __umulhi(n, f.mul);
}
__device__ __forceinline__ void fdivmod(uint32_t n, const FastDiv f, uint32_t& q,
                                        uint32_t& r) {
  q = fdiv(n, f);
  r = n - q * f.d;
}

constexpr int kNoQuantum = 0x7fffffff;

struct MaxPlusParams {
  const float* walls;
  const float* rocks;
  const float* level;
  float* out;
  int E, R, H, W, h;
  int Ph, Pw;
  int hp;           // rock columns padded to a multiple of VC
  int Ws;           // smem wall row stride (floats), Ws % 4 == 0, (Ws/4) odd
  int wall_stride;  // smem floats per wall  (H * Ws)
  int rock_stride;  // smem floats per rock copy (h * hp + 4)
  int G;            // environments per CTA (group)
  int RC;           // rotations per CTA
  int rchunks;      // ceil(R / RC)
  int strips;       // strips per output row; strip k starts at column k*(T-1)
  int ngroups;      // ceil(E / G)
  float threshold;
  int tma_wall;     // wall rows can be bulk-copied (W % 4 == 0, 16-B aligned base)
  int tma_rock;     // rock rows can be bulk-copied (h % 4 == 0, 16-B aligned base)
  int stage_out;    // staged kernel: score maps leave by bulk TMA store
  int band, nbands; // direct kernel: output rows per CTA / CTAs per (group, chunk)
  int qlog2;        // stream kernel: raw values are expected to be multiples of
                    // 2^qlog2 (kNoQuantum: no hint, float sweep only)
  float qscale, qunit;   // 2^-qlog2, 2^qlog2
  int wsw16, need16;     // 16-bit layout: wall row stride (words), words a row read needs
  // stream kernel (maxplus_stream_kernel)
  int nslot;        // environments resident in the compute-layout ring
  int ipe;          // items (rotation, output row, strip) per environment
  long long items;  // E * ipe
  int units;        // ceil(items / 32): one warp-pass each
  FastDiv dPh, dStrips, dRC, dW, dH, dh, dhp, dW4, dh4, dIpe, dNslot;
};

// Block of T x VC (add, max) cells: acc[t] = max(acc[t], row[t+v] + nv[v]).
// PAIRED: nvs[v] = nv[v+1] is the one-column-shifted rock row, loaded from its
// own smem copy so that (nvs[v], nvs[v+1]) for even v is an aligned register
// pair holding (nv[v+1], nv[v+2]).
//
// IMAX: the 3-input max is VIMNMX3 on the float bit patterns (signed 32-bit
// integer order) instead of FMNMX3.  Exact whenever every operand is a
// non-negative float or -inf, i.e. no wall or live rock value of the tile has
// its sign bit set: non-negative floats order like their bit patterns, and -inf
// (0xff800000) is below all of them as a signed integer too.  The callers test
// that precondition per environment and fall back to FMNMX3 otherwise.
template <bool IMAX>
__device__ __forceinline__ float max3(float a, float b, float c) {
  if constexpr (IMAX)
    return __int_as_float(__vimax3_s32(__float_as_int(a), __float_as_int(b),
                                       __float_as_int(c)));
  else
    return fmax3(a, b, c);
}

template <int T, int VC, bool PAIRED, bool IMAX = false>
__device__ __forceinline__ void cell_block(float (&acc)[T],
                                           const float (&row)[4 * ((T + VC + 2) / 4)],
                                           const float (&nv)[VC],
                                           const float (&nvs)[VC]) {
  if constexpr (!PAIRED) {
#pragma unroll
    for (int v = 0; v < VC; ++v) {
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = fmaxf(acc[t], row[t + v] + nv[v]);
    }
  } else {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if ((t & 1) == 0) {
        // even output column: wall index k = t+v is even for even v.
#pragma unroll
        for (int v = 0; v < VC; v += 2) {
          float s0, s1;
          fadd2(s0, s1, row[t + v], row[t + v + 1], nv[v], nv[v + 1]);
          acc[t] = max3<IMAX>(acc[t], s0, s1);
        }
      } else {
        // odd output column: pair odd v with v+1 (k = t+v even), using the
        // shifted rock row; v = 0 and v = VC-1 stay single.
        float e0 = row[t] + nv[0];
        float e1 = row[t + VC - 1] + nv[VC - 1];
        acc[t] = max3<IMAX>(acc[t], e0, e1);
#pragma unroll
        for (int v = 1; v + 1 < VC; v += 2) {
          float s0, s1;
          fadd2(s0, s1, row[t + v], row[t + v + 1], nvs[v - 1], nvs[v]);
          acc[t] = max3<IMAX>(acc[t], s0, s1);
        }
      }
    }
  }
}

// All (add, max) cells of one item: T outputs of one output row against one rock.
template <int T, int VC, int PAIRED, bool IMAX = false>   // PAIRED 0: FADD + FMNMX per cell, 1: FADD2 + 3-input max per cell pair
__device__ __forceinline__ void sweep_item(float (&acc)[T], const float* wbase,
                                           const float* rbase, const float* sbase,
                                           int h, int hp, int Ws) {
  constexpr int NR4 = (T + VC + 2) / 4;   // float4 loads per wall row chunk
  static_assert((T - 1) % 4 == 0 && 4 * NR4 >= T + VC - 1, "tile shape");
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = kNegInf;
  // One flat loop over the (rock row, column chunk) pairs: rock rows are
  // contiguous (hp floats each), the wall pointer skips to the next row after
  // the last chunk of a row.
  const int cpr = hp / VC, nchunks = h * cpr, wskip = Ws - hp;
  int j = 0;
#pragma unroll 2
  for (int c = 0; c < nchunks; ++c) {
    float row[4 * NR4];
    float nv[VC];
    float nvs[VC];
#pragma unroll
    for (int k = 0; k < NR4; ++k) {
      const float4 x = lds128(wbase + 4 * k);
      row[4 * k + 0] = x.x;
      row[4 * k + 1] = x.y;
      row[4 * k + 2] = x.z;
      row[4 * k + 3] = x.w;
    }
#pragma unroll
    for (int k = 0; k < VC / 4; ++k) {
      const float4 x = lds128(rbase + 4 * k);
      nv[4 * k + 0] = x.x;
      nv[4 * k + 1] = x.y;
      nv[4 * k + 2] = x.z;
      nv[4 * k + 3] = x.w;
      if constexpr (PAIRED != 0) {
        const float4 y = lds128(sbase + 4 * k);
        nvs[4 * k + 0] = y.x;
        nvs[4 * k + 1] = y.y;
        nvs[4 * k + 2] = y.z;
        nvs[4 * k + 3] = y.w;
      }
    }
    cell_block<T, VC, PAIRED == 1, IMAX>(acc, row, nv, nvs);
    wbase += VC;
    rbase += VC;
    sbase += VC;
    if (++j == cpr) {
      j = 0;
      wbase += wskip;
    }
  }
}

// ---- 16-bit fixed-point sweep (DPX) --------------------------------------------- //
// When every wall and rock value of an environment is a non-negative multiple of
// one power of two q and smaller than 2^14 q (heightmaps straight from the
// rasteriser: the reference's float32 depth->elevation formulas leave multiples of
// ulp(1000) = 2^-14 m, observer.py:259-260/274-275), the float32 sums are exact and
// so is 16-bit integer arithmetic on the counts x / q.  The sweep then needs ONE
// instruction per two cells, VIADDMNMX.S16x2 (acc = max(acc, wall + rock) on two
// packed 16-bit lanes), instead of FADD2 + VIMNMX3.
//
// Packed layout: 32-bit words of two consecutive columns.  wall A[m] = (w[2m],
// w[2m+1]), wall B[m] = (w[2m+1], w[2m+2]) (even / odd output columns), rock
// N[m] = (n[2m], n[2m+1]); masked rock cells hold kMask16, so their sums stay
// negative and never win.  acc[t] holds the running maxima over even (low half)
// and odd (high half) rock columns of output t.
constexpr int kMask16 = -20000;          // 16383 + kMask16 < 0, 2 * kMask16 > -32768... never added twice
constexpr uint32_t kAccInit16 = 0x80008000u;

__device__ __forceinline__ uint4 lds128u(const uint32_t* p) {
  return *reinterpret_cast<const uint4*>(p);
}

// All cells of one item: T outputs of one output row against one rock.
// wa/wb: the item's first word of the A / B wall copies (row stride wsw words),
// rk: the rock's first word (rows are hp/2 words, contiguous).
template <int T, int VC>
__device__ __forceinline__ void sweep_item16(uint32_t (&acc)[T], const uint32_t* wa,
                                             const uint32_t* wb, const uint32_t* rk, int h,
                                             int hp, int wsw) {
  static_assert((T - 1) % 8 == 0 && VC % 8 == 0, "16-bit sweep: 16-byte aligned strips");
  constexpr int PV = VC / 2;                     // rock pair words per chunk
  constexpr int NA = (T + 1) / 2 + PV - 1;       // A words: even t, pairs t/2 .. t/2+PV-1
  constexpr int NB = (T - 1) / 2 + PV - 1;       // B words: odd t
  constexpr int NA4 = (NA + 3) / 4, NB4 = (NB + 3) / 4;
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = kAccInit16;
  const int cpr = hp / VC, nchunks = h * cpr, wskip = wsw - hp / 2;
  int j = 0;
#pragma unroll 2
  for (int c = 0; c < nchunks; ++c) {
    uint32_t a[4 * NA4], b[4 * NB4], n[PV];
#pragma unroll
    for (int k = 0; k < NA4; ++k) {
      const uint4 x = lds128u(wa + 4 * k);
      a[4 * k] = x.x; a[4 * k + 1] = x.y; a[4 * k + 2] = x.z; a[4 * k + 3] = x.w;
    }
#pragma unroll
    for (int k = 0; k < NB4; ++k) {
      const uint4 x = lds128u(wb + 4 * k);
      b[4 * k] = x.x; b[4 * k + 1] = x.y; b[4 * k + 2] = x.z; b[4 * k + 3] = x.w;
    }
#pragma unroll
    for (int k = 0; k < PV / 4; ++k) {
      const uint4 x = lds128u(rk + 4 * k);
      n[4 * k] = x.x; n[4 * k + 1] = x.y; n[4 * k + 2] = x.z; n[4 * k + 3] = x.w;
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int v = 0; v < PV; ++v) {
        const uint32_t w = (t & 1) ? b[(t - 1) / 2 + v] : a[t / 2 + v];
        acc[t] = __viaddmax_s16x2(w, n[v], acc[t]);
      }
    }
    wa += PV;
    wb += PV;
    rk += PV;
    if (++j == cpr) {
      j = 0;
      wa += wskip;
      wb += wskip;
    }
  }
}

// IEEE x / level.  Zero numerators (rock background, empty wall) are common and
// would take __fdiv_rn's slow path; +-0 / level = +-0 * level bit for bit (the
// level is a finite, non-zero goal height).
__device__ __forceinline__ float div_level(float x, float level) {
  return x == 0.f ? __fmul_rn(x, level) : __fdiv_rn(x, level);
}

// Division by a goal level that is a power of two (the reference default,
// max_z - object_z = 0.25, is one) is an exact scaling: x / 2^k == x * 2^-k bit
// for bit (both are the correctly rounded value of the same real number, also in
// the subnormal and overflow ranges), so the prep pass may multiply instead.
// `inv` is 0 when the level is not a (normal) power of two whose inverse is
// normal too, and the IEEE division is used.
__device__ __forceinline__ float pow2_inverse(float level) {
  const uint32_t b = __float_as_uint(level);
  const uint32_t ex = (b >> 23) & 0xff;
  if ((b & 0x807fffffu) != 0u || ex < 64 || ex > 190) return 0.f;
  return __uint_as_float((254u - ex) << 23);
}
__device__ __forceinline__ float div_level(float x, float level, float inv) {
  return inv != 0.f ? __fmul_rn(x, inv) : div_level(x, level);
}

// Normalise + mask one rock value (baselines.py:24-25, :32).
__device__ __forceinline__ float prep_rock(float n, bool scaled, float level, float inv,
                                           float thr, bool& dead) {
  if (scaled) n = div_level(n, level, inv);
  const bool live = n > thr;
  dead = dead || !live;
  return live ? n : kNegInf;
}

// ---- host-side tile selection ------------------------------------------------ //
struct Choice {
  int T, VC;
};

// Per-thread tile widths with an instantiation: T = 4m+1 outputs, strips pitched
// S = 4m apart (16-B aligned starts).  The reference geometries have
// Pw = 2^a - 2^b + 1 == 1 (mod 4), which these cover with no wasted column.
static const int kTs[] = {5, 9, 13, 17, 21, 25};

inline int strips_for(int Pw, int T) {
  return Pw <= T ? 1 : (Pw - T + (T - 1) - 1) / (T - 1) + 1;
}

inline Choice choose_tile(int Pw, int h) {
  Choice c;
  c.VC = h >= 13 ? 16 : (h >= 5 ? 8 : 4);
  if (const char* s = getenv("SRL_MP_VC")) {   // tuning override
    const int v = atoi(s);
    if (v == 4 || v == 8 || v == 16) c.VC = v;
  }
  int best = kTs[0];
  double best_cost = 1e30;
  for (int T : kTs) {
    if (c.VC == 16 && T > 21) continue;   // register budget (112/thread)
    const int strips = strips_for(Pw, T);
    const double waste = (double)strips * T / Pw;
    const double loads = ((T + c.VC + 2) / 4 + c.VC / 2) / (double)(T * c.VC);
    const double cost = waste * (1.03 + loads);
    if (cost < best_cost - 1e-12) {
      best_cost = cost;
      best = T;
    }
  }
  c.T = best;
  return c;
}

inline int pick_threads(int items, int cap) {
  if (items <= cap) return round_up(items < 32 ? 32 : items, 32);
  int best_t = cap;
  double best_w = 1e9;
  for (int t = 192; t <= cap; t += 32) {   // several passes: waste the fewest lanes
    const int passes = (items + t - 1) / t;
    const double w = (double)passes * t / items;
    if (w < best_w - 1e-9) {
      best_w = w;
      best_t = t;
    }
  }
  return best_t;
}

}  // namespace srl
