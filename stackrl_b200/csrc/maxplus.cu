// Batched max-plus "drop" search (reference: stackrl/baselines.py:21-43,
// get_inputs + height; the Python double loop over positions at :34-41).
//
//   out[e,r,i,j] = max_{u,v}( n[e,r,u,v] > thr ? o[e,i+u,j+v] + n[e,r,u,v] : 0 )
//   o = wall/level, n = rock/level  (IEEE float32 division, then float32 add)
//
// Design ("maxplus_f32"):
//   * Shared-memory compute layout per CTA: G walls on a padded row stride
//     (Ws/4 odd => the per-row LDS.128 of 8 consecutive lanes hit 8 distinct
//     16-B bank groups) and, per rock rotation, the normalised rock with the
//     `n > thr` mask folded in as -inf plus a one-column-shifted copy of it.
//     With the mask folded in the inner loop has no select: o + (-inf) = -inf
//     never wins the max.  The reference's "masked cell contributes 0" (quirk
//     Q2) becomes one max(acc, 0) at the end, applied only when the rock really
//     has a masked cell.
//   * Each thread owns T = 4m+1 consecutive outputs of one output row (strips
//     pitched 4m apart).  Per rock row it loads T+VC-1 wall values and VC rock
//     values with LDS.128 and runs the fully unrolled T x VC block of (add,
//     max) cells out of registers.
//   * sm_100 instruction mix: two adds issue as one FADD2 (add.rn.f32x2, FMA
//     pipe) on aligned register pairs and fold into the accumulator with one
//     3-input FMNMX3 (ALU pipe): ~1.03 issue slots per cell instead of 2.  Max
//     is exact and order independent and each add is one IEEE rn add, so the
//     result is bit-identical to numpy's.
//   * Two kernels share that sweep:
//       - maxplus_staged_kernel (small walls, e.g. 32x32 / 64x64): persistent
//         CTAs, 2 per SM.  Group k+1's raw walls and rocks arrive by two 1-D
//         bulk TMA copies (cp.async.bulk + mbarrier) while group k is swept;
//         a prep pass converts raw -> compute layout (division by the goal
//         level, mask, shifted copy); score maps are staged in shared memory
//         and leave with one bulk TMA store per group.
//       - maxplus_direct_kernel (big walls, e.g. 128x128 with 36 rotations):
//         one CTA per (environment, rotation chunk); wall rows are bulk-copied
//         straight onto the padded stride and normalised in place.
//   * No tensor cores: max-plus is not a dense contraction.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "maxplus_core.cuh"

namespace srl {

// --------------------------------------------------------------------------- //
// Staged persistent kernel (small walls).
// --------------------------------------------------------------------------- //
template <int T, int VC, int PAIRED>
__global__ void __launch_bounds__(288, 2)
maxplus_staged_kernel(const MaxPlusParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = p.H, W = p.W, h = p.h, hp = p.hp, Ws = p.Ws, R = p.R, G = p.G;
  const int Ph = p.Ph, Pw = p.Pw, P = Ph * Pw;
  constexpr int S = T - 1;

  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  int* masked = reinterpret_cast<int*>(smem_raw + 16);          // [2][G*R]
  const int slots = G * R;
  const int flag_bytes = round_up(2 * slots * 4, 16);
  float* raw_wall = reinterpret_cast<float*>(smem_raw + 16 + flag_bytes);
  float* raw_rock = raw_wall + G * H * W;
  float* wall_s = raw_rock + slots * h * h;
  float* rock_s = wall_s + G * p.wall_stride;
  float* rock_sh = rock_s + slots * p.rock_stride;
  float* out_s = rock_sh + slots * p.rock_stride;               // [G*R*P] if stage_out

  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  // One-time fills of everything the per-group prep never rewrites: wall pad
  // columns [W, Ws) (finite), rock pad columns [h, hp) and the last real column
  // of the shifted copy (-inf: never win, never count as masked).
  for (int k = tid; k < G * p.wall_stride; k += nthreads) wall_s[k] = 0.f;
  for (int k = tid; k < 2 * slots * p.rock_stride; k += nthreads) rock_s[k] = kNegInf;
  for (int k = tid; k < 2 * slots; k += nthreads) masked[k] = 0;
  __syncthreads();

  auto issue_loads = [&](int g) {
    const int e0 = g * G;
    const int Gv = min(G, p.E - e0);
    const uint32_t wb = (uint32_t)Gv * H * W * 4, rb = (uint32_t)Gv * R * h * h * 4;
    mbar_arrive_expect_tx(bar, wb + rb);
    tma_load_1d(raw_wall, p.walls + (size_t)e0 * H * W, wb, bar);
    tma_load_1d(raw_rock, p.rocks + (size_t)e0 * R * h * h, rb, bar);
  };

  if (tid == 0 && (int)blockIdx.x < p.ngroups) issue_loads(blockIdx.x);

  const int W4 = W / 4, h4 = h / 4;
  const bool scaled = p.level != nullptr;
  int it = 0;
  for (int g = blockIdx.x; g < p.ngroups; g += gridDim.x, ++it) {
    const int e0 = g * G;
    const int Gv = min(G, p.E - e0);
    int* flags = masked + (it & 1) * slots;
    int* flags_next = masked + ((it + 1) & 1) * slots;

    mbar_wait(bar, it & 1);

    // ---- prep: raw -> compute layout ---------------------------------------- //
    for (int k = tid; k < slots; k += nthreads) flags_next[k] = 0;
    for (uint32_t q = tid; q < (uint32_t)(Gv * H * W4); q += nthreads) {
      uint32_t row, c4;
      fdivmod(q, p.dW4, row, c4);
      float4 x = lds128(raw_wall + 4 * q);
      if (scaled) {
        const float lv = __ldg(p.level + e0 + fdiv(row, p.dH));
        const float inv = pow2_inverse(lv);
        x.x = div_level(x.x, lv, inv);
        x.y = div_level(x.y, lv, inv);
        x.z = div_level(x.z, lv, inv);
        x.w = div_level(x.w, lv, inv);
      }
      *reinterpret_cast<float4*>(wall_s + row * Ws + 4 * c4) = x;
    }
    for (uint32_t q = tid; q < (uint32_t)(Gv * R * h * h4); q += nthreads) {
      uint32_t rrow, c4, slot, u;
      fdivmod(q, p.dh4, rrow, c4);
      fdivmod(rrow, p.dh, slot, u);
      const float lv = scaled ? __ldg(p.level + e0 + fdiv(slot, p.dRC)) : 1.f;
      float4 x = lds128(raw_rock + 4 * q);
      bool dead = false;
      const float inv = scaled ? pow2_inverse(lv) : 0.f;
      x.x = prep_rock(x.x, scaled, lv, inv, p.threshold, dead);
      x.y = prep_rock(x.y, scaled, lv, inv, p.threshold, dead);
      x.z = prep_rock(x.z, scaled, lv, inv, p.threshold, dead);
      x.w = prep_rock(x.w, scaled, lv, inv, p.threshold, dead);
      if (dead) flags[slot] = 1;
      float* dst = rock_s + slot * p.rock_stride + u * hp + 4 * c4;
      *reinterpret_cast<float4*>(dst) = x;
      if constexpr (PAIRED != 0) {
        float* sh = rock_sh + slot * p.rock_stride + u * hp + 4 * c4;
        if (c4 != 0) sh[-1] = x.x;
        sh[0] = x.y;
        sh[1] = x.z;
        sh[2] = x.w;
      }
    }
    // The previous group's bulk store must have finished READING out_s before
    // this group's sweep overwrites it.
    if (tid == 0 && p.stage_out) tma_store_wait_read();
    __syncthreads();

    // ---- prefetch the next group while this one is swept ---------------------- //
    // (raw_* were last read through the generic proxy in the prep above)
    if (tid == 0 && g + (int)gridDim.x < p.ngroups) {
      fence_proxy_async();
      issue_loads(g + gridDim.x);
    }

    // ---- sweep ---------------------------------------------------------------- //
    const int items = Gv * R * p.strips * Ph;
    for (int item = tid; item < items; item += nthreads) {
      uint32_t rest, i, strip, slot;
      fdivmod((uint32_t)item, p.dPh, rest, i);
      fdivmod(rest, p.dStrips, slot, strip);
      const uint32_t el = fdiv(slot, p.dRC);
      float acc[T];
      sweep_item<T, VC, PAIRED>(acc, wall_s + el * p.wall_stride + i * Ws + strip * S,
                                rock_s + slot * p.rock_stride,
                                rock_sh + slot * p.rock_stride, h, hp, Ws);
      const bool floor0 = flags[slot] != 0;
      const int ncols = ((int)strip == p.strips - 1) ? min(T, Pw - (int)strip * S) : S;
      const size_t off = (size_t)slot * P + i * Pw + strip * S;
      if (p.stage_out) {
        float* o = out_s + off;
#pragma unroll
        for (int t = 0; t < T; ++t)
          if (t < ncols) o[t] = floor0 ? fmaxf(acc[t], 0.f) : acc[t];
      } else {
        float* o = p.out + (size_t)e0 * R * P + off;
#pragma unroll
        for (int t = 0; t < T; ++t)
          if (t < ncols) __stcs(o + t, floor0 ? fmaxf(acc[t], 0.f) : acc[t]);
      }
    }
    if (p.stage_out) {
      fence_proxy_async();          // generic-proxy writes -> visible to the TMA store
      __syncthreads();
      if (tid == 0) {
        tma_store_1d(p.out + (size_t)e0 * R * P, out_s, (uint32_t)Gv * R * P * 4);
        tma_store_commit();
      }
    } else {
      __syncthreads();              // compute layout is rewritten by the next prep
    }
  }
  if (tid == 0 && p.stage_out) tma_store_wait_all();
}

// --------------------------------------------------------------------------- //
// Stream kernel (small walls; default): warp-specialised, no CTA-wide barrier.
//
// One persistent CTA per SM = kConsumers sweep warps + kProducers prep warps.
// The CTA owns a contiguous range of the global item stream (item = one strip of
// one output row of one (environment, rotation); 32 consecutive items = one
// "unit" = one warp pass), cut in whole units so that every SM sub-partition
// carries the same number of warp passes whatever the map size (the 17 x 17 maps
// of the 32/16 geometry give 136 items per environment = 4.25 warps).
//   producers (one group of kProducers warps, environments strictly in order):
//     raw wall + rocks of environment k arrive by bulk TMA into raw buffer
//     k % kRawDepth; the group converts them to the compute layout in ring slot
//     k % nslot (division by the goal level, mask as -inf, shifted copy, "has
//     masked cell" / "has negative value" flags), re-arms the TMA for
//     environment k + kRawDepth and publishes ready[slot] = k + 1.
//   consumers: wait until ready[slot] names the environment(s) of their 32
//     items, sweep, stage the 32 row segments (contiguous in the output tensor)
//     in a per-warp buffer, arrive on empty[slot] with the number of items they
//     finished, and send the segment block off with one bulk TMA store.
// A slot is rewritten only after all `ipe` items of its previous environment
// arrived on empty[slot]; production is sequential, so a waiter is never more
// than one mbarrier phase away, and `ready` carries the environment number
// itself, so a consumer that runs far ahead cannot mistake an older tenant.
// --------------------------------------------------------------------------- //
#ifndef SRL_CONSUMERS
#define SRL_CONSUMERS 16
#endif
#ifndef SRL_PRODUCERS
#define SRL_PRODUCERS 8
#endif
// Register split of the warp-specialised kernel (setmaxnreg): 24 warps start at
// the launch-bound cap of 80 registers; the two producer warpgroups give
// registers back to the CTA pool, the four consumer warpgroups take them.  The
// pool only holds what was released, so 8 * (80 - P) >= 16 * (C - 80) must hold
// (P = 64, C = 88 is the measured optimum: the 32/16 geometry is producer-bound,
// the 64/16 one consumer-bound).
#ifndef SRL_PRODUCER_REGS
#define SRL_PRODUCER_REGS 64
#endif
#ifndef SRL_CONSUMER_REGS
#define SRL_CONSUMER_REGS 88
#endif
constexpr int kConsumers = SRL_CONSUMERS, kProducers = SRL_PRODUCERS, kRawDepth = 2;

// Producer-side conversion of one environment, raw -> compute layout, by the
// kProducerThreads threads of the producer group (thread index pt).  Returns the
// OR of the bit patterns of every value that takes part in a sum (sign bit set
// <=> some value is negative: the integer-max sweep is not usable then).
template <int MODE>
__device__ __forceinline__ float scale_level(float x, float lv, float inv) {
  if constexpr (MODE == 0) return x;
  else if constexpr (MODE == 1) return __fmul_rn(x, inv);
  else return div_level(x, lv);
}
template <int MODE>
__device__ __forceinline__ uint32_t convert_env(const MaxPlusParams& p, const float* raw,
                                                float* wall_s, float* rock_s,
                                                float* rock_sh, int* flags, float lv,
                                                float inv, int pt) {
  const int H = p.H, W = p.W, h = p.h, hp = p.hp, Ws = p.Ws, R = p.R;
  const uint32_t W4 = W / 4, h4 = h / 4;
  const int lane = pt & 31;
  uint32_t neg = 0;
  const uint32_t nw = (uint32_t)H * W4;
#pragma unroll 2
  for (uint32_t q = pt; q < nw; q += (kProducers * 32)) {
    uint32_t row, c4;
    fdivmod(q, p.dW4, row, c4);
    float4 x = lds128(raw + 4 * q);
    x.x = scale_level<MODE>(x.x, lv, inv);
    x.y = scale_level<MODE>(x.y, lv, inv);
    x.z = scale_level<MODE>(x.z, lv, inv);
    x.w = scale_level<MODE>(x.w, lv, inv);
    neg |= __float_as_uint(x.x) | __float_as_uint(x.y) | __float_as_uint(x.z) |
           __float_as_uint(x.w);
    *reinterpret_cast<float4*>(wall_s + row * Ws + 4 * c4) = x;
  }
  // Live rock values are > threshold: they can only be negative when the
  // threshold is.
  const bool rock_sign = p.threshold < 0.f;
  const float thr = p.threshold;
  const float* rraw = raw + H * W;
  const uint32_t nr = (uint32_t)R * h * h4;
  // The trip count is the same for all lanes of a warp (the tail is handled by
  // clamping q), so the shuffle below is always executed by full warps.
  for (uint32_t q0 = pt - lane; q0 < nr; q0 += (kProducers * 32)) {
    const uint32_t q = q0 + lane;
    const bool in = q < nr;
    const uint32_t qq = in ? q : nr - 1;
    uint32_t rrow, c4, slot, u;
    fdivmod(qq, p.dh4, rrow, c4);
    fdivmod(rrow, p.dh, slot, u);
    float4 x = lds128(rraw + 4 * qq);
    x.x = scale_level<MODE>(x.x, lv, inv);
    x.y = scale_level<MODE>(x.y, lv, inv);
    x.z = scale_level<MODE>(x.z, lv, inv);
    x.w = scale_level<MODE>(x.w, lv, inv);
    const bool lx = x.x > thr, ly = x.y > thr, lz = x.z > thr, lw = x.w > thr;
    x.x = lx ? x.x : kNegInf;
    x.y = ly ? x.y : kNegInf;
    x.z = lz ? x.z : kNegInf;
    x.w = lw ? x.w : kNegInf;
    if (in && !(lx && ly && lz && lw)) flags[slot] = 1;
    if (rock_sign)   // -inf marks a masked cell, not a negative value
      neg |= (lx ? __float_as_uint(x.x) : 0u) | (ly ? __float_as_uint(x.y) : 0u) |
             (lz ? __float_as_uint(x.z) : 0u) | (lw ? __float_as_uint(x.w) : 0u);
    // Shifted copy: sh[c] = n[c + 1]; the first value of the next float4 of the
    // same rock row sits in the next lane (or is fetched directly by lane 31).
    float nx = __shfl_down_sync(0xffffffffu, x.x, 1);
    if (c4 == h4 - 1) {
      nx = kNegInf;
    } else if (lane == 31) {
      nx = scale_level<MODE>(rraw[4 * qq + 4], lv, inv);
      nx = nx > thr ? nx : kNegInf;
    }
    if (in) {
      const uint32_t off = slot * p.rock_stride + u * hp + 4 * c4;
      *reinterpret_cast<float4*>(rock_s + off) = x;
      *reinterpret_cast<float4*>(rock_sh + off) = make_float4(x.y, x.z, x.w, nx);
    }
  }
  return neg;
}

constexpr int kStreamThreads = (kConsumers + kProducers) * 32;
constexpr int kProducerThreads = kProducers * 32;
static_assert(SRL_PRODUCER_REGS == 0 ||
                  kProducers * (80 - SRL_PRODUCER_REGS) >= kConsumers * (SRL_CONSUMER_REGS - 80),
              "setmaxnreg: the consumers cannot take more registers than the producers release");

// ---- 16-bit fixed-point layout (see maxplus_core.cuh, "16-bit fixed-point sweep") -- //
struct Layout16 {
  int wsw;        // wall row stride in words (two columns per word), wsw % 4 == 0, (wsw/4) odd
  int rstride;    // words per rock: h * hp / 2 + 4
  int words;      // words per environment: 2 * H * wsw + R * rstride
};
__host__ __device__ inline Layout16 layout16(int H, int R, int h, int hp, int need_words) {
  Layout16 l;
  l.wsw = round_up(need_words, 4);
  if ((l.wsw / 4) % 2 == 0) l.wsw += 4;
  l.rstride = h * hp / 2 + 4;
  l.words = 2 * H * l.wsw + R * l.rstride;
  return l;
}

// x * sc must be an integer count in [0, 2^14): t + 1.5 * 2^23 keeps an integer t
// exactly in the low mantissa bits.
__device__ __forceinline__ bool count16(float x, float sc, int& v) {
  const float t = __fmul_rn(x, sc);
  const float r = __fadd_rn(t, 12582912.f);
  v = __float_as_int(r) - 0x4B400000;
  return __fadd_rn(r, -12582912.f) == t && (unsigned)v < 16384u;
}
__device__ __forceinline__ uint32_t pack16(int lo, int hi) {
  return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410);
}

// Pad words of one slot for the 16-bit layout (written when a slot changes layout).
__device__ __forceinline__ void init_pads16(uint32_t* slot, const Layout16 l, int H, int W,
                                            int R, int h, int hp, int pt) {
  const int padw = l.wsw - W / 2;
  for (int k = pt; k < 2 * H * padw; k += kProducers * 32) {
    const int row = k / padw, c = k - row * padw;
    slot[row * l.wsw + W / 2 + c] = 0u;
  }
  uint32_t* rocks = slot + 2 * H * l.wsw;
  const uint32_t m2 = pack16(kMask16, kMask16);
  for (int k = pt; k < R * 4; k += kProducers * 32)
    rocks[(k >> 2) * l.rstride + h * hp / 2 + (k & 3)] = m2;
  const int padr = (hp - h) / 2;
  for (int k = pt; k < R * h * padr; k += kProducers * 32) {
    const int row = k / padr, c = k - row * padr;
    const int r = row / h, u = row - r * h;
    rocks[r * l.rstride + u * (hp / 2) + h / 2 + c] = m2;
  }
}
// Pad floats of one slot for the float layout.
__device__ __forceinline__ void init_padsf(float* slot, const MaxPlusParams& p, int pt) {
  const int H = p.H, W = p.W, h = p.h, hp = p.hp, Ws = p.Ws, R = p.R;
  const int padw = Ws - W, padr = hp - h;
  for (int k = pt; k < H * padw; k += kProducers * 32) {
    const int row = k / padw, c = k - row * padw;
    slot[row * Ws + W + c] = 0.f;
  }
  for (int k = pt; k < 2 * R * 4; k += kProducers * 32)
    slot[p.wall_stride + (k >> 2) * p.rock_stride + h * hp + (k & 3)] = kNegInf;
  for (int k = pt; k < 2 * R * h * padr; k += kProducers * 32) {
    const int row = k / padr, c = k - row * padr;
    const int r = row / h, u = row - r * h;
    slot[p.wall_stride + r * p.rock_stride + u * hp + h + c] = kNegInf;
  }
}

// raw -> 16-bit layout.  Returns false as soon as a value is not a count in
// [0, 2^14) (the caller then converts the environment to the float layout).
// `scaled`: live test is (x * inv) > thr like the float path.
__device__ __forceinline__ bool convert_env16(const MaxPlusParams& p, const float* raw,
                                              uint32_t* slot, const Layout16 l, int* flags,
                                              float inv, float sc, int pt) {
  const int H = p.H, W = p.W, h = p.h, hp = p.hp, R = p.R;
  const uint32_t W4 = W / 4, h4 = h / 4;
  const int lane = pt & 31;
  bool ok = true;
  uint32_t* A = slot;
  uint32_t* B = slot + H * l.wsw;
  uint32_t* rocks = slot + 2 * H * l.wsw;
  const uint32_t nw = (uint32_t)H * W4;
  for (uint32_t q0 = pt - lane; q0 < nw; q0 += kProducers * 32) {
    const uint32_t q = q0 + lane;
    const bool in = q < nw;
    const uint32_t qq = in ? q : nw - 1;
    uint32_t row, c4;
    fdivmod(qq, p.dW4, row, c4);
    const float4 x = lds128(raw + 4 * qq);
    int v0, v1, v2, v3;
    ok = count16(x.x, sc, v0) & count16(x.y, sc, v1) & count16(x.z, sc, v2) &
         count16(x.w, sc, v3) & ok;
    int nx = __shfl_down_sync(0xffffffffu, v0, 1);
    if (c4 == W4 - 1) {
      nx = 0;                                    // w[W]: pad column
    } else if (lane == 31) {
      ok = count16(raw[4 * qq + 4], sc, nx) & ok;
    }
    if (in) {
      *reinterpret_cast<uint2*>(A + row * l.wsw + 2 * c4) =
          make_uint2(pack16(v0, v1), pack16(v2, v3));
      *reinterpret_cast<uint2*>(B + row * l.wsw + 2 * c4) =
          make_uint2(pack16(v1, v2), pack16(v3, nx));
    }
  }
  const float thr = p.threshold;
  const float* rraw = raw + H * W;
  const uint32_t nr = (uint32_t)R * h * h4;
  for (uint32_t q = pt; q < nr; q += kProducers * 32) {
    uint32_t rrow, c4, slot_i, u;
    fdivmod(q, p.dh4, rrow, c4);
    fdivmod(rrow, p.dh, slot_i, u);
    const float4 x = lds128(rraw + 4 * q);
    const bool lx = __fmul_rn(x.x, inv) > thr, ly = __fmul_rn(x.y, inv) > thr,
               lz = __fmul_rn(x.z, inv) > thr, lw = __fmul_rn(x.w, inv) > thr;
    int v0, v1, v2, v3;
    const bool o0 = count16(x.x, sc, v0), o1 = count16(x.y, sc, v1),
               o2 = count16(x.z, sc, v2), o3 = count16(x.w, sc, v3);
    ok = ok & (o0 | !lx) & (o1 | !ly) & (o2 | !lz) & (o3 | !lw);
    if (!(lx && ly && lz && lw)) flags[slot_i] = 1;
    *reinterpret_cast<uint2*>(rocks + slot_i * l.rstride + u * (hp / 2) + 2 * c4) =
        make_uint2(pack16(lx ? v0 : kMask16, ly ? v1 : kMask16),
                   pack16(lz ? v2 : kMask16, lw ? v3 : kMask16));
  }
  return ok;
}

template <int T, int VC>
__global__ void __launch_bounds__(kStreamThreads, 1)
maxplus_stream_kernel(const MaxPlusParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = p.H, W = p.W, h = p.h, hp = p.hp, Ws = p.Ws, R = p.R;
  const int Ph = p.Ph, Pw = p.Pw, P = Ph * Pw, nslot = p.nslot, ipe = p.ipe;
  constexpr int S = T - 1;

  // ---- shared memory carve-up ---------------------------------------------- //
  uint64_t* empty = reinterpret_cast<uint64_t*>(smem_raw);         // [nslot]
  uint64_t* full = empty + nslot;                                  // [nslot]
  uint64_t* rawbar = full + nslot;                                 // [kRawDepth]
  int* ready = reinterpret_cast<int*>(rawbar + kRawDepth);         // [nslot]
  int* negative = ready + nslot;                                   // [nslot]
  int* exact16 = negative + nslot;                                 // [nslot] 16-bit layout?
  int* lay = exact16 + nslot;                                      // [nslot] producer's copy
  float* scale_out = reinterpret_cast<float*>(lay + nslot);        // [nslot] count -> value
  int* bad16 = reinterpret_cast<int*>(scale_out + nslot);          // [1]
  int* masked = bad16 + 1;                                         // [nslot][R]
  const int head =
      round_up((2 * nslot + kRawDepth) * 8 + (5 * nslot + 1 + nslot * R) * 4, 16);
  const Layout16 l16 = layout16(H, R, h, hp, p.need16);
  const int raw_floats = H * W + R * h * h;
  const int env_floats = p.wall_stride + 2 * R * p.rock_stride;
  float* raw = reinterpret_cast<float*>(smem_raw + head);          // [kRawDepth][raw]
  float* ring = raw + kRawDepth * raw_floats;                      // [nslot][env]
  float* stage = ring + (size_t)nslot * env_floats;                // [kConsumers][32*T]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- this CTA's share of the item stream ----------------------------------- //
  const long long u0 = (long long)p.units * blockIdx.x / gridDim.x;
  const long long u1 = (long long)p.units * (blockIdx.x + 1) / gridDim.x;
  const long long it0 = u0 * 32, it1 = min(u1 * 32, p.items);
  if (it0 >= it1) return;
  const int env_first = (int)(it0 / ipe), env_last = (int)((it1 - 1) / ipe);
  const int nenv = env_last - env_first + 1;
  const long long base = (long long)env_first * ipe;   // stream position of local 0

  if (tid == 0) {
    for (int s = 0; s < nslot; ++s) {
      mbar_init(empty + s, ipe);
      mbar_init(full + s, 1);
    }
    for (int k = 0; k < kRawDepth; ++k) mbar_init(rawbar + k, 1);
    fence_barrier_init();
  }
  for (int k = tid; k < nslot; k += kStreamThreads) {
    ready[k] = 0;
    lay[k] = 0;
    exact16[k] = 0;
  }
  // One-time fills of what the producers never rewrite: wall pad columns
  // [W, Ws) (finite) and, in both rock copies, the pad columns [h, hp) and the 4
  // floats behind each rock (-inf: never win, never count as masked).
  {
    const int padw = Ws - W, padr = hp - h;
    for (int k = tid; k < nslot * H * padw; k += kStreamThreads) {
      const int row = k / padw, c = k - row * padw;
      const int s = row / H, r = row - s * H;
      ring[(size_t)s * env_floats + r * Ws + W + c] = 0.f;
    }
    for (int k = tid; k < nslot * 2 * R * 4; k += kStreamThreads) {
      const int rk = k >> 2, s = rk / (2 * R), r = rk - s * 2 * R;
      ring[(size_t)s * env_floats + p.wall_stride + r * p.rock_stride + h * hp + (k & 3)] =
          kNegInf;
    }
    if (padr > 0) {
      for (int k = tid; k < nslot * 2 * R * h * padr; k += kStreamThreads) {
        const int row = k / padr, c = k - row * padr;
        const int rk = row / h, u = row - rk * h;
        const int s = rk / (2 * R), r = rk - s * 2 * R;
        ring[(size_t)s * env_floats + p.wall_stride + r * p.rock_stride + u * hp + h + c] =
            kNegInf;
      }
    }
  }
  __syncthreads();

  if (warp >= kConsumers) {
    // ============================ producer group ============================ //
    if constexpr (kConsumers % 4 == 0 && kProducers % 4 == 0 && SRL_PRODUCER_REGS > 0)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SRL_PRODUCER_REGS));
    const int pt = tid - kConsumers * 32;          // 0 .. kProducerThreads-1
    const uint32_t wb = (uint32_t)H * W * 4, rb = (uint32_t)R * h * h * 4;
    auto issue = [&](int k) {
      const size_t e = (size_t)(env_first + k);
      float* dst = raw + (k % kRawDepth) * raw_floats;
      uint64_t* bar = rawbar + (k % kRawDepth);
      mbar_arrive_expect_tx(bar, wb + rb);
      tma_load_1d(dst, p.walls + e * H * W, wb, bar);
      tma_load_1d(dst + H * W, p.rocks + e * R * h * h, rb, bar);
    };
    if (pt == 0)
      for (int k = 0; k < kRawDepth && k < nenv; ++k) issue(k);
    const bool scaled = p.level != nullptr;
    const int W4 = W / 4, h4 = h / 4;
    for (int k = 0; k < nenv; ++k) {
      uint32_t use, s;
      fdivmod((uint32_t)k, p.dNslot, use, s);
      const float* myraw = raw + (k % kRawDepth) * raw_floats;
      float* wall_s = ring + (size_t)s * env_floats;
      float* rock_s = wall_s + p.wall_stride;
      float* rock_sh = rock_s + R * p.rock_stride;
      int* flags = masked + s * R;
      const float lv = scaled ? __ldg(p.level + env_first + k) : 1.f;
      const float inv = scaled ? pow2_inverse(lv) : 0.f;
      // One warp waits on the mbarriers (slot drained, raw data landed); the others
      // block in the named barrier, which costs no issue slots.
      if (pt < 32) {
        if (use > 0) mbar_wait_parked(empty + s, (use - 1) & 1);
        mbar_wait_parked(rawbar + (k % kRawDepth), (k / kRawDepth) & 1);
      }
      named_bar_sync(1, kProducerThreads);
      for (int r = pt; r < R; r += kProducerThreads) flags[r] = 0;
      if (pt == 0) {
        negative[s] = 0;
        *bad16 = 0;
      }
      // Scale mode is uniform per environment: 0 none, 1 exact multiply by the
      // inverse of a power-of-two level, 2 IEEE division.
      const int mode = !scaled ? 0 : (inv != 0.f ? 1 : 2);
      const bool want16 = p.qlog2 != kNoQuantum && mode != 2;
      uint32_t* slot16 = reinterpret_cast<uint32_t*>(wall_s);
      if (want16 && lay[s] != 1) init_pads16(slot16, l16, H, W, R, h, hp, pt);
      if (!want16 && lay[s] != 0) init_padsf(wall_s, p, pt);
      named_bar_sync(1, kProducerThreads);
      bool is16 = false;
      if (want16) {
        // Optimistic 16-bit fixed-point layout; any value that is not a count in
        // [0, 2^14) sends the environment to the float layout below.
        const float inv1 = mode == 1 ? inv : 1.f;
        if (!convert_env16(p, myraw, slot16, l16, flags, inv1, p.qscale, pt)) *bad16 = 1;
        named_bar_sync(1, kProducerThreads);
        is16 = *bad16 == 0;
        if (!is16) init_padsf(wall_s, p, pt);
      }
      if (!is16) {
        uint32_t neg;
        if (mode == 0)
          neg = convert_env<0>(p, myraw, wall_s, rock_s, rock_sh, flags, lv, inv, pt);
        else if (mode == 1)
          neg = convert_env<1>(p, myraw, wall_s, rock_s, rock_sh, flags, lv, inv, pt);
        else
          neg = convert_env<2>(p, myraw, wall_s, rock_s, rock_sh, flags, lv, inv, pt);
        if ((neg >> 31) != 0u) negative[s] = 1;
      }
      named_bar_sync(1, kProducerThreads);   // slot written, raw buffer read
      if (pt == 0) {
        if (k + kRawDepth < nenv) {
          fence_proxy_async();        // generic reads of the raw buffer before its async rewrite
          issue(k + kRawDepth);
        }
        lay[s] = is16 ? 1 : 0;
        exact16[s] = is16 ? 1 : 0;
        scale_out[s] = p.qunit * (inv != 0.f ? inv : 1.f);
        st_release_shared(ready + s, k + 1);
        mbar_arrive(full + s);
        // Items of a boundary environment that belong to a neighbouring CTA.
        uint32_t missing = 0;
        if (k == 0) missing += (uint32_t)(it0 - base);
        if (k == nenv - 1) missing += (uint32_t)(base + (long long)nenv * ipe - it1);
        if (missing) mbar_arrive_count(empty + s, missing);
      }
    }
    return;
  }

  // ============================= consumer warp ================================ //
  if constexpr (kConsumers % 4 == 0 && kProducers % 4 == 0 && SRL_CONSUMER_REGS > 0)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SRL_CONSUMER_REGS));
  const int cw = warp;                            // consumer index
  float* mystage = stage + cw * (32 * T);
  const bool out_aligned = (((uintptr_t)p.out) & 15) == 0;
  for (long long q = u0 + cw; q < u1; q += kConsumers) {
    const long long pos0 = q * 32;
    const int nvalid = (int)min((long long)32, it1 - pos0);
    const bool valid = lane < nvalid;
    // local position (clamped for the idle lanes of the very last unit)
    const uint32_t lp = (uint32_t)(pos0 - base) + (uint32_t)min(lane, nvalid - 1);
    uint32_t k, rem, slot_i, rest, i, strip, use, s;
    fdivmod(lp, p.dIpe, k, rem);
    fdivmod(rem, p.dStrips, rest, strip);
    fdivmod(rest, p.dPh, slot_i, i);
    fdivmod(k, p.dNslot, use, s);
    {
      // Wait (warp-uniformly) until every environment the unit touches is in
      // its ring slot.
      const uint32_t k_lo = __shfl_sync(0xffffffffu, k, 0);
      const uint32_t k_hi = __shfl_sync(0xffffffffu, k, 31);
      for (uint32_t kk = k_lo; kk <= k_hi; ++kk) {
        uint32_t uu, ss;
        fdivmod(kk, p.dNslot, uu, ss);
        // `ready` decides; the mbarrier only parks the warp while it waits.  (Its
        // parity alone cannot tell tenant k from tenant k - 2 * nslot.)
        while (ld_acquire_shared(ready + ss) != (int)kk + 1) {
          if (mbar_try_wait_hint(full + ss, uu & 1, 100000u)) __nanosleep(200);
        }
      }
      __syncwarp();
    }
    const float* wall_s = ring + (size_t)s * env_floats;
    const float* rock_s = wall_s + p.wall_stride + slot_i * p.rock_stride;
    const float* rock_sh = rock_s + R * p.rock_stride;
    const bool any_neg = __any_sync(0xffffffffu, negative[s] != 0);
    const bool is16 = exact16[s] != 0;
    float acc[T];
    if constexpr ((T - 1) % 8 == 0 && VC % 8 == 0) {
      if (is16) {
        // 16-bit fixed-point environment: one VIADDMNMX.S16x2 per two cells.
        const uint32_t* slot16 = reinterpret_cast<const uint32_t*>(wall_s);
        const uint32_t* wa = slot16 + i * l16.wsw + strip * (S / 2);
        uint32_t acc16[T];
        sweep_item16<T, VC>(acc16, wa, wa + H * l16.wsw,
                            slot16 + 2 * H * l16.wsw + slot_i * l16.rstride, h, hp, l16.wsw);
        const float unit = scale_out[s];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int lo = (int)(short)(acc16[t] & 0xffffu), hi = (int)(short)(acc16[t] >> 16);
          const int m = max(lo, hi);
          acc[t] = m < 0 ? kNegInf : __fmul_rn((float)m, unit);
        }
      }
    }
    if (!is16) {
      if (any_neg)
        sweep_item<T, VC, 1, false>(acc, wall_s + i * Ws + strip * S, rock_s, rock_sh, h, hp,
                                    Ws);
      else
        sweep_item<T, VC, 1, true>(acc, wall_s + i * Ws + strip * S, rock_s, rock_sh, h, hp,
                                   Ws);
    }
    __syncwarp();
    // Decode the item again instead of keeping its indices live across the
    // sweep (the register file is the scarce resource there).
    {
      uint32_t lp2 = lp;
      asm volatile("" : "+r"(lp2));
      fdivmod(lp2, p.dIpe, k, rem);
      fdivmod(rem, p.dStrips, rest, strip);
      fdivmod(rest, p.dPh, slot_i, i);
      fdivmod(k, p.dNslot, use, s);
    }
    const bool floor0 = masked[s * R + slot_i] != 0;
    const int ncols = ((int)strip == p.strips - 1) ? min(T, Pw - (int)strip * S) : S;
    // Output offset of this item relative to the first item of the unit: the
    // row segments of consecutive items are contiguous in out [E,R,Ph,Pw].
    const long long goff = (long long)(env_first + k) * R * P + (long long)slot_i * P +
                           i * Pw + strip * S;
    const long long goff0 = __shfl_sync(0xffffffffu, goff, 0);
    const int so = (int)(goff - goff0);

    if (lane == 0) tma_store_wait_read();    // previous block has left mystage
    __syncwarp();
    if (valid) {
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (t < ncols) mystage[so + t] = floor0 ? fmaxf(acc[t], 0.f) : acc[t];
    }
    // Everything this warp read from the ring is in registers now: release the
    // environments.  The last lane of each environment inside the unit arrives
    // with the number of its items in the unit.
    const int total = __shfl_sync(0xffffffffu, so + ncols, nvalid - 1);
    fence_proxy_async();
    __syncwarp();
    if (valid && (lane == nvalid - 1 || rem == (uint32_t)ipe - 1))
      mbar_arrive_count(empty + s, min(rem, (uint32_t)lane) + 1);
    float* gdst = p.out + goff0;
    if (out_aligned && ((goff0 | total) & 3) == 0) {
      if (lane == 0) {
        tma_store_1d(gdst, mystage, (uint32_t)total * 4);
        tma_store_commit();
      }
    } else {
      for (int t = lane; t < total; t += 32) __stcs(gdst + t, mystage[t]);
      __syncwarp();
    }
  }
  if (lane == 0) tma_store_wait_all();
}

// --------------------------------------------------------------------------- //
// Direct kernel (big walls): one CTA per (environment group, rotation chunk).
// --------------------------------------------------------------------------- //
template <int T, int VC, int PAIRED>
__global__ void __launch_bounds__(288, 2)
maxplus_direct_kernel(const MaxPlusParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  int* masked = reinterpret_cast<int*>(smem_raw + 16);
  const int flag_bytes = round_up(p.G * p.RC * 4, 16);
  float* wall_s = reinterpret_cast<float*>(smem_raw + 16 + flag_bytes);
  float* rock_s = wall_s + p.G * p.wall_stride;
  float* rock_sh = rock_s + p.G * p.RC * p.rock_stride;
  constexpr int S = T - 1;

  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  // blockIdx = ((group * rchunks) + rchunk) * nbands + band.  Bands (p.nbands > 1
  // only with one environment per CTA) cut the output rows so that a single
  // observation still spreads over the chip.
  const int bandi = blockIdx.x % p.nbands;
  const int gc = blockIdx.x / p.nbands;
  const int group = gc / p.rchunks;
  const int rchunk = gc % p.rchunks;
  const int i0 = bandi * p.band;                            // first output row
  const int rows_out = min(p.band, p.Ph - i0);
  const int Hb = rows_out + p.h - 1;                        // wall rows this CTA needs
  const int e0 = group * p.G;
  const int r0 = rchunk * p.RC;
  const int Gv = min(p.G, p.E - e0);
  const int RCv = min(p.RC, p.R - r0);
  const int H = p.H, W = p.W, h = p.h, hp = p.hp, Ws = p.Ws;
  // Row r of environment el of the group: local row el * Hb + r, global row
  // (e0 + el) * H + i0 + r.
  const FastDiv dHb = {Hb == 1 ? 0u : (uint32_t)(((1ull << 32) + Hb - 1) / Hb), (uint32_t)Hb};
  const FastDiv dRCv = {RCv == 1 ? 0u : (uint32_t)(((1ull << 32) + RCv - 1) / RCv),
                        (uint32_t)RCv};

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  for (int k = tid; k < p.G * p.RC; k += nthreads) masked[k] = 0;
  __syncthreads();

  // ---- bulk TMA loads (warp 0) + padding fills (everyone) -------------------- //
  const bool any_tma = p.tma_wall || p.tma_rock;
  if (tid < 32 && any_tma) {
    if (tid == 0) {
      uint32_t bytes = 0;
      if (p.tma_wall) bytes += (uint32_t)Gv * Hb * W * 4;
      if (p.tma_rock) bytes += (uint32_t)Gv * RCv * h * h * 4;
      mbar_arrive_expect_tx(bar, bytes);
    }
    __syncwarp();
    if (p.tma_wall) {
      for (uint32_t k = tid; k < (uint32_t)(Gv * Hb); k += 32) {
        uint32_t el, r;
        fdivmod(k, dHb, el, r);
        tma_load_1d(wall_s + (el * H + r) * Ws,
                    p.walls + ((size_t)(e0 + el) * H + i0 + r) * W, W * 4, bar);
      }
    }
    if (p.tma_rock) {
      for (uint32_t k = tid; k < (uint32_t)(Gv * RCv * h); k += 32) {
        uint32_t er, u, el, r;
        fdivmod(k, p.dh, er, u);
        fdivmod(er, dRCv, el, r);
        tma_load_1d(rock_s + (el * p.RC + r) * p.rock_stride + u * hp,
                    p.rocks + (((size_t)(e0 + el) * p.R + r0 + r) * h + u) * h,
                    h * 4, bar);
      }
    }
  }
  if (!p.tma_wall) {
    for (uint32_t k = tid; k < (uint32_t)(Gv * Hb * W); k += nthreads) {
      uint32_t row, c, el, r;
      fdivmod(k, p.dW, row, c);
      fdivmod(row, dHb, el, r);
      wall_s[(el * H + r) * Ws + c] =
          __ldg(p.walls + ((size_t)(e0 + el) * H + i0 + r) * W + c);
    }
  }
  if (!p.tma_rock) {
    for (uint32_t k = tid; k < (uint32_t)(Gv * RCv * h * h); k += nthreads) {
      uint32_t rrow, c, er, u, el, r;
      fdivmod(k, p.dh, rrow, c);
      fdivmod(rrow, p.dh, er, u);
      fdivmod(er, dRCv, el, r);
      rock_s[(el * p.RC + r) * p.rock_stride + u * hp + c] =
          __ldg(p.rocks + (((size_t)(e0 + el) * p.R + r0 + r) * h + u) * h + c);
    }
  }
  // Wall pad columns [W, Ws) must be finite.  (Rock pad columns [h, hp) and the
  // shifted copy are written by the normalise pass below.)
  {
    const int padw = Ws - W;
    for (int k = tid; k < Gv * Hb * padw; k += nthreads) {
      const int row = k / padw, el = row / Hb, r = row - el * Hb;
      wall_s[(el * H + r) * Ws + W + k % padw] = 0.f;
    }
  }
  if (any_tma) mbar_wait(bar, 0);
  __syncthreads();

  // ---- normalise in place, fold the mask, build the shifted copy ------------- //
  const bool scaled = p.level != nullptr;
  if (scaled) {
    for (uint32_t k = tid; k < (uint32_t)(Gv * Hb * W); k += nthreads) {
      uint32_t row, c, el, r;
      fdivmod(k, p.dW, row, c);
      fdivmod(row, dHb, el, r);
      float* q = wall_s + (el * H + r) * Ws + c;
      *q = div_level(*q, __ldg(p.level + e0 + el));
    }
  }
  for (uint32_t k = tid; k < (uint32_t)(Gv * RCv * h * hp); k += nthreads) {
    uint32_t rrow, c, er, u, el, r;
    fdivmod(k, p.dhp, rrow, c);
    fdivmod(rrow, p.dh, er, u);
    fdivmod(er, dRCv, el, r);
    const int slot = el * p.RC + r;
    float* q = rock_s + slot * p.rock_stride + u * hp + c;
    float n = kNegInf;
    if ((int)c < h) {
      bool dead = false;
      n = prep_rock(*q, scaled, scaled ? __ldg(p.level + e0 + el) : 1.f, 0.f, p.threshold,
                    dead);
      if (dead) masked[slot] = 1;
    }
    *q = n;
    if constexpr (PAIRED != 0) {
      float* sh = rock_sh + slot * p.rock_stride + u * hp + c;
      if (c != 0) sh[-1] = n;
      if ((int)c == hp - 1) sh[0] = kNegInf;
    }
  }
  __syncthreads();

  // ---- sweep ------------------------------------------------------------------ //
  const int Ph = p.Ph, Pw = p.Pw;
  // p.dPh divides by the band height (== Ph without banding); the last band may be
  // shorter.
  const int items = Gv * RCv * p.strips * p.band;
  for (int item = tid; item < items; item += nthreads) {
    uint32_t rest, i, strip, er, el, r;
    fdivmod((uint32_t)item, p.dPh, rest, i);
    if ((int)i >= rows_out) continue;
    fdivmod(rest, p.dStrips, er, strip);
    fdivmod(er, dRCv, el, r);
    const int slot = el * p.RC + r;
    float acc[T];
    sweep_item<T, VC, PAIRED>(acc, wall_s + el * p.wall_stride + i * Ws + strip * S,
                              rock_s + slot * p.rock_stride,
                              rock_sh + slot * p.rock_stride, h, hp, Ws);
    const bool floor0 = masked[slot] != 0;
    float* orow = p.out +
                  (((size_t)(e0 + el) * p.R + r0 + r) * Ph + i0 + i) * (size_t)Pw +
                  strip * S;
    // The last column of a strip is the first of the next one; only the last
    // strip stores it.
    const int ncols = ((int)strip == p.strips - 1) ? min(T, Pw - (int)strip * S) : S;
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (t < ncols) __stcs(orow + t, floor0 ? fmaxf(acc[t], 0.f) : acc[t]);
  }
}

// --------------------------------------------------------------------------- //
// Host-side dispatch.
// --------------------------------------------------------------------------- //
namespace {

template <int T, int VC>
int launch(const MaxPlusParams& p, bool staged, int blocks, int threads, size_t smem,
           int paired, cudaStream_t stream) {
#define SRL_LAUNCH(...)                                                              \
  do {                                                                               \
    auto k = __VA_ARGS__;                                                            \
    SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                  (int)smem));                                       \
    k<<<blocks, threads, smem, stream>>>(p);                                         \
  } while (0)
  if (staged) {
    if (paired == 0) SRL_LAUNCH(maxplus_staged_kernel<T, VC, 0>);
    else SRL_LAUNCH(maxplus_staged_kernel<T, VC, 1>);
  } else {
    if (paired == 0) SRL_LAUNCH(maxplus_direct_kernel<T, VC, 0>);
    else SRL_LAUNCH(maxplus_direct_kernel<T, VC, 1>);
  }
#undef SRL_LAUNCH
  return check_launch("maxplus_f32 kernel");
}

}  // namespace

constexpr int kStreamNotTaken = 1;   // stream_only: the call does not fit the stream kernel

static int maxplus_f32_impl(const float* walls, const float* rocks, const float* level,
                            float* out, int E, int R, int H, int W, int h, float threshold,
                            int variant, int quantum_log2, int forced_T,
                            cudaStream_t stream, bool stream_only = false) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h, SRL_E_INVALID,
              "maxplus_f32: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && rocks && out, SRL_E_INVALID, "maxplus_f32: null pointer");
  SRL_REQUIRE(H <= 4096 && W <= 4096 && R <= 4096, SRL_E_UNSUPPORTED,
              "maxplus_f32: dimension above 4096");

  MaxPlusParams p;
  p.walls = walls; p.rocks = rocks; p.level = level; p.out = out;
  p.E = E; p.R = R; p.H = H; p.W = W; p.h = h;
  p.Ph = H - h + 1; p.Pw = W - h + 1;
  p.threshold = threshold;
  p.band = p.Ph;
  p.nbands = 1;
  const int P = p.Ph * p.Pw;

  const int sms = sm_count();
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "maxplus_f32: no CUDA device");
  Choice c = choose_tile(p.Pw, h);
  // Latency-bound calls (one observation, a handful of maps): far fewer items
  // than lanes on the chip, so take the narrowest tile -- more, shorter threads.
  if ((long long)E * R * strips_for(p.Pw, c.T) * p.Ph < (long long)sms * 256) c.T = kTs[0];
  bool tile_fixed = forced_T != 0;
  if (forced_T) c.T = forced_T;
  if (const char* st = getenv("SRL_MP_T")) {   // test / tuning override
    const int t = atoi(st);
    for (int known : kTs)
      if (t == known) {
        c.T = t;
        tile_fixed = true;
      }
  }
  // The widest tile with 16-column chunks is outside choose_tile's register budget (the
  // staged / direct kernels), but the stream kernel's consumers hold it: where it wastes fewer
  // columns (49 output columns are two strips of 25 instead of three of 17: config 4, 1.60 ->
  // 1.56 ms) the stream kernel is tried with it first.
  if (!tile_fixed && !stream_only && c.VC == 16 && c.T != 25) {
    auto cost_of = [&](int TT) {
      const double waste = (double)strips_for(p.Pw, TT) * TT / p.Pw;
      const double loads = ((TT + c.VC + 2) / 4 + c.VC / 2) / (double)(TT * c.VC);
      return waste * (1.03 + loads);
    };
    if (cost_of(25) < 0.985 * cost_of(c.T) &&
        (long long)E * R * strips_for(p.Pw, 25) * p.Ph >= (long long)sms * 256) {
      const int rc = maxplus_f32_impl(walls, rocks, level, out, E, R, H, W, h, threshold,
                                      variant, quantum_log2, 25, stream, true);
      if (rc != kStreamNotTaken) return rc;
    }
  }
  const int T = c.T, VC = c.VC;
  const int paired = variant != 0;   // 0: FADD + FMNMX, 1: FADD2 + FMNMX3 (default)
  p.hp = round_up(h, VC);
  p.strips = strips_for(p.Pw, T);
  // Columns a thread may touch: strip start + (hp - VC) + 4*NR4 floats.
  const int nr4 = (T + VC + 2) / 4;
  const int need = (p.strips - 1) * (T - 1) + (p.hp - VC) + 4 * nr4;
  p.Ws = round_up(need > W ? need : W, 4);
  if ((p.Ws / 4) % 2 == 0) p.Ws += 4;
  p.wall_stride = H * p.Ws;
  p.rock_stride = h * p.hp + 4;
  p.tma_wall = (W % 4 == 0) && (((uintptr_t)walls) % 16 == 0);
  p.tma_rock = (h % 4 == 0) && (((uintptr_t)rocks) % 16 == 0);

  const size_t kBudget = 110 * 1024;     // per CTA, two CTAs per SM
  const size_t kMax = 220 * 1024;
  const size_t wall_bytes = (size_t)p.wall_stride * 4;
  const size_t rock_bytes = (size_t)p.rock_stride * 4 * 2;
  const int kThreads = 288;

  // ---- staged persistent kernel: whole environments (all R rotations) -------- //
  auto staged_smem = [&](int G, bool stage_out) {
    return 16 + (size_t)round_up(2 * G * R * 4, 16) +
           (size_t)G * H * W * 4 + (size_t)G * R * h * h * 4 +     // raw staging
           G * wall_bytes + (size_t)G * R * rock_bytes +           // compute layout
           (stage_out ? (size_t)G * R * P * 4 : 0);
  };
  auto staged_fits = [&](int G, bool stage_out, size_t cap) {
    return staged_smem(G, stage_out) <= cap && (size_t)G * H * W < 65536 &&
           (size_t)G * R * h * p.hp < 65536 && (size_t)G * R * P < 65536 * 4;
  };
  const bool can_stage_out = ((size_t)R * P) % 4 == 0 && ((uintptr_t)out) % 16 == 0;
  bool staged = p.tma_wall && p.tma_rock && staged_fits(1, can_stage_out, kBudget);
  if (const char* s = getenv("SRL_MP_MODE")) {
    if (atoi(s) == 0) staged = false;
  }

  // ---- stream kernel: warp-specialised, one persistent CTA per SM -------------- //
  {
    const int T1 = T;
    p.ipe = R * p.Ph * p.strips;
    p.items = (long long)E * p.ipe;
    p.units = (int)((p.items + 31) / 32);
    const size_t raw_b = ((size_t)H * W + (size_t)R * h * h) * 4;
    const size_t env_b = wall_bytes + (size_t)R * rock_bytes;
    const size_t stage_b = (size_t)kConsumers * 32 * T1 * 4;
    auto stream_smem = [&](int ns) {
      return (size_t)round_up((2 * ns + kRawDepth) * 8 + (5 * ns + 1 + ns * R) * 4, 16) +
             kRawDepth * raw_b + ns * env_b + stage_b;
    };
    const size_t kStreamMax = 227 * 1024;
    int ns = 0;
    while (ns < 24 && stream_smem(ns + 1) <= kStreamMax) ++ns;
    if (const char* sn = getenv("SRL_MP_NSLOT")) {   // tuning override
      const int v = atoi(sn);
      if (v >= 3 && v < ns) ns = v;
    }
    int mode = 2;
    if (const char* sm = getenv("SRL_MP_MODE")) mode = atoi(sm);
    const int blocks_s = p.units < sms ? p.units : sms;
    // index spaces of the fast divisions (n * d < 2^32) and the TMA tx count
    const double per_cta = (double)p.items / blocks_s + 2.0 * p.ipe + 64;
    // A consumer warp waits until every environment its 32-item unit touches is
    // resident, and a ring slot is recycled only when all `ipe` items of its tenant
    // are done: a unit that spans more environments than the ring has slots would
    // wait for a slot that only its own completion can free.  32 consecutive items
    // touch at most (30 + ipe) / ipe + 1 environments.
    const int unit_span = (30 + p.ipe) / p.ipe + 1;
    const bool ok = paired && p.tma_wall && p.tma_rock && ns >= 3 && mode == 2 &&
                    unit_span <= ns &&
                    raw_b < (1u << 20) && per_cta * p.ipe < 4.0e9 && p.ipe < (1 << 20) &&
                    (size_t)H * W < 65536 && (size_t)R * h * p.hp < 65536 &&
                    p.items < (1ll << 40);
    if (ok) {
      // 16-bit fixed-point sweep: needs the quantum hint, 16-byte aligned strips
      // and a packed layout that fits the float layout's slot.
      p.qlog2 = kNoQuantum;
      p.need16 = 4;
      if (quantum_log2 != kNoQuantum && quantum_log2 > -120 && quantum_log2 < 120 &&
          (T - 1) % 8 == 0 && VC % 8 == 0) {
        const int pv = VC / 2;
        const int na4 = ((T + 1) / 2 + pv - 1 + 3) / 4, nb4 = ((T - 1) / 2 + pv - 1 + 3) / 4;
        int need = ((p.strips - 1) * (T - 1) + (p.hp - VC)) / 2 + 4 * (na4 > nb4 ? na4 : nb4);
        if (need < W / 2 + 1) need = W / 2 + 1;
        const Layout16 l = layout16(H, R, h, p.hp, need);
        if ((size_t)l.words * 4 <= env_b) {
          p.qlog2 = quantum_log2;
          p.need16 = need;
          p.qscale = ldexpf(1.f, -quantum_log2);
          p.qunit = ldexpf(1.f, quantum_log2);
        }
      }
      p.nslot = ns;
      p.G = 1; p.RC = R; p.rchunks = 1; p.stage_out = 1; p.ngroups = E;
      p.dPh = make_fastdiv(p.Ph); p.dStrips = make_fastdiv(p.strips);
      p.dRC = make_fastdiv(R); p.dW = make_fastdiv(W); p.dH = make_fastdiv(H);
      p.dh = make_fastdiv(h); p.dhp = make_fastdiv(p.hp);
      p.dW4 = make_fastdiv(W / 4); p.dh4 = make_fastdiv(h / 4);
      p.dIpe = make_fastdiv(p.ipe); p.dNslot = make_fastdiv(ns);
      const size_t smem_s = stream_smem(ns);
#define SRL_MP_CASE(TT, VV)                                                          \
  if (T == TT && VC == VV) {                                                         \
    auto k = maxplus_stream_kernel<TT, VV>;                                          \
    SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                  (int)smem_s));                                     \
    k<<<blocks_s, kStreamThreads, smem_s, stream>>>(p);                              \
    return check_launch("maxplus_stream_kernel");                                    \
  }
#define SRL_MP_ROW(VV)                                                            \
  SRL_MP_CASE(5, VV) SRL_MP_CASE(9, VV) SRL_MP_CASE(13, VV) SRL_MP_CASE(17, VV)   \
  SRL_MP_CASE(21, VV) SRL_MP_CASE(25, VV)
      SRL_MP_ROW(4)
      SRL_MP_ROW(8)
      SRL_MP_ROW(16)
#undef SRL_MP_ROW
#undef SRL_MP_CASE
    }
  }
  if (stream_only) return kStreamNotTaken;

  int G = 1, RC = R, blocks, threads;
  size_t smem;
  if (staged) {
    const int items_per_env = R * p.strips * p.Ph;
    while (G < 32 && G < E && staged_fits(G + 1, can_stage_out, kBudget) &&
           (G + 1) * items_per_env <= kThreads)
      ++G;
    if (const char* s = getenv("SRL_MP_G")) {
      const int g = atoi(s);
      if (g >= 1 && staged_fits(g, can_stage_out, kMax)) G = g;
    }
    p.stage_out = can_stage_out;
    smem = staged_smem(G, can_stage_out);
    threads = pick_threads(G * items_per_env, kThreads);
    p.ngroups = (E + G - 1) / G;
    const int per_sm = smem <= kBudget ? 2 : 1;
    blocks = p.ngroups < sms * per_sm ? p.ngroups : sms * per_sm;
  } else {
    auto direct_smem = [&](int g, int rc) {
      return 16 + (size_t)round_up(g * rc * 4, 16) + g * wall_bytes +
             (size_t)g * rc * rock_bytes;
    };
    while (RC > 1 && (direct_smem(1, RC) > kBudget ||
                      (size_t)RC * p.strips * p.Ph >= (1u << 20) ||
                      (size_t)RC * h * p.hp >= 65536))
      RC = (RC + 1) / 2;
    SRL_REQUIRE(direct_smem(1, RC) <= kMax, SRL_E_UNSUPPORTED,
                "maxplus_f32: one wall (%dx%d) + one rock (%d) exceed shared memory",
                H, W, h);
    // One CTA sweeps its items in whole passes of its threads: with few items per
    // CTA (one rotation of a 97-row map is 582 items at T = 17) the tile width that
    // fills the passes best wins over the one with the fewest wasted columns.
    if (!tile_fixed && (long long)E * R * p.strips * p.Ph >= (long long)sms * 256) {
      auto cost_of = [&](int TT) {
        const int st = strips_for(p.Pw, TT);
        const int items = RC * st * p.Ph;
        const int th = pick_threads(items, kThreads);
        const int passes = (items + th - 1) / th;
        const double util = (double)items / ((double)passes * th);
        const double waste = (double)st * TT / p.Pw;
        const double loads = ((TT + VC + 2) / 4 + VC / 2) / (double)(TT * VC);
        return waste * (1.03 + loads) / util;
      };
      int best_T = T;
      double best_cost = cost_of(T) * 0.97;       // switch only for a clear gain
      for (int TT : kTs) {
        if (VC == 16 && TT > 21) continue;
        const double cst = cost_of(TT);
        if (cst < best_cost) {
          best_cost = cst;
          best_T = TT;
        }
      }
      if (best_T != T)
        return maxplus_f32_impl(walls, rocks, level, out, E, R, H, W, h, threshold, variant,
                                quantum_log2, best_T, stream);
    }
    const int items_per_env = RC * p.strips * p.Ph;
    if (RC == R) {
      while (G < 32 && G < E && direct_smem(G + 1, RC) <= kBudget &&
             (G + 1) * items_per_env <= kThreads &&
             (size_t)(G + 1) * H * W < 65536 && (size_t)(G + 1) * RC * h * p.hp < 65536)
        ++G;
    }
    p.stage_out = 0;
    smem = direct_smem(G, RC);
    p.ngroups = (E + G - 1) / G;
    blocks = p.ngroups * ((R + RC - 1) / RC);
    // Few CTAs (a single observation, a handful of maps): cut the output rows
    // into bands so that the work spreads over the SMs.
    if (G == 1 && blocks < 2 * sms && p.Ph > 1) {
      int nb = (2 * sms + blocks - 1) / blocks;
      if (nb > p.Ph) nb = p.Ph;
      p.band = (p.Ph + nb - 1) / nb;
      p.nbands = (p.Ph + p.band - 1) / p.band;
    }
    threads = pick_threads(G * RC * p.strips * p.band, kThreads);
    blocks *= p.nbands;
  }
  if (const char* s = getenv("SRL_MP_THREADS")) {
    const int t = atoi(s);
    if (t >= 32 && t <= 288 && t % 32 == 0) threads = t;
  }
  p.G = G; p.RC = RC; p.rchunks = (R + RC - 1) / RC;
  p.dPh = make_fastdiv(staged ? p.Ph : p.band); p.dStrips = make_fastdiv(p.strips);
  p.dRC = make_fastdiv(RC); p.dW = make_fastdiv(W); p.dH = make_fastdiv(H);
  p.dh = make_fastdiv(h); p.dhp = make_fastdiv(p.hp);
  p.dW4 = make_fastdiv(W >= 4 ? W / 4 : 1); p.dh4 = make_fastdiv(h >= 4 ? h / 4 : 1);

#define SRL_MP_CASE(TT, VV)                                                     \
  if (T == TT && VC == VV)                                                      \
    return launch<TT, VV>(p, staged, blocks, threads, smem, paired, stream);
#define SRL_MP_ROW(VV)                                                            \
  SRL_MP_CASE(5, VV) SRL_MP_CASE(9, VV) SRL_MP_CASE(13, VV) SRL_MP_CASE(17, VV)   \
  SRL_MP_CASE(21, VV) SRL_MP_CASE(25, VV)
  SRL_MP_ROW(4)
  SRL_MP_ROW(8)
  SRL_MP_ROW(16)
#undef SRL_MP_ROW
#undef SRL_MP_CASE
  return fail(SRL_E_UNSUPPORTED, "maxplus_f32: no kernel for T=%d VC=%d", T, VC);
}

int maxplus_f32(const float* walls, const float* rocks, const float* level,
                float* out, int E, int R, int H, int W, int h, float threshold,
                int variant, int quantum_log2, cudaStream_t stream) {
  return maxplus_f32_impl(walls, rocks, level, out, E, R, H, W, h, threshold, variant,
                          quantum_log2, 0, stream);
}

}  // namespace srl
