// Siamese correlation layer on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Reference: stackrl.nets.correlation (stackrl/nets/layers.py:21-38; used by
// PseudoSiamFCN / DeepQSiamFCN, nets/models.py:89, 182), per sample
//   out[i, j] = sum_{u, v, c} x[i+u, j+v, c] * f[u, v, c]        (channels-last)
// = tf.nn.conv2d(x[None], f[..., None], 'VALID'): a convolution with ONE output
// channel, so a direct implicit GEMM has N = 1.  The contraction is made dense by
// splitting it in two:
//
//   G[(r, j), u] = sum_{v, c} x[r, j+v, c] * f[u, v, c]     (GEMM, tensor cores)
//   out[i, j]    = sum_u G[(i+u, j), u]                      (diagonal sum, epilogue)
//
// M = the pixels (r, j) of one image row (tile of 128 pixels j), N = the filter rows
// u, K = (v, c) = one filter row.  The A operand of tap v is the SAME image row
// shifted by v pixels, so one row staged in shared memory serves all taps: the
// shared-memory matrix descriptor's start address just moves by v rows.  That needs
// rows that are 16 bytes apart (the no-swizzle K-major core-matrix layout), hence
// the row is staged "chunk planar": plane q holds channels [4q, 4q+4) of every
// pixel, 16 bytes per pixel.  The filter is staged the same way (plane (v, q):
// 16 bytes per filter row u) and stays resident for the whole sample.
//
// Precision: the reference layer computes in float32.  Every operand is split into a
// hi and a lo part whose sum carries 21-22 significant bits, and three tensor-core
// products hi*hi + hi*lo + lo*hi are accumulated in float32; a TMEM accumulator only
// sums the (v, c) of one filter row (K = w*C), the 32-term sum over u runs in the
// epilogue in float32.  Two operand formats:
//   * FP16 pairs (default when C is a multiple of 16): hi = fp16(s*x), lo = fp16(s*x -
//     hi), with s a power of two per CTA that puts the largest |x| of the rows the CTA
//     reads (the largest |f| of the filter) just below 2^14 -- no overflow, and what
//     underflows is below 2^-39 of the largest value; the output is un-scaled by exact
//     powers of two.  16 K elements per MMA instead of 8: half the MMAs.
//   * TF32 pairs (C a multiple of 8, or SRL_SIAM_TC=tf32): hi = the value with the low
//     13 mantissa bits cleared, lo = value - hi (exact), no scaling.
// Both operands come from shared memory and N is only the filter height, so the tensor
// core's operand reads (128 x 32 B of image per MMA), not its arithmetic, bound the
// kernel (ncu: tensor pipe 22 % busy, its shared-memory wavefronts at 67 % of peak).  The
// two products that share the image's hi part are therefore ONE MMA: the filter's hi and
// lo planes are stacked along N (columns [0, N) = hi*hi, [N, 2N) = hi*lo), the image's lo
// part multiplies the hi planes alone into its own columns [64, 64 + N), and the epilogue
// adds the three column groups: 11 KB of operand reads per K step instead of 15.
//
// Warp roles (288 threads, one CTA = one band of output rows of one sample):
//   warp 0        allocates TMEM, then one elected lane issues every tcgen05.mma
//   warps 1..4    epilogue: tcgen05.ld of a finished row tile, diagonal accumulation into
//                 a 32-row ring in shared memory, finished output rows to global memory
//   warps 5..8    stage image rows: coalesced loads -> hi / lo planes, up to three rows
//                 ahead of the MMAs
// (staging and epilogue in the same warps put load latency + epilogue on one critical
// path per row: 3.7 us against 2.5 us of MMAs.)
// Synchronisation is mbarriers only (stage full / stage free via tcgen05.commit /
// accumulator full / accumulator empty); the two TMEM accumulators alternate.
#include <algorithm>
#include <cstdlib>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

constexpr int kTcThreads = 288;
constexpr int kTcWorkers = 128;       // warps 1..4 (epilogue) and warps 5..8 (staging)
constexpr int kTcStages = 3;          // image rows in flight
constexpr int kTileM = 128;           // pixels per tile (UMMA M)

struct SiamTcParams {
  const float* x;     // [B, H, W, C]
  const float* f;     // [B, h, wd, C]
  float* out;         // [B, Ph, Pw]
  int H, W, C, h, wd, Ph, Pw;
  int Q;              // 16-byte chunks per pixel: C / 4 (TF32) or C / 8 (FP16)
  int N;              // filter rows rounded up to a multiple of 16 (UMMA N)
  int npx;            // pixels staged per row tile: kTileM + wd - 1
  int plane;          // bytes between consecutive planes of a staged row (16 * odd)
  int bplane;         // bytes of one filter plane: 16 * N
  int bands, band;    // output-row bands per sample, rows per band
  int jblocks;        // 128-pixel column blocks
};

// ---- tcgen05 / TMEM wrappers (PTX ISA 8.6+, sm_100a) --------------------------------- //
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 or FP16 operands, float32 accumulate.
template <bool kHalf>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                     uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                     uint32_t accumulate) {
  // executed by the whole (converged) warp with warp-uniform operands; one elected
  // lane issues -- the operands then stay on the uniform datapath
  if (kHalf) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// The two MMAs of one K step under ONE election: D1 (+)= A1 * B with descriptor i1,
// D2 (+)= A2 * B with descriptor i2 (A1, A2 share the high descriptor word).
template <bool kHalf>
__device__ __forceinline__ void umma_pair(uint32_t d1, uint32_t d2, uint32_t a1_lo, uint32_t a2_lo,
                                          uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t i1,
                                          uint32_t i2, uint32_t accumulate) {
#define SRL_UMMA_PAIR(KIND)                                                              \
  asm volatile(                                                                          \
      "{\n\t"                                                                            \
      ".reg .pred p, e;\n\t"                                                             \
      ".reg .b64 da1, da2, db;\n\t"                                                      \
      "mov.b64 da1, {%2, %4};\n\t"                                                       \
      "mov.b64 da2, {%3, %4};\n\t"                                                       \
      "mov.b64 db, {%5, %6};\n\t"                                                        \
      "elect.sync _|e, 0xffffffff;\n\t"                                                  \
      "setp.ne.b32 p, %9, 0;\n\t"                                                        \
      "@e tcgen05.mma.cta_group::1.kind::" KIND " [%0], da1, db, %7, p;\n\t"             \
      "@e tcgen05.mma.cta_group::1.kind::" KIND " [%1], da2, db, %8, p;\n\t"             \
      "}\n" ::"r"(d1),                                                                   \
      "r"(d2), "r"(a1_lo), "r"(a2_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(i1), "r"(i2), \
      "r"(accumulate)                                                                    \
      : "memory")
  if (kHalf) {
    SRL_UMMA_PAIR("f16");
  } else {
    SRL_UMMA_PAIR("tf32");
  }
#undef SRL_UMMA_PAIR
}
// mbarrier arrive once every tcgen05.mma issued so far by the elected lane has completed
// (elect.sync picks the same lane every time for a full mask).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

// Shared-memory matrix descriptor, no swizzle, K-major: rows of a core matrix are 16
// bytes apart, `sbo` bytes between 8-row groups, `lbo` bytes between the 16-byte
// chunks along K (cute::UMMA::SmemDescriptor, version 1 = sm_100).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 at bit 4), A and B
// formats at bits 7 and 10 (F16 = 0, TF32 = 2), both K-major, N >> 3 at bit 17, M >> 4
// at bit 24.
template <bool kHalf>
__device__ __forceinline__ uint32_t instr_desc(int n) {
  const uint32_t fmt = kHalf ? 0u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

// One 16-byte K chunk of an operand row: 4 channels as TF32 pairs or 8 channels as FP16
// pairs (scaled by the power of two `s`).  `src` is 16-byte aligned.
template <bool kHalf>
struct Chunk {
  float4 a, b;          // b: channels 4..7 of the FP16 chunk
  __device__ __forceinline__ void zero() {
    a = make_float4(0.f, 0.f, 0.f, 0.f);
    b = a;
  }
  __device__ __forceinline__ void load(const float* src) {
    a = __ldg(reinterpret_cast<const float4*>(src));
    if (kHalf) b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  }
  __device__ __forceinline__ void split(float s, uint4& hi, uint4& lo) const {
    if (kHalf) {
      const float v[8] = {a.x * s, a.y * s, a.z * s, a.w * s, b.x * s, b.y * s, b.z * s, b.w * s};
      uint32_t h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __half2 hh = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * k] - back.x, v[2 * k + 1] - back.y);
        h[k] = *reinterpret_cast<const uint32_t*>(&hh);
        l[k] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      hi = make_uint4(h[0], h[1], h[2], h[3]);
      lo = make_uint4(l[0], l[1], l[2], l[3]);
    } else {
      const float4 th = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
      hi = make_uint4(__float_as_uint(th.x), __float_as_uint(th.y), __float_as_uint(th.z),
                      __float_as_uint(th.w));
      lo = make_uint4(__float_as_uint(a.x - th.x), __float_as_uint(a.y - th.y),
                      __float_as_uint(a.z - th.z), __float_as_uint(a.w - th.w));
    }
  }
};

// Largest |value| of n floats (n a multiple of 4, 16-byte aligned), over the CTA; NaNs
// are ignored.  `red` is shared scratch of one float per warp.
__device__ __forceinline__ float cta_abs_max(const float* src, int n, float* red) {
  float m = 0.f;
  for (int k = threadIdx.x; k < n / 4; k += kTcThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + k);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < kTcThreads / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  return m;
}
// The power of two that puts `m` into [2^13, 2^14) (1 for zero / non-finite input; the
// exponent is kept within +-100 so that scale and inverse are normal numbers).
__device__ __forceinline__ float pow2_scale(float m, float& inverse) {
  int e = 0;
  if (m > 0.f && m < 3.0e38f) e = min(max(13 - ilogbf(m), -100), 100);
  inverse = __uint_as_float((uint32_t)(127 - e) << 23);
  return __uint_as_float((uint32_t)(127 + e) << 23);
}

template <bool kHalf>
__global__ void __launch_bounds__(kTcThreads, 1) siam_tc_kernel(const SiamTcParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  // (shuffled so that the compiler knows the warp index is warp-uniform: the role
  // branches and everything the MMA issuer computes then stay on the uniform datapath)
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int Q = p.Q, N = p.N, wd = p.wd;
  // ---- shared memory ----------------------------------------------------------------- //
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* full = bars;                        // [kTcStages] stage staged by the workers
  uint64_t* freeb = bars + kTcStages;           // [kTcStages] stage consumed by the MMAs
  uint64_t* acc_full = bars + 2 * kTcStages;    // [2] accumulator written
  uint64_t* acc_empty = acc_full + 2;           // [2] accumulator read back
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);        // [16] reduction scratch
  constexpr int kCw = kHalf ? 8 : 4;                           // channels per 16-byte chunk
  unsigned char* at = smem_raw + 256;
  unsigned char* bmat = at;                                   // [wd * Q planes][hi, lo][N * 16 B]
  at += (size_t)2 * wd * Q * p.bplane;
  unsigned char* amat = at;                                   // [stages][2 parts][Q planes]
  const int stage_bytes = 2 * Q * p.plane;
  at += (size_t)kTcStages * stage_bytes;
  float* ring = reinterpret_cast<float*>(at);                 // [32][kTileM] partial output rows

  const int sample = blockIdx.x / (p.bands * p.jblocks);
  const int rest = blockIdx.x - sample * (p.bands * p.jblocks);
  const int bandi = rest / p.jblocks, jb = rest - bandi * p.jblocks;
  const int i0 = bandi * p.band, i1 = min(i0 + p.band, p.Ph);     // output rows of this CTA
  const int j0 = jb * kTileM;                                      // first output column
  const int r0 = i0, r1 = i1 + p.h - 1;                           // image rows needed
  const int nrows = r1 - r0;
  const float* xs = p.x + (size_t)sample * p.H * p.W * p.C;
  const float* fs = p.f + (size_t)sample * p.h * wd * p.C;

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(full + s, kTcWorkers);
      mbar_init(freeb + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full + b, 1);
      mbar_init(acc_empty + b, kTcWorkers);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  // ---- filter planes, hi and lo (all threads; resident for the whole CTA) ------------- //
  // plane (v, q) holds f[u, v, 4q .. 4q+4) for u = 0 .. N-1 (zero rows beyond h), its lo
  // part right behind its hi part: together they are one K-major operand of 2N rows
  float sx = 1.f, sf = 1.f, inv_x = 1.f, inv_f = 1.f;
  if (kHalf) {
    sx = pow2_scale(cta_abs_max(xs + (size_t)r0 * p.W * p.C, nrows * p.W * p.C, red), inv_x);
    sf = pow2_scale(cta_abs_max(fs, p.h * wd * p.C, red), inv_f);
  }
  for (int k = tid; k < wd * Q * N; k += kTcThreads) {
    const int u = k % N, vq = k / N;
    const int v = vq / Q, q = vq - v * Q;
    Chunk<kHalf> val;
    val.zero();
    if (u < p.h) val.load(fs + ((size_t)u * wd + v) * p.C + kCw * q);
    uint4 hi, lo;
    val.split(sf, hi, lo);
    *reinterpret_cast<uint4*>(bmat + (size_t)(2 * vq) * p.bplane + 16 * u) = hi;
    *reinterpret_cast<uint4*>(bmat + (size_t)(2 * vq + 1) * p.bplane + 16 * u) = lo;
  }
  for (int k = tid; k < 32 * kTileM; k += kTcThreads) ring[k] = 0.f;
  fence_proxy_async();            // generic-proxy writes of the filter -> tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // =============================== MMA issuer ========================================= //
    // every lane follows the barriers (cheap, keeps the warp converged for the final
    // block barrier); one lane issues
    // Everything an MMA names must sit in uniform registers (UTCHMMA takes its descriptors,
    // the TMEM address and the instruction descriptor from the uniform datapath): a value
    // ptxas cannot prove warp-uniform costs an R2UR round trip per operand and MMA, and with
    // MMAs this small (16-32 clocks of tensor work) the single issuing warp, not the tensor
    // core, sets the pace.  redux.sync returns in a uniform register; the descriptors are
    // rebuilt from the (uniform) loop counters instead of being carried in registers.
    auto uni = [](uint32_t v) { return __reduce_or_sync(0xffffffffu, v); };
    const uint32_t idesc = uni(instr_desc<kHalf>(N)), idesc2 = uni(instr_desc<kHalf>(2 * N));
    const uint32_t tmem_u = uni(tmem);
    const uint32_t a_base = smem_u32(amat), b_base = smem_u32(bmat);
    const uint64_t a_tmpl = smem_desc(0, p.plane, 128), b_tmpl = smem_desc(0, 2 * p.bplane, 128);
    const uint32_t a_top = uni((uint32_t)(a_tmpl >> 32)), b_top = uni((uint32_t)(b_tmpl >> 32));
    const uint32_t a_mid = (uint32_t)a_tmpl, b_mid = (uint32_t)b_tmpl;     // LBO field, bits 16..29
    const uint32_t qstep = uni((uint32_t)(2 * p.plane) >> 4), bstep = uni((uint32_t)(4 * p.bplane) >> 4);
    const uint32_t lo_off = uni((uint32_t)(Q * p.plane) >> 4);
    const uint32_t bh0 = uni(b_mid | (b_base >> 4));
    const uint32_t a_first = uni(a_mid | (a_base >> 4)), a_stage = uni((uint32_t)stage_bytes >> 4);
    const int Q2 = Q >> 1;
    for (int k = 0; k < nrows; ++k) {
      const int s = k % kTcStages, b = k & 1;
      mbar_wait(full + s, (k / kTcStages) & 1);
      if (k >= 2) mbar_wait(acc_empty + b, ((k >> 1) - 1) & 1);
      tc_fence_after();
      {
        // two accumulators per row tile: hi * [hi | lo] in columns [0, 2N), lo * hi in
        // [64, 64 + N) -- MMAs into different columns do not wait for each other
        const uint32_t d = tmem_u + (uint32_t)(b * 128);
        // Descriptors differ only in their start-address field (low 14 bits, 16-byte units).
        const uint32_t a_hi0 = a_first + (uint32_t)s * a_stage;
        if (Q2 == 1) {
          // one K step per tap: tap v moves the image operand by one 16-byte row and the
          // filter operand by one pair of planes
#pragma unroll 4
          for (int v = 0; v < wd; ++v) {
            const uint32_t ah = a_hi0 + (uint32_t)v, bh = bh0 + (uint32_t)v * bstep;
            umma_pair<kHalf>(d, d + 64u, ah, ah + lo_off, a_top, bh, b_top, idesc2, idesc,
                             (uint32_t)v);
          }
        } else {
          for (int v = 0; v < wd; ++v) {
#pragma unroll 2
            for (int q2 = 0; q2 < Q2; ++q2) {
              const uint32_t ah = a_hi0 + (uint32_t)v + (uint32_t)q2 * qstep;
              const uint32_t bh = bh0 + (uint32_t)(v * Q2 + q2) * bstep;
              umma_pair<kHalf>(d, d + 64u, ah, ah + lo_off, a_top, bh, b_top, idesc2, idesc,
                               (uint32_t)(v | q2));
            }
          }
        }
        umma_commit(freeb + s);        // the stage may be overwritten
        umma_commit(acc_full + b);     // the accumulator may be read
      }
      __syncwarp();
    }
  } else {
    // ====================== workers: warps 1..4 epilogue, warps 5..8 staging ============== //
    const bool stager = warp >= 5;
    const int wt = stager ? tid - 160 : tid - 32;          // 0 .. 127 within the group
    const int npx = p.npx;
    auto stage_row = [&](int k) {
      const int s = k % kTcStages;
      if (k >= kTcStages) mbar_wait(freeb + s, ((k / kTcStages) - 1) & 1);
      const int r = r0 + k;
      unsigned char* hi_base = amat + (size_t)s * stage_bytes;
      unsigned char* lo_base = hi_base + (size_t)Q * p.plane;
      const float* xrow = xs + (size_t)r * p.W * p.C;
      // all of a thread's loads of the row first (they are independent), then the
      // conversions: one exposed memory latency per row instead of one per element
      constexpr int kBatch = kHalf ? 3 : 6;
      const int total = npx * Q;
      for (int e0 = wt; e0 < total; e0 += kBatch * kTcWorkers) {
        Chunk<kHalf> vals[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int e = e0 + u * kTcWorkers;
          const int px = e / Q, q = e - px * Q;
          vals[u].zero();
          if (e < total && j0 + px < p.W) vals[u].load(xrow + (size_t)(j0 + px) * p.C + kCw * q);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int e = e0 + u * kTcWorkers;
          if (e < total) {
            const int px = e / Q, q = e - px * Q;
            uint4 hi, lo;
            vals[u].split(sx, hi, lo);
            *reinterpret_cast<uint4*>(hi_base + (size_t)q * p.plane + 16 * px) = hi;
            *reinterpret_cast<uint4*>(lo_base + (size_t)q * p.plane + 16 * px) = lo;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(full + s);
    };
    auto epilogue = [&](int k) {
      const int b = k & 1;
      mbar_wait(acc_full + b, (k >> 1) & 1);
      tc_fence_after();
      // this warp's quarter of the TMEM lanes: lane = pixel of the tile
      const int quarter = warp & 3;
      const int pix = quarter * 32 + lane;
      float g[32];
      {
        const uint32_t t0 = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * 128);
        float g2[32];
        tmem_ld32(t0, g);
        tmem_ld32(t0 + 64u, g2);
        if (N == 32) {
#pragma unroll
          for (int u = 0; u < 32; ++u) g[u] += g2[u];
          tmem_ld32(t0 + 32u, g2);
#pragma unroll
          for (int u = 0; u < 32; ++u) g[u] += g2[u];
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) g[u] = (g[u] + g2[u]) + g[16 + u];
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty + b);
      // diagonal sum: G[(r, j), u] belongs to output row r - u
      const int r = r0 + k;
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int i = r - u;
        if (u < p.h && i >= i0 && i < i1) ring[(i & 31) * kTileM + pix] += g[u];
      }
      // output row r - (h - 1) has all its h terms now
      named_bar_sync(1, kTcWorkers);
      const int done = r - (p.h - 1);
      if (done >= i0 && done < i1) {
        float* slot = ring + (done & 31) * kTileM;
        const int j = j0 + wt;
        if (j < p.Pw) p.out[((size_t)sample * p.Ph + done) * p.Pw + j] = (slot[wt] * inv_x) * inv_f;
        slot[wt] = 0.f;
      }
      named_bar_sync(1, kTcWorkers);
    };
    if (stager) {
      for (int k = 0; k < nrows; ++k) stage_row(k);
    } else {
      for (int k = 0; k < nrows; ++k) epilogue(k);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

// Returns SRL_E_UNSUPPORTED for shapes outside the tensor-core path (the caller then
// uses the FP32 kernel of siam.cu).
int siam_correlation_tc(const float* x, const float* f, float* out, int B, int H, int W, int C,
                        int h, int wd, cudaStream_t stream) {
  const int Ph = H - h + 1, Pw = W - wd + 1;
  if (C % 8 != 0 || h > 32 || h < 1 || wd < 1 || Ph < 1 || Pw < 1 ||
      ((uintptr_t)x & 15) != 0 || ((uintptr_t)f & 15) != 0)
    return SRL_E_UNSUPPORTED;
  SiamTcParams p;
  p.x = x;
  p.f = f;
  p.out = out;
  p.H = H; p.W = W; p.C = C; p.h = h; p.wd = wd; p.Ph = Ph; p.Pw = Pw;
  // FP16 pairs when the channels fill whole 16-element K steps, else TF32 pairs
  // (SRL_SIAM_TC=tf32 forces the latter: tests, A/B)
  const char* fmt = getenv("SRL_SIAM_TC");
  const bool half = C % 16 == 0 && !(fmt && fmt[0] == 't');
  p.Q = half ? C / 8 : C / 4;
  p.N = (h + 15) / 16 * 16;
  p.npx = kTileM + wd - 1;
  p.plane = 16 * (p.npx | 1);                  // odd number of 16-byte slots: no bank conflicts
  p.bplane = 16 * p.N;
  p.jblocks = (Pw + kTileM - 1) / kTileM;
  const size_t smem = 256 + (size_t)2 * wd * p.Q * p.bplane +
                      (size_t)kTcStages * 2 * p.Q * p.plane + (size_t)32 * kTileM * 4;
  if (smem > 227 * 1024 || p.plane >= (1 << 18) || p.bplane >= (1 << 18)) return SRL_E_UNSUPPORTED;
  // bands of output rows: enough CTAs for every SM, not more rows re-staged than needed
  const int sms = std::max(1, sm_count());
  int bands = 1;
  while ((long long)B * p.jblocks * bands < sms && bands < Ph &&
         (Ph + bands) / (bands + 1) >= 8)
    ++bands;
  p.bands = bands;
  p.band = (Ph + bands - 1) / bands;
  p.bands = (Ph + p.band - 1) / p.band;
  auto kernel = half ? siam_tc_kernel<true> : siam_tc_kernel<false>;
  SRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<B * p.bands * p.jblocks, kTcThreads, smem, stream>>>(p);
  return check_launch("siam_tc_kernel");
}

}  // namespace srl
