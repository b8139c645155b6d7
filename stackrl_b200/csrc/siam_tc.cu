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
// Precision: the reference layer computes in float32.  Every operand is split into
// hi = the value with the low 13 mantissa bits cleared (exactly a TF32 number) and
// lo = value - hi (exact), and three tensor-core products hi*hi + hi*lo + lo*hi
// are accumulated in float32 (3xTF32); a TMEM accumulator only sums the (v, c) of
// one filter row (K = w*C), the 32-term sum over u runs in the epilogue in float32.
// Both operands come from shared memory and N is only the filter height, so the tensor
// core's operand reads (128 x 32 B of image per MMA), not its arithmetic, bound the
// kernel (ncu: tensor pipe 22 % busy, its shared-memory wavefronts at 67 % of peak).  The
// two products that share the image's hi part are therefore ONE MMA: the filter's hi and
// lo planes are stacked along N (columns [0, N) = hi*hi, [N, 2N) = hi*lo), the image's lo
// part multiplies the hi planes alone into columns [0, N), and the epilogue adds the two
// column groups: 11 KB of operand reads per K step instead of 15.
//
// Warp roles (160 threads, one CTA = one band of output rows of one sample):
//   warp 0        allocates TMEM, then one elected lane issues every tcgen05.mma
//   warps 1..4    (a) stage image rows: coalesced loads -> hi / lo planes, three rows
//                 ahead of the MMAs; (b) epilogue: tcgen05.ld of a finished row tile,
//                 diagonal accumulation into a 32-row ring in shared memory, finished
//                 output rows to global memory.
// Synchronisation is mbarriers only (stage full / stage free via tcgen05.commit /
// accumulator full / accumulator empty); the two TMEM accumulators alternate.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

constexpr int kTcThreads = 160;
constexpr int kTcWorkers = 128;       // warps 1..4
constexpr int kTcStages = 3;          // image rows in flight
constexpr int kTileM = 128;           // pixels per tile (UMMA M)

struct SiamTcParams {
  const float* x;     // [B, H, W, C]
  const float* f;     // [B, h, wd, C]
  float* out;         // [B, Ph, Pw]
  int H, W, C, h, wd, Ph, Pw;
  int Q;              // C / 4: 16-byte chunks per pixel
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
// D[tmem] (+)= A[smem] * B[smem], TF32 operands, float32 accumulate.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                          uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  // executed by the whole (converged) warp with warp-uniform operands; one elected
  // lane issues -- the operands then stay on the uniform datapath
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
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both
// K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__device__ __forceinline__ uint32_t instr_desc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

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
  unsigned char* at = smem_raw + 128;
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
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // ---- filter planes, hi and lo (all threads; resident for the whole CTA) ------------- //
  // plane (v, q) holds f[u, v, 4q .. 4q+4) for u = 0 .. N-1 (zero rows beyond h), its lo
  // part right behind its hi part: together they are one K-major operand of 2N rows
  for (int k = tid; k < wd * Q * N; k += kTcThreads) {
    const int u = k % N, vq = k / N;
    const int v = vq / Q, q = vq - v * Q;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u < p.h)
      val = __ldg(reinterpret_cast<const float4*>(fs + ((size_t)u * wd + v) * p.C + 4 * q));
    const float4 hi = make_float4(tf32_hi(val.x), tf32_hi(val.y), tf32_hi(val.z), tf32_hi(val.w));
    const float4 lo = make_float4(val.x - hi.x, val.y - hi.y, val.z - hi.z, val.w - hi.w);
    *reinterpret_cast<float4*>(bmat + (size_t)(2 * vq) * p.bplane + 16 * u) = hi;
    *reinterpret_cast<float4*>(bmat + (size_t)(2 * vq + 1) * p.bplane + 16 * u) = lo;
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
    const uint32_t idesc = instr_desc(N), idesc2 = instr_desc(2 * N);
    const uint32_t a_base = smem_u32(amat), b_base = smem_u32(bmat);
    for (int k = 0; k < nrows; ++k) {
      const int s = k % kTcStages, b = k & 1;
      mbar_wait(full + s, (k / kTcStages) & 1);
      if (k >= 2) mbar_wait(acc_empty + b, ((k >> 1) - 1) & 1);
      tc_fence_after();
      {
        const uint32_t d = tmem + (uint32_t)(b * 64);
        // Descriptors differ only in their start-address field (low 14 bits, 16-byte
        // units): one add per operand and MMA instead of rebuilding 64-bit words.
        const uint32_t a_hi = a_base + (uint32_t)(s * stage_bytes);
        const uint64_t a_tmpl = smem_desc(0, p.plane, 128), b_tmpl = smem_desc(0, 2 * p.bplane, 128);
        const uint32_t a_top = (uint32_t)(a_tmpl >> 32), b_top = (uint32_t)(b_tmpl >> 32);
        const uint32_t a_mid = (uint32_t)a_tmpl, b_mid = (uint32_t)b_tmpl;   // LBO field, bits 16..29
        const uint32_t a_hi0 = a_mid | (a_hi >> 4), a_lo0 = a_mid | ((a_hi + (uint32_t)(Q * p.plane)) >> 4);
        const uint32_t qstep = (uint32_t)(2 * p.plane) >> 4, bstep = (uint32_t)(4 * p.bplane) >> 4;
        uint32_t bh = b_mid | (b_base >> 4);
        uint32_t first = 0;
        for (int v = 0; v < wd; ++v) {
          uint32_t ah = a_hi0 + (uint32_t)v, al = a_lo0 + (uint32_t)v;
#pragma unroll 2
          for (int q = 0; q < Q; q += 2) {
            umma_tf32(d, ah, a_top, bh, b_top, idesc2, first);    // hi * [hi | lo]
            umma_tf32(d, al, a_top, bh, b_top, idesc, 1u);        // lo * hi
            first = 1u;
            ah += qstep; al += qstep; bh += bstep;
          }
        }
        umma_commit(freeb + s);        // the stage may be overwritten
        umma_commit(acc_full + b);     // the accumulator may be read
      }
      __syncwarp();
    }
  } else {
    // =============================== workers ============================================= //
    const int wt = tid - 32;                      // 0 .. 127
    const int npx = p.npx;
    auto stage_row = [&](int k) {
      const int s = k % kTcStages;
      if (k >= kTcStages) mbar_wait(freeb + s, ((k / kTcStages) - 1) & 1);
      const int r = r0 + k;
      unsigned char* hi_base = amat + (size_t)s * stage_bytes;
      unsigned char* lo_base = hi_base + (size_t)Q * p.plane;
      const float* xrow = xs + (size_t)r * p.W * p.C;
      for (int e = wt; e < npx * Q; e += kTcWorkers) {
        const int px = e / Q, q = e - px * Q;
        const int col = j0 + px;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < p.W) val = __ldg(reinterpret_cast<const float4*>(xrow + (size_t)col * p.C + 4 * q));
        const float4 hi = make_float4(tf32_hi(val.x), tf32_hi(val.y), tf32_hi(val.z),
                                      tf32_hi(val.w));
        const float4 lo = make_float4(val.x - hi.x, val.y - hi.y, val.z - hi.z, val.w - hi.w);
        *reinterpret_cast<float4*>(hi_base + (size_t)q * p.plane + 16 * px) = hi;
        *reinterpret_cast<float4*>(lo_base + (size_t)q * p.plane + 16 * px) = lo;
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
        const uint32_t t0 = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * 64);
        tmem_ld32(t0, g);
        if (N == 32) {
          float g2[32];
          tmem_ld32(t0 + 32u, g2);
#pragma unroll
          for (int u = 0; u < 32; ++u) g[u] += g2[u];
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) g[u] += g[16 + u];
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
        if (j < p.Pw) p.out[((size_t)sample * p.Ph + done) * p.Pw + j] = slot[wt];
        slot[wt] = 0.f;
      }
      named_bar_sync(1, kTcWorkers);
    };
    for (int k = 0; k < nrows + 2; ++k) {
      if (k < nrows) stage_row(k);
      if (k >= 2) epilogue(k - 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
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
  p.Q = C / 4;
  p.N = (h + 15) / 16 * 16;
  p.npx = kTileM + wd - 1;
  p.plane = 16 * (p.npx | 1);                  // odd number of 16-byte slots: no bank conflicts
  p.bplane = 16 * p.N;
  p.jblocks = (Pw + kTileM - 1) / kTileM;
  const size_t smem = 128 + (size_t)2 * wd * p.Q * p.bplane +
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
  SRL_CUDA(cudaFuncSetAttribute(siam_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  siam_tc_kernel<<<B * p.bands * p.jblocks, kTcThreads, smem, stream>>>(p);
  return check_launch("siam_tc_kernel");
}

}  // namespace srl
