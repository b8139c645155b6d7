// Siamese correlation layer (reference: stackrl/nets/layers.py:21-38,
// used by PseudoSiamFCN / DeepQSiamFCN, nets/models.py:89, 182):
//   out[b, i, j] = sum_{u, v, c} x[b, i+u, j+v, c] * w[b, u, v, c]
// i.e. tf.nn.conv2d(x[b][None], w[b][..., None], strides=1, padding='VALID') for
// every sample b, channels-last like the reference's tensors.  It is the
// sum-product twin of the max-plus drop search over the same window geometry.
//
// float32 products and float32 FMA accumulation like the reference's float32
// convolution; tensor cores would mean TF32/BF16 operands, i.e. LESS precision than
// the reference computes in, so the kernel stays on the FP32 pipe and is built to
// keep it fed from shared memory:
//  * a CTA owns a band of output rows x a block of output columns of one sample and
//    walks the channels four at a time (one 16-byte slot per pixel): the input patch
//    and the filter of that channel quad are staged in shared memory;
//  * a thread owns 2 output rows x 8 output columns.  For every INPUT row it slides
//    an 8-slot register window along the filter columns (one new LDS.128 per step)
//    and feeds the row to both of its output rows (filter rows u and u-1; the first
//    and the last input row of the window meet one filter row only):
//    32 FFMA2 (fma.rn.f32x2: the even and the odd channels of a quad in one
//    instruction, accumulator pairs) per 3 LDS.128, two of them warp-wide broadcasts;
//  * the patch rows carry one spare slot after every 8 pixels, so the 8 lanes of a
//    quarter warp (column blocks 8 pixels = 128 B apart) hit different banks;
//  * the accumulator pairs are folded into per-output totals after every input row
//    (512 products per pair), a two-level sum that stays within ~2e-6 of the exact
//    result for the 16k-product windows of the reference geometry.
// Inputs must be finite when the filter width is not a multiple of 8: its zero-padded
// columns multiply pixels just outside the window (0 * inf would give NaN where the
// reference gives a finite value).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

constexpr int kSiamTI = 2;      // output rows per thread
constexpr int kSiamTJ = 8;      // output columns per thread
constexpr int kSiamMaxThreads = 384;

struct SiamParams {
  const float* x;     // [B, H, W, C]
  const float* w;     // [B, h, wd, C]
  float* out;         // [B, Ph, Pw]
  int H, W, C, h, wd, Ph, Pw;
  int BI, BJ;         // output rows / columns per CTA (multiples of 2 / 8)
  int nbi, nbj;       // bands per sample
  int ncg;            // column groups per CTA = BJ / 8
  int rows_in;        // BI + h - 1
  int cols_in;        // BJ + wdp - 1
  int xpitch;         // slots per staged patch row (skewed)
  int wdp;            // filter columns rounded up to a multiple of 8
};

// FFMA2 (fma.rn.f32x2, sm_100): two float32 FMAs per issued instruction.  A pixel's
// channel quad is two 64-bit register pairs; an accumulator pair holds the partial
// sums of the even and of the odd channels.
typedef unsigned long long f32x2;
struct Quad {
  f32x2 lo, hi;      // channels (0, 1) and (2, 3)
};
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 dot4(f32x2 acc, const Quad& a, const Quad& b) {
  return ffma2(a.hi, b.hi, ffma2(a.lo, b.lo, acc));
}
__device__ __forceinline__ Quad lds_quad(const float4* p) {
  const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
  Quad q;
  q.lo = v.x;
  q.hi = v.y;
  return q;
}
__device__ __forceinline__ float pair_sum(f32x2 v) {
  float e, o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(e), "=f"(o) : "l"(v));
  return __fadd_rn(e, o);
}

// One input row of the thread's strip against filter row(s): ROWS = 3 feeds output
// row 0 with `wrow0` and output row 1 with `wrow1`; ROWS = 1 / 2 only the first /
// second (the first and the last input row of a thread's window meet one filter row).
template <int ROWS>
__device__ __forceinline__ void sweep_row(const float4* __restrict__ xrow,
                                          const float4* __restrict__ wrow0,
                                          const float4* __restrict__ wrow1, int wdp,
                                          float (&tot)[2][8]) {
  f32x2 acc[2][8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[a][c] = 0ull;
  Quad xw[8];
#pragma unroll
  for (int t = 0; t < 7; ++t) xw[t] = lds_quad(xrow + t);
  for (int vc = 0; vc < wdp; vc += 8) {
    const float4* xc = xrow + (vc >> 3) * 9;
#pragma unroll
    for (int vv = 0; vv < 8; ++vv) {
      // column (vc + vv + 7) of the thread's strip: skewed slot 7 for vv = 0,
      // 8 + vv after the spare slot otherwise
      xw[(7 + vv) & 7] = lds_quad(xc + (vv == 0 ? 7 : 8 + vv));
      Quad w0, w1;
      if constexpr (ROWS & 1) w0 = lds_quad(wrow0 + vc + vv);
      if constexpr (ROWS & 2) w1 = lds_quad(wrow1 + vc + vv);
#pragma unroll
      for (int tj = 0; tj < 8; ++tj) {
        const Quad xv = xw[(tj + vv) & 7];
        if constexpr (ROWS & 1) acc[0][tj] = dot4(acc[0][tj], xv, w0);
        if constexpr (ROWS & 2) acc[1][tj] = dot4(acc[1][tj], xv, w1);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if constexpr (ROWS & 1) tot[0][c] = __fadd_rn(tot[0][c], pair_sum(acc[0][c]));
    if constexpr (ROWS & 2) tot[1][c] = __fadd_rn(tot[1][c], pair_sum(acc[1][c]));
  }
}

// Channels [c0, c0+4) of one pixel (zero beyond C).
__device__ __forceinline__ float4 load_quad(const float* __restrict__ px, int c0, int C,
                                            bool vec) {
  if (vec) return __ldg(reinterpret_cast<const float4*>(px + c0));
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c0 < C) q.x = __ldg(px + c0);
  if (c0 + 1 < C) q.y = __ldg(px + c0 + 1);
  if (c0 + 2 < C) q.z = __ldg(px + c0 + 2);
  if (c0 + 3 < C) q.w = __ldg(px + c0 + 3);
  return q;
}

__global__ void __launch_bounds__(kSiamMaxThreads)
siam_correlation_kernel(const SiamParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* xs = reinterpret_cast<float4*>(smem_raw);               // [rows_in][xpitch]
  float4* ws = xs + (size_t)p.rows_in * p.xpitch;                 // [h][wdp]

  int blk = blockIdx.x;
  const int bj = blk % p.nbj; blk /= p.nbj;
  const int bi = blk % p.nbi;
  const int b = blk / p.nbi;
  const int i0 = bi * p.BI, j0 = bj * p.BJ;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int rg = tid / p.ncg, cg = tid - rg * p.ncg;
  const bool active = rg * kSiamTI < p.BI;
  const bool vec = (p.C & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.w)) & 15) == 0;
  const float* xb = p.x + (size_t)b * p.H * p.W * p.C;
  const float* wb = p.w + (size_t)b * p.h * p.wd * p.C;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  float tot[kSiamTI][kSiamTJ];
#pragma unroll
  for (int a = 0; a < kSiamTI; ++a)
#pragma unroll
    for (int c = 0; c < kSiamTJ; ++c) tot[a][c] = 0.f;

  const int nquads = (p.C + 3) >> 2;
  for (int cq = 0; cq < nquads; ++cq) {
    const int c0 = cq * 4;
    __syncthreads();                               // previous quad's sweep is done
    // patch rows by warp, pixels by lane (no index division; a warp's loads of one
    // row are one contiguous run of the image)
    for (int rr = tid >> 5; rr < p.rows_in; rr += nthr >> 5) {
      const int gr = i0 + rr;
      const float* src = xb + ((size_t)gr * p.W + j0) * p.C;
      float4* dst = xs + (size_t)rr * p.xpitch;
      const int live = gr < p.H ? min(p.cols_in, p.W - j0) : 0;   // pixels inside the image
#pragma unroll 4
      for (int q = tid & 31; q < p.cols_in; q += 32)
        dst[q + (q >> 3)] = q < live ? load_quad(src + (size_t)q * p.C, c0, p.C, vec) : zero4;
    }
    for (int k = tid; k < p.h * p.wdp; k += nthr) {
      const int u = k / p.wdp, v = k - u * p.wdp;
      float4 val = zero4;
      if (v < p.wd) val = load_quad(wb + ((size_t)u * p.wd + v) * p.C, c0, p.C, vec);
      ws[(size_t)u * p.wdp + v] = val;
    }
    __syncthreads();
    if (!active) continue;

    // input row (rg*2 + rr) feeds output row 0 with filter row rr and output row 1
    // with filter row rr-1
    const float4* xbase = xs + (size_t)(rg * kSiamTI) * p.xpitch + cg * 9;
    sweep_row<1>(xbase, ws, ws, p.wdp, tot);
    for (int rr = 1; rr < p.h; ++rr)
      sweep_row<3>(xbase + (size_t)rr * p.xpitch, ws + (size_t)rr * p.wdp,
                   ws + (size_t)(rr - 1) * p.wdp, p.wdp, tot);
    sweep_row<2>(xbase + (size_t)p.h * p.xpitch, ws, ws + (size_t)(p.h - 1) * p.wdp, p.wdp, tot);
  }

  if (!active) return;
#pragma unroll
  for (int a = 0; a < kSiamTI; ++a) {
    const int li = rg * kSiamTI + a;
    const int i = i0 + li;
    if (li >= p.BI || i >= p.Ph) continue;
    float* orow = p.out + ((size_t)b * p.Ph + i) * p.Pw;
#pragma unroll
    for (int c = 0; c < kSiamTJ; ++c) {
      const int j = j0 + cg * kSiamTJ + c;
      if (j < p.Pw) orow[j] = tot[a][c];
    }
  }
}

}  // namespace

int siam_correlation_f32(const float* x, const float* w, float* out, int B, int H, int W,
                         int C, int h, int wd, cudaStream_t stream) {
  SRL_REQUIRE(B >= 0 && C >= 1 && h >= 1 && wd >= 1 && H >= h && W >= wd, SRL_E_INVALID,
              "siam_correlation: bad shape B=%d H=%d W=%d C=%d h=%d w=%d", B, H, W, C, h, wd);
  if (B == 0) return SRL_OK;
  SRL_REQUIRE(x && w && out, SRL_E_INVALID, "siam_correlation: null pointer");
  // Tensor-core path (siam_tc.cu: tcgen05, 3xTF32) for the shapes it covers and batches
  // that fill the GPU with whole samples (one CTA per sample: a band split re-stages
  // the filter and h - 1 image rows per band, which the FP32 kernel does not pay);
  // the FP32 FFMA2 kernel below for the rest.  SRL_SIAM_MODE: 0 FP32 only, 1 automatic
  // (default), 2 tensor cores whenever the shape allows (tests, A/B).
  const int sms = sm_count();
  int mode = 1;
  if (const char* m = getenv("SRL_SIAM_MODE")) mode = atoi(m);
  const long long whole = (long long)B * ((W - wd + 1 + 127) / 128);
  if (mode == 2 || (mode == 1 && 4 * whole >= 3ll * sms)) {
    const int rc = siam_correlation_tc(x, w, out, B, H, W, C, h, wd, stream);
    if (rc != SRL_E_UNSUPPORTED) return rc;
  }
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "siam_correlation: no CUDA device");
  SiamParams p;
  p.x = x; p.w = w; p.out = out;
  p.H = H; p.W = W; p.C = C; p.h = h; p.wd = wd;
  p.Ph = H - h + 1; p.Pw = W - wd + 1;
  p.wdp = (wd + 7) & ~7;
  // Tile search: a CTA covers ceil(Ph/nbi) rows x ceil(Pw/nbj) columns (rounded to the
  // thread tile) with one thread per 2x8 outputs.  Every thread does the same work, so
  // a launch costs (waves of CTAs) x (warps per SM sub-partition); ties go to the
  // smaller halo.  The patch of the tile must fit shared memory.
  const size_t filt = (size_t)h * p.wdp * 16;
  auto band = [&](int n, int parts, int unit) { return (((n + parts - 1) / parts) + unit - 1) / unit * unit; };
  int force_bi = 0, force_bj = 0;
  if (const char* t = getenv("SRL_SIAM_TILE")) sscanf(t, "%d,%d", &force_bi, &force_bj);   // tuning override
  double best = 1e300;
  int best_nbi = 0, best_nbj = 0;
  for (int nbj = 1; nbj <= (p.Pw + kSiamTJ - 1) / kSiamTJ; ++nbj) {
    const int BJ = band(p.Pw, nbj, kSiamTJ), ncg = BJ / kSiamTJ;
    if ((p.Pw + BJ - 1) / BJ != nbj || ncg > kSiamMaxThreads) continue;
    const int cols_in = BJ + p.wdp - 1, xpitch = cols_in + (cols_in >> 3) + 1;
    for (int nbi = 1; nbi <= (p.Ph + kSiamTI - 1) / kSiamTI; ++nbi) {
      const int BI = band(p.Ph, nbi, kSiamTI);
      if ((p.Ph + BI - 1) / BI != nbi) continue;
      const int threads = (BI / kSiamTI) * ncg;
      const size_t smem = (size_t)(BI + h - 1) * xpitch * 16 + filt;
      if (threads > kSiamMaxThreads || smem > 220 * 1024) continue;
      const int warps = (threads + 31) / 32;
      const size_t ctas = (size_t)B * nbi * nbj;
      const int per_sm = (int)std::min<size_t>((220 * 1024) / smem, (size_t)(kSiamMaxThreads / (warps * 32)));
      const double waves = (double)((ctas + (size_t)sms * per_sm - 1) / ((size_t)sms * per_sm));
      const double rounds = (double)((warps * per_sm + 3) / 4);
      const double halo = (double)(BI + h - 1) * cols_in / ((double)BI * BJ);
      double cost = waves * rounds * (1. + 0.02 * halo);
      if (force_bi) cost = (nbi == force_bi && nbj == force_bj) ? 0. : 1e299;
      if (cost < best) {
        best = cost;
        best_nbi = nbi;
        best_nbj = nbj;
      }
    }
  }
  SRL_REQUIRE(best_nbi > 0, SRL_E_UNSUPPORTED,
              "siam_correlation: a %dx%d filter on %d-column rows exceeds shared memory", h, wd,
              W);
  p.nbi = best_nbi; p.nbj = best_nbj;
  p.BJ = band(p.Pw, p.nbj, kSiamTJ);
  p.ncg = p.BJ / kSiamTJ;
  p.cols_in = p.BJ + p.wdp - 1;
  p.xpitch = p.cols_in + (p.cols_in >> 3) + 1;
  p.BI = band(p.Ph, p.nbi, kSiamTI);
  auto smem_for = [&](int bi) { return (size_t)(bi + h - 1) * p.xpitch * 16 + filt; };
  p.rows_in = p.BI + h - 1;
  const size_t smem = smem_for(p.BI);
  int threads = (p.BI / kSiamTI) * p.ncg;
  threads = (threads + 31) / 32 * 32;
  SRL_REQUIRE(threads <= kSiamMaxThreads, SRL_E_UNSUPPORTED, "siam_correlation: tile %dx%d", p.BI,
              p.BJ);
  SRL_CUDA(cudaFuncSetAttribute(siam_correlation_kernel,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t grid = (size_t)B * p.nbi * p.nbj;
  SRL_REQUIRE(grid <= 0x7fffffffu, SRL_E_UNSUPPORTED, "siam_correlation: batch too large");
  siam_correlation_kernel<<<(unsigned)grid, threads, smem, stream>>>(p);
  return check_launch("siam_correlation_kernel");
}

}  // namespace srl
