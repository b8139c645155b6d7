// Gradients of the Siamese correlation layer (stackrl/nets/layers.py:21-38): the DQN
// trains THROUGH the layer (nets/models.py:89, 182 inside DQN.train's gradient tape),
// so a replacement needs the two vector-Jacobian products of
//   out[i, j] = sum_{u, v, c} x[i+u, j+v, c] * f[u, v, c]:
//   grad_f[u, v, c] = sum_{i, j} g[i, j] * x[i+u, j+v, c]      (x correlated with g)
//   grad_x[r, s, c] = sum_{u, v} g[r-u, s-v] * f[u, v, c]      (g scattered through f)
// float32 FMA kernels (the contraction with g has a single input channel: no dense
// GEMM shape without an im2col of g); float32 accumulation, folded into per-output
// totals once per filter / image row like the forward FP32 kernel.
#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

constexpr int kGradThreads = 128;

// ---- grad_f: one CTA per (sample, filter row u); thread = (block of 4 filter columns,
// channel); per output row i the image row i+u and g's row i are staged -------------- //
__global__ void __launch_bounds__(kGradThreads)
siam_grad_filter_kernel(const float* __restrict__ x, const float* __restrict__ g,
                        float* __restrict__ grad_f, int H, int W, int C, int h, int wd, int Ph,
                        int Pw) {
  extern __shared__ __align__(16) float gsm[];
  float* xrow = gsm;                  // [W * C]
  float* grow = gsm + W * C;          // [Pw]
  const int b = blockIdx.x / h, u = blockIdx.x - b * h;
  const float* xs = x + (size_t)b * H * W * C;
  const float* gs = g + (size_t)b * Ph * Pw;
  const int nvb = (wd + 3) / 4;                       // blocks of 4 filter columns
  for (int item0 = 0; item0 < nvb * C; item0 += kGradThreads) {
    const int item = item0 + threadIdx.x;
    const bool live = item < nvb * C;
    const int vb = live ? item / C : 0, c = live ? item - vb * C : 0;
    const int v0 = vb * 4;
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < Ph; ++i) {
      __syncthreads();
      for (int k = threadIdx.x; k < W * C; k += kGradThreads)
        xrow[k] = __ldg(xs + (size_t)(i + u) * W * C + k);
      for (int k = threadIdx.x; k < Pw; k += kGradThreads) grow[k] = __ldg(gs + (size_t)i * Pw + k);
      __syncthreads();
      if (!live) continue;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      // sliding window over j: x[(j + v0 + t) * C + c], t = 0..3
      float w0 = xrow[min(v0, W - 1) * C + c], w1 = xrow[min(v0 + 1, W - 1) * C + c],
            w2 = xrow[min(v0 + 2, W - 1) * C + c];
      for (int j = 0; j < Pw; ++j) {
        const float w3 = xrow[min(j + v0 + 3, W - 1) * C + c];
        const float gv = grow[j];
        acc[0] = __fmaf_rn(gv, w0, acc[0]);
        acc[1] = __fmaf_rn(gv, w1, acc[1]);
        acc[2] = __fmaf_rn(gv, w2, acc[2]);
        acc[3] = __fmaf_rn(gv, w3, acc[3]);
        w0 = w1; w1 = w2; w2 = w3;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) tot[t] = __fadd_rn(tot[t], acc[t]);
    }
    if (live) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (v0 + t < wd)
          grad_f[(((size_t)b * h + u) * wd + v0 + t) * C + c] = tot[t];
    }
  }
}

// ---- grad_x: one CTA per (sample, image row r); thread = pixel s, all channels in
// blocks of 4; the filter rows and g rows that reach row r are read through L1 --------- //
__global__ void __launch_bounds__(kGradThreads)
siam_grad_input_kernel(const float* __restrict__ f, const float* __restrict__ g,
                       float* __restrict__ grad_x, int H, int W, int C, int h, int wd, int Ph,
                       int Pw) {
  extern __shared__ __align__(16) float gsm[];
  float* frow = gsm;                  // [wd * C] one filter row
  float* grow = gsm + wd * C;         // [Pw + 2 * (wd - 1)] one g row, zero padded
  const int b = blockIdx.x / H, r = blockIdx.x - b * H;
  const float* fs = f + (size_t)b * h * wd * C;
  const float* gs = g + (size_t)b * Ph * Pw;
  const int pad = wd - 1;
  for (int s0 = 0; s0 < W; s0 += kGradThreads) {
    const int s = s0 + threadIdx.x;
    for (int c0 = 0; c0 < C; c0 += 4) {
      float tot[4] = {0.f, 0.f, 0.f, 0.f};
      const int u_lo = max(0, r - (Ph - 1)), u_hi = min(h - 1, r);
      for (int u = u_lo; u <= u_hi; ++u) {
        __syncthreads();
        for (int k = threadIdx.x; k < wd * C; k += kGradThreads)
          frow[k] = __ldg(fs + (size_t)u * wd * C + k);
        for (int k = threadIdx.x; k < Pw + 2 * pad; k += kGradThreads) {
          const int j = k - pad;
          grow[k] = (j >= 0 && j < Pw) ? __ldg(gs + (size_t)(r - u) * Pw + j) : 0.f;
        }
        __syncthreads();
        if (s >= W) continue;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int v = 0; v < wd; ++v) {
          const float gv = grow[s - v + pad];               // g[r-u, s-v], 0 outside
          const float* fp = frow + v * C + c0;
          acc[0] = __fmaf_rn(gv, fp[0], acc[0]);
          if (c0 + 1 < C) acc[1] = __fmaf_rn(gv, fp[1], acc[1]);
          if (c0 + 2 < C) acc[2] = __fmaf_rn(gv, fp[2], acc[2]);
          if (c0 + 3 < C) acc[3] = __fmaf_rn(gv, fp[3], acc[3]);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) tot[t] = __fadd_rn(tot[t], acc[t]);
      }
      if (s < W) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (c0 + t < C) grad_x[(((size_t)b * H + r) * W + s) * C + c0 + t] = tot[t];
      }
    }
  }
}

}  // namespace

int siam_correlation_grad_f32(const float* x, const float* f, const float* g, float* grad_x,
                              float* grad_f, int B, int H, int W, int C, int h, int wd,
                              cudaStream_t stream) {
  SRL_REQUIRE(B >= 0 && C >= 1 && h >= 1 && wd >= 1 && H >= h && W >= wd, SRL_E_INVALID,
              "siam_correlation_grad: bad shape B=%d H=%d W=%d C=%d h=%d w=%d", B, H, W, C, h,
              wd);
  if (B == 0) return SRL_OK;
  SRL_REQUIRE(g && (grad_x == nullptr || f) && (grad_f == nullptr || x), SRL_E_INVALID,
              "siam_correlation_grad: null pointer");
  const int Ph = H - h + 1, Pw = W - wd + 1;
  if (grad_f) {
    const size_t smem = ((size_t)W * C + Pw) * 4;
    SRL_REQUIRE(smem <= 200 * 1024, SRL_E_UNSUPPORTED,
                "siam_correlation_grad: image row of %d x %d floats exceeds shared memory", W, C);
    SRL_CUDA(cudaFuncSetAttribute(siam_grad_filter_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    siam_grad_filter_kernel<<<B * h, kGradThreads, smem, stream>>>(x, g, grad_f, H, W, C, h, wd,
                                                                   Ph, Pw);
    const int rc = check_launch("siam_grad_filter_kernel");
    if (rc != SRL_OK) return rc;
  }
  if (grad_x) {
    const size_t smem = ((size_t)wd * C + Pw + 2 * (wd - 1)) * 4;
    SRL_REQUIRE(smem <= 200 * 1024, SRL_E_UNSUPPORTED,
                "siam_correlation_grad: filter row of %d x %d floats exceeds shared memory", wd,
                C);
    SRL_CUDA(cudaFuncSetAttribute(siam_grad_input_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    siam_grad_input_kernel<<<B * H, kGradThreads, smem, stream>>>(f, g, grad_x, H, W, C, h, wd,
                                                                  Ph, Pw);
    return check_launch("siam_grad_input_kernel");
  }
  return SRL_OK;
}

}  // namespace srl
