// Issue-rate micro-benchmark for the max-plus instruction mix (SURVEY 8d asks
// the builder to measure the FP32 add+max rate: MEASURED_PEAKS.json has no
// non-tensor FP32 figure).  The loop body is the register-only part of
// maxplus_f32_kernel's cell block (T=16 accumulators x VC=16 rock values per
// round, no memory traffic), written with volatile asm so nothing is hoisted
// or folded.  cells/s from this kernel is the roofline `peak` of bench.py.
#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

constexpr int kT = 16, kV = 16;

__device__ __forceinline__ float vadd(float a, float b) {
  float d;
  asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float vmax(float a, float b) {
  float d;
  asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float vmax3(float a, float b, float c) {
  float d;
  asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void vadd2(float& lo, float& hi, float a_lo, float a_hi,
                                      float b_lo, float b_hi) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a_lo), "f"(a_hi));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b_lo), "f"(b_hi));
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(c));
}

template <int VARIANT>
__global__ void __launch_bounds__(256) addmax_kernel(const float* in, float* out,
                                                      int iters) {
  // Every add takes a loop-carried accumulator as an operand, so ptxas can
  // neither hoist it out of the loop nor fold the max.  The dependency distance
  // is kT cells, far beyond the 4-cycle pipe latency.
  float nv[kV];
  float acc[kT];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kV; ++k) nv[k] = in[(tid * 3 + k) & 1023];
#pragma unroll
  for (int t = 0; t < kT; ++t) acc[t] = in[(tid + t) & 1023];

  for (int it = 0; it < iters; ++it) {
    if constexpr (VARIANT == 0) {
#pragma unroll
      for (int v = 0; v < kV; ++v)
#pragma unroll
        for (int t = 0; t < kT; ++t)
          acc[t] = vmax(acc[t], vadd(acc[(t + 8) % kT], nv[v]));
    } else if constexpr (VARIANT == 1) {
#pragma unroll
      for (int v = 0; v < kV; v += 2)
#pragma unroll
        for (int t = 0; t < kT; ++t)
          acc[t] = vmax3(acc[t], vadd(acc[(t + 8) % kT], nv[v]),
                         vadd(acc[(t + 9) % kT], nv[v + 1]));
    } else {
#pragma unroll
      for (int v = 0; v < kV; v += 2)
#pragma unroll
        for (int t = 0; t < kT; ++t) {
          float s0, s1;
          const int k = ((t + 8) % kT) & ~1;     // even-aligned register pair
          // (odd t swaps the rock pair so the two adds are not common
          // subexpressions of the even neighbour's)
          if (t & 1) vadd2(s0, s1, acc[k], acc[k + 1], nv[v + 1], nv[v]);
          else vadd2(s0, s1, acc[k], acc[k + 1], nv[v], nv[v + 1]);
          acc[t] = vmax3(acc[t], s0, s1);
        }
    }
  }
  float r = acc[0];
#pragma unroll
  for (int t = 1; t < kT; ++t) r = fmaxf(r, acc[t]);
  out[tid] = r;
}

}  // namespace

int microbench_addmax(int variant, int iters, double* host_cells_per_s) {
  SRL_REQUIRE(host_cells_per_s != nullptr && iters > 0 && variant >= 0 && variant <= 2,
              SRL_E_INVALID, "microbench_addmax: bad arguments");
  const int sms = sm_count();
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "microbench_addmax: no device");
  const int threads = 256, blocks = sms * 4;
  float *in = nullptr, *out = nullptr;
  SRL_CUDA(cudaMalloc(&in, 1024 * sizeof(float)));
  SRL_CUDA(cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)));
  float host_in[1024];
  for (int k = 0; k < 1024; ++k) host_in[k] = (float)((k * 37) % 101) * 0.01f;
  SRL_CUDA(cudaMemcpy(in, host_in, sizeof(host_in), cudaMemcpyHostToDevice));
  cudaEvent_t t0, t1;
  SRL_CUDA(cudaEventCreate(&t0));
  SRL_CUDA(cudaEventCreate(&t1));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {   // rep 0 is the warm-up
    SRL_CUDA(cudaEventRecord(t0));
    if (variant == 0) addmax_kernel<0><<<blocks, threads>>>(in, out, iters);
    else if (variant == 1) addmax_kernel<1><<<blocks, threads>>>(in, out, iters);
    else addmax_kernel<2><<<blocks, threads>>>(in, out, iters);
    SRL_CUDA(cudaEventRecord(t1));
    SRL_CUDA(cudaEventSynchronize(t1));
    float ms = 0;
    SRL_CUDA(cudaEventElapsedTime(&ms, t0, t1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  int rc = check_launch("addmax_kernel");
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(in);
  cudaFree(out);
  if (rc != SRL_OK) return rc;
  const double cells = (double)blocks * threads * (double)iters * kT * kV;
  *host_cells_per_s = cells / (best_ms * 1e-3);
  return SRL_OK;
}

}  // namespace srl
