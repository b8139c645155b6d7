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

__device__ __forceinline__ float vimax3(float a, float b, float c) {
  // signed-integer 3-input max on the float bit patterns (VIMNMX3): orders
  // non-negative floats and -inf exactly like the float max.
  return __int_as_float(
      __vimax3_s32(__float_as_int(a), __float_as_int(b), __float_as_int(c)));
}
__device__ __forceinline__ float vimax(float a, float b) {
  int d;
  asm volatile("max.s32 %0, %1, %2;"
               : "=r"(d)
               : "r"(__float_as_int(a)), "r"(__float_as_int(b)));
  return __int_as_float(d);
}

// VARIANT: 0 FADD+FMNMX | 1 2xFADD+FMNMX3 | 2 FADD2+FMNMX3 (float sweep, any sign)
//          3 FADD only | 4 FMNMX only | 5 FMNMX3 only | 6 FADD2 only
//          7 FADD2+VIMNMX3(s32) (float sweep, non-negative tiles) | 8 VIMNMX3 only
//          9 FADD+VIMNMX(s32) | 14 warp-specialised FADD2 / VIMNMX3 | 15 basic-block
//          separated groups | 17 VIADDMNMX.S16x2 (fixed-point sweep) | 18 VIADDMNMX.S32
//          (uint8 integer-key sweep)
// Every op takes a loop-carried accumulator as an operand, so ptxas can neither
// hoist it out of the loop nor fold it.  Dependency distance is >= 8 ops.
// "cells" are counted as if each variant did the full (add, max) pair work of
// its mixed counterpart: variants 3-6, 8 report the rate of the single op class
// in cell units (2 cells per FADD2 / FMNMX3 / VIMNMX3, 1 per FADD / FMNMX).
template <int VARIANT>
__global__ void __launch_bounds__(256) addmax_kernel(const float* in, float* out,
                                                      int iters, int never) {
  float nv[kV];
  float acc[kT];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kV; ++k) nv[k] = in[(tid * 3 + k) & 1023];
#pragma unroll
  for (int t = 0; t < kT; ++t) acc[t] = in[(tid + t) & 1023];

  float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float magic = __int_as_float(never);
  for (int it = 0; it < (VARIANT >= 17 ? 0 : iters); ++it) {
#pragma unroll
    for (int v = 0; v < kV; v += 2) {
#pragma unroll
      for (int t = 0; t < (VARIANT >= 14 ? 0 : kT); ++t) {
        const int k = ((t + 8) % kT) & ~1;     // even-aligned register pair
        const int k1 = (t + 8) % kT, k2 = (t + 9) % kT;
        if constexpr (VARIANT == 0) {
          acc[t] = vmax(acc[t], vadd(acc[k1], nv[v]));
          acc[t] = vmax(acc[t], vadd(acc[k2], nv[v + 1]));
        } else if constexpr (VARIANT == 1) {
          acc[t] = vmax3(acc[t], vadd(acc[k1], nv[v]), vadd(acc[k2], nv[v + 1]));
        } else if constexpr (VARIANT == 2 || VARIANT == 7) {
          float s0, s1;
          // (odd t swaps the rock pair so the two adds are not common
          // subexpressions of the even neighbour's)
          if (t & 1) vadd2(s0, s1, acc[k], acc[k + 1], nv[v + 1], nv[v]);
          else vadd2(s0, s1, acc[k], acc[k + 1], nv[v], nv[v + 1]);
          if constexpr (VARIANT == 2) acc[t] = vmax3(acc[t], s0, s1);
          else acc[t] = vimax3(acc[t], s0, s1);
        } else if constexpr (VARIANT == 3) {
          // two DEPENDENT adds, both live (round 1 overwrote the first result, which
          // ptxas removed: the reported rate counted instructions that never ran)
          acc[t] = vadd(vadd(acc[k1], nv[v]), nv[v + 1]);
        } else if constexpr (VARIANT == 4) {
          // two maxima that cannot fuse into one FMNMX3: the inner one feeds an add-free
          // chain through a second accumulator
          acc[t] = vmax(acc[k1], nv[v]);
          acc[k2] = vmax(acc[t], nv[v + 1]);
        } else if constexpr (VARIANT == 5) {
          acc[t] = vmax3(acc[k1], acc[k2], nv[v]);
        } else if constexpr (VARIANT == 6) {
          if ((t & 1) == 0) {
            float s0, s1;
            vadd2(s0, s1, acc[k], acc[k + 1], nv[v], nv[v + 1]);
            acc[(t + 4) % kT] = s0;
            acc[(t + 5) % kT] = s1;
          }
        } else if constexpr (VARIANT == 8) {
          acc[t] = vimax3(acc[k1], acc[k2], nv[v]);
        } else if constexpr (VARIANT == 9) {
          acc[t] = vimax(acc[t], vadd(acc[k1], nv[v]));
          acc[t] = vimax(acc[t], vadd(acc[k2], nv[v + 1]));
        }
      }
      // Variant 14: warp-specialised -- even warps issue only FADD2 (rock pair
      // reused), odd warps only VIMNMX3: do the two pipes overlap when the
      // instructions come from different warps?
      if constexpr (VARIANT == 14) {
        if ((threadIdx.x >> 5) & 1) {
#pragma unroll
          for (int t = 0; t < kT; ++t)
            acc[t] = vimax3(acc[(t + 8) % kT], acc[(t + 9) % kT], nv[v]);
        } else {
#pragma unroll
          for (int t = 0; t < kT; ++t) {
            const int k = (2 * t + 8) % kT;
            float s0, s1;
            vadd2(s0, s1, acc[k], acc[k + 1], nv[v], nv[v + 1]);
            acc[(2 * t + 4) % kT] = s0;
            acc[(2 * t + 5) % kT] = s1;
          }
        }
      }
      // Variant 15: groups of 8 adds and 8 maxes kept apart by never-taken
      // data-dependent branches (basic-block boundaries ptxas cannot schedule
      // across), so that the shared rock pair is served by the operand-reuse cache.
      if constexpr (VARIANT == 15) {
        constexpr int G = 8;
#pragma unroll
        for (int t0 = 0; t0 < kT; t0 += G) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            // G distinct wall pairs, one rock pair per group (as in the sweep)
            const int k = (2 * g + 8) % kT, vv = (v + (t0 ? 2 : 0)) % kV;
            vadd2(s0[g], s1[g], acc[k], acc[k + 1], nv[vv], nv[vv + 1]);
          }
          if (s1[G - 1] == magic) goto done;      // never true; ends the add block
#pragma unroll
          for (int g = 0; g < G; ++g) {
            acc[t0 + g] = vimax3(acc[t0 + g], s0[g], s1[g]);
          }
          if (acc[t0 + G - 1] == magic) goto done;  // ends the max block
        }
      }
    }
  }
  // Variants 17/18: DPX fused integer add + max, 16x2 packed (two cells per
  // instruction) and 32-bit (one cell per instruction).
  if constexpr (VARIANT == 17 || VARIANT == 18) {
    unsigned ia[kT], in_[kV];
#pragma unroll
    for (int t = 0; t < kT; ++t) ia[t] = __float_as_uint(acc[t]);
#pragma unroll
    for (int k = 0; k < kV; ++k) in_[k] = __float_as_uint(nv[k]) & 0x00ff00ffu;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int v = 0; v < kV; ++v) {
#pragma unroll
        for (int t = 0; t < kT; ++t) {
          const int k1 = (t + 8) % kT;
          if constexpr (VARIANT == 17) ia[t] = __viaddmax_s16x2(ia[k1], in_[v], ia[t]);
          else ia[t] = (unsigned)__viaddmax_s32((int)ia[k1], (int)in_[v], (int)ia[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < kT; ++t) acc[t] = __uint_as_float(ia[t]);
  }
done:
  float r = acc[0];
  if constexpr (VARIANT == 15) {
#pragma unroll
    for (int g = 0; g < 8; ++g) r = fmaxf(r, fmaxf(s0[g], s1[g]));
  }
#pragma unroll
  for (int t = 1; t < kT; ++t) r = fmaxf(r, acc[t]);
  out[tid] = r;
}


// ---- FP32 FMA issue rate (roofline denominator of siam_correlation_kernel) ----- //
// VARIANT 0: FFMA, three 32-bit register operands | 1: FFMA2 (fma.rn.f32x2) in the
// kernel's pattern: 8 consecutive instructions share one multiplicand pair
// (operand-reuse cache) | 2: FFMA2 with three distinct register pairs.
__device__ __forceinline__ unsigned long long vfma2(unsigned long long a, unsigned long long b,
                                                    unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float vfma(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}

template <int VARIANT>
__global__ void __launch_bounds__(256) fma_kernel(const float* in, float* out, int iters) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc[16], x[8], w[2];
  float accf[32], xf[8], wf[4];
#pragma unroll
  for (int t = 0; t < 16; ++t) acc[t] = pack2(in[(tid + t) & 1023], in[(tid + 2 * t + 1) & 1023]);
#pragma unroll
  for (int t = 0; t < 8; ++t) x[t] = pack2(in[(tid * 3 + t) & 1023], in[(tid * 5 + t) & 1023]);
  w[0] = pack2(in[tid & 1023] * 1e-3f, in[(tid + 7) & 1023] * 1e-3f);
  w[1] = pack2(in[(tid + 9) & 1023] * 1e-3f, in[(tid + 11) & 1023] * 1e-3f);
#pragma unroll
  for (int t = 0; t < 32; ++t) accf[t] = in[(tid + t) & 1023];
#pragma unroll
  for (int t = 0; t < 8; ++t) xf[t] = in[(tid * 3 + t) & 1023];
#pragma unroll
  for (int t = 0; t < 4; ++t) wf[t] = in[(tid + 13 * t) & 1023] * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if constexpr (VARIANT == 0) {
#pragma unroll
        for (int t = 0; t < 32; ++t) accf[t] = vfma(xf[t & 7], wf[(t >> 3) & 3], accf[t]);
      } else if constexpr (VARIANT == 1) {
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[t] = vfma2(x[t & 7], w[t >> 3], acc[t]);
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[t] = vfma2(x[t & 7], acc[(t + 8) & 15], acc[t]);
      }
    }
  }
  float r = 0.f;
  if constexpr (VARIANT == 0) {
#pragma unroll
    for (int t = 0; t < 32; ++t) r += accf[t];
  } else {
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[t]));
      r += lo + hi;
    }
  }
  out[tid] = r;
}

}  // namespace

int microbench_addmax(int variant, int iters, double* host_cells_per_s) {
  SRL_REQUIRE(host_cells_per_s != nullptr && iters > 0 && (variant <= 9 || variant == 14 || variant == 15 || variant == 17 || variant == 18) &&
                  variant >= 0,
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
    switch (variant) {
#define SRL_MB(V) case V: addmax_kernel<V><<<blocks, threads>>>(in, out, iters, 0x7fc12345); break;
      SRL_MB(0) SRL_MB(1) SRL_MB(2) SRL_MB(3) SRL_MB(4) SRL_MB(5) SRL_MB(6) SRL_MB(7)
      SRL_MB(8) SRL_MB(9) SRL_MB(14) SRL_MB(15) SRL_MB(17) SRL_MB(18)
#undef SRL_MB
    }
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
  double cells = (double)blocks * threads * (double)iters * kT * kV;
  if (variant == 6 || variant == 14) cells *= 0.5;
  if (variant == 17) cells *= 2.0;   // two (add, max) cells per instruction   // one FADD2 per two (v, t) steps
  *host_cells_per_s = cells / (best_ms * 1e-3);
  return SRL_OK;
}

int microbench_fma(int variant, int iters, double* host_fma_per_s) {
  SRL_REQUIRE(host_fma_per_s != nullptr && iters > 0 && variant >= 0 && variant <= 2,
              SRL_E_INVALID, "microbench_fma: bad arguments");
  const int sms = sm_count();
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "microbench_fma: no device");
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
    if (variant == 0) fma_kernel<0><<<blocks, threads>>>(in, out, iters);
    else if (variant == 1) fma_kernel<1><<<blocks, threads>>>(in, out, iters);
    else fma_kernel<2><<<blocks, threads>>>(in, out, iters);
    SRL_CUDA(cudaEventRecord(t1));
    SRL_CUDA(cudaEventSynchronize(t1));
    float ms = 0;
    SRL_CUDA(cudaEventElapsedTime(&ms, t0, t1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  int rc = check_launch("fma_kernel");
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(in);
  cudaFree(out);
  if (rc != SRL_OK) return rc;
  *host_fma_per_s = (double)blocks * threads * (double)iters * 4 * 32 / (best_ms * 1e-3);
  return SRL_OK;
}

}  // namespace srl
