// Shared device/host helpers for the stackrl_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "stackrl_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "stackrl_b200 kernels are written for sm_100a (B200) only"
#endif

namespace srl {

// ---- error plumbing (capi.cu) ---------------------------------------------- //
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define SRL_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) return ::srl::fail((code), __VA_ARGS__); \
  } while (0)

#define SRL_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess)                                                  \
      return ::srl::fail(SRL_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

int sm_count();

#ifdef __CUDACC__
#define SRL_HD __host__ __device__
#else
#define SRL_HD
#endif
SRL_HD inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

#ifdef __CUDACC__

constexpr float kNegInf = -__builtin_huge_valf();

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) ------------------ //
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}

// Make the barrier initialisation visible to the async (TMA) proxy.
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Order prior generic-proxy smem accesses before later async-proxy ones.
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar,
                                                      uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_arrive_count(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

// Named barrier among `count` threads (count % 32 == 0) of the CTA.
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void st_release_shared(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v)
               : "memory");
}
__device__ __forceinline__ int ld_acquire_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p))
               : "memory");
  return v;
}

// One try_wait with a suspend-time hint (ns): the warp is parked by the hardware
// until the phase completes or the hint expires.  Returns true when the phase
// with the given parity has completed.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t phase,
                                                   uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Blocking wait that parks the warp instead of spinning.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t phase) {
  while (!mbar_try_wait_hint(bar, phase, 1000000u)) {
  }
}

// global -> shared::cta bulk copy; bytes % 16 == 0, both addresses 16-B aligned.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem,
                                            uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// shared::cta -> global bulk store (bulk async-group completion).
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem,
                                             uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                   dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Wait until the smem sources of all committed bulk stores have been read.
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- sm_100 FP32 max-plus instruction mix ---------------------------------- //
// FMNMX3: d = max(a, b, c) in one ALU-pipe instruction.
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// FADD2: two IEEE round-to-nearest float32 adds in one FMA-pipe instruction.
__device__ __forceinline__ void fadd2(float& lo, float& hi, float a_lo, float a_hi,
                                      float b_lo, float b_hi) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a_lo), "f"(a_hi));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b_lo), "f"(b_hi));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(c));
}

__device__ __forceinline__ float4 lds128(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}

#endif  // __CUDACC__

}  // namespace srl
