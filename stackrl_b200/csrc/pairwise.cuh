// numpy's pairwise summation order and the per-dtype arithmetic of the reference,
// shared by difference.cu and corrcoef.cu.
#pragma once

#include "common.cuh"

namespace srl {

constexpr int kMaxOps = 96;
struct PwProg {
  int n_ops;
  short start[kMaxOps];
  short len[kMaxOps];   // > 0: leaf [start, start+len); 0: add the two top entries
};

inline void build_prog(int start, int n, PwProg& prog) {
  if (n <= 128) {
    prog.start[prog.n_ops] = (short)start;
    prog.len[prog.n_ops++] = (short)n;
    return;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  build_prog(start, n2, prog);
  build_prog(start + n2, n - n2, prog);
  prog.start[prog.n_ops] = 0;
  prog.len[prog.n_ops++] = 0;
}

// numpy's pairwise sum of term(0..n-1) following `prog`; term(k) must be called
// with k increasing by one (callers keep running row/column counters).
template <typename F>
__device__ __forceinline__ double pairwise_sum(const PwProg& prog, F term) {
  double stack[8];
  int sp = 0;
  for (int op = 0; op < prog.n_ops; ++op) {
    const int len = prog.len[op];
    if (len == 0) {
      --sp;
      stack[sp - 1] = __dadd_rn(stack[sp - 1], stack[sp]);
      continue;
    }
    double res;
    if (len < 8) {
      res = 0.;
      for (int k = 0; k < len; ++k) res = __dadd_rn(res, term());
    } else {
      double r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = term();
      const int body = len - len % 8;
      for (int i = 8; i < body; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], term());
      }
      res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                      __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
      for (int i = body; i < len; ++i) res = __dadd_rn(res, term());
    }
    stack[sp++] = res;
  }
  return stack[0];
}

__device__ __forceinline__ float div_level(float x, float level) {
  return x == 0.f ? __fmul_rn(x, level) : __fdiv_rn(x, level);
}

// Arithmetic type of the reference for an observation dtype: float32 obs stay
// float32, uint8 / uint8 is float64 in numpy (get_inputs, baselines.py:21-26).
template <typename In> struct Arith;
template <> struct Arith<float> {
  typedef float C;
  static __device__ __forceinline__ float norm(float x, float g, bool scaled) {
    return scaled ? div_level(x, g) : x;
  }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
};
template <> struct Arith<uint8_t> {
  typedef double C;
  static __device__ __forceinline__ double norm(uint8_t x, uint8_t g, bool scaled) {
    return scaled ? __ddiv_rn((double)x, (double)g) : (double)x;
  }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
};


// Pairwise sum in the arithmetic type C itself (np.sum of a float32 array
// accumulates in float32).
template <typename A, typename F>
__device__ __forceinline__ typename A::C pairwise_sum_t(const PwProg& prog, F term) {
  typedef typename A::C C;
  C stack[8];
  int sp = 0;
  for (int op = 0; op < prog.n_ops; ++op) {
    const int len = prog.len[op];
    if (len == 0) {
      --sp;
      stack[sp - 1] = A::add(stack[sp - 1], stack[sp]);
      continue;
    }
    C res;
    if (len < 8) {
      res = C(0);
      for (int k = 0; k < len; ++k) res = A::add(res, term());
    } else {
      C r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = term();
      const int body = len - len % 8;
      for (int i = 8; i < body; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = A::add(r[k], term());
      }
      res = A::add(A::add(A::add(r[0], r[1]), A::add(r[2], r[3])),
                   A::add(A::add(r[4], r[5]), A::add(r[6], r[7])));
      for (int i = body; i < len; ++i) res = A::add(res, term());
    }
    stack[sp++] = res;
  }
  return stack[0];
}

}  // namespace srl
