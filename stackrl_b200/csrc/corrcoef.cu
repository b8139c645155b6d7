// baselines.corrcoef(localized=True) (reference: stackrl/baselines.py:79-114,
// the masked Python-loop variant):
//
//   m      = n > 0,  c = count(m)
//   n'     = n - sum(where(m, n, 0)) / c ;  n_var = sum(where(m, n'^2, 0))
//   per position:  o' = win - sum(where(m, win, 0)) / c
//                  o_var = sum(where(m, o'^2, 0))
//                  f = o_var != 0 ? sum(where(m, n' * o', 0)) / sqrt(n_var * o_var) : 0
//
// Every sum is np.sum over a contiguous [h,h] block, i.e. numpy's pairwise order
// (pairwise.cuh) in the array's own type.  For uint8 observations everything is
// float64 (uint8 / uint8 is float64 in get_inputs).  For float32 observations
// numpy's promotion rules make it mixed: `c` is a numpy integer, so both means
// are float64; `n -= mean` rounds the float64 difference back to float32 (n'
// and n_var stay float32), while `win - mean` is a float64 array, hence o_var,
// the covariance sum, the product n_var * o_var, sqrt and the quotient are all
// float64.  Bit-exact for both dtypes.
#include "common.cuh"
#include "kernels.h"
#include "pairwise.cuh"

namespace srl {

namespace {

template <typename C> __device__ __forceinline__ C sqrt_rn(C x);
template <> __device__ __forceinline__ float sqrt_rn<float>(float x) { return __fsqrt_rn(x); }
template <> __device__ __forceinline__ double sqrt_rn<double>(double x) { return __dsqrt_rn(x); }
template <typename C> __device__ __forceinline__ C div_rn(C a, C b);
template <> __device__ __forceinline__ float div_rn<float>(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double div_rn<double>(double a, double b) { return __ddiv_rn(a, b); }

// One thread per rock: centred rock n' (masked cells keep n - mean too, they are
// never summed), stats = (n_var, count).
template <typename In>
__global__ void __launch_bounds__(128)
corrcoef_rock_kernel(const In* __restrict__ rocks, const In* __restrict__ level,
                     typename Arith<In>::C* __restrict__ centred,
                     typename Arith<In>::C* __restrict__ stats, int nrocks, int R, int h,
                     const PwProg prog) {
  typedef Arith<In> A;
  typedef typename A::C C;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrocks) return;
  const In* rock = rocks + (size_t)idx * h * h;
  C* out = centred + (size_t)idx * h * h;
  const bool scaled = level != nullptr;
  const In g = scaled ? level[idx / R] : In(1);
  int k = 0, count = 0;
  auto term1 = [&]() {
    const C n = A::norm(rock[k++], g, scaled);
    if (n > C(0)) {
      ++count;
      return n;
    }
    return C(0);
  };
  const C total = pairwise_sum_t<A>(prog, term1);
  const double mean = __ddiv_rn((double)total, (double)count);
  k = 0;
  auto term2 = [&]() {
    const C n = A::norm(rock[k], g, scaled);
    const C c = (C)__dsub_rn((double)n, mean);     // n -= mean (in place: back to C)
    out[k++] = c;
    return n > C(0) ? A::mul(c, c) : C(0);
  };
  const C n_var = pairwise_sum_t<A>(prog, term2);
  stats[2 * (size_t)idx] = n_var;
  stats[2 * (size_t)idx + 1] = (C)count;
}

template <typename In>
__global__ void __launch_bounds__(128)
corrcoef_localized_kernel(const In* __restrict__ walls, const In* __restrict__ rocks,
                          const In* __restrict__ level,
                          const typename Arith<In>::C* __restrict__ centred,
                          const typename Arith<In>::C* __restrict__ stats,
                          double* __restrict__ out, int R, int H, int W, int h, int band,
                          int nbands, const PwProg prog) {
  typedef Arith<In> A;
  typedef typename A::C C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Ph = H - h + 1, Pw = W - h + 1;
  int b = blockIdx.x;
  const int bandi = b % nbands; b /= nbands;
  const int r = b % R;
  const int e = b / R;
  const int i0 = bandi * band;
  const int rows_out = min(band, Ph - i0);
  const int rows_in = rows_out + h - 1;
  C* rock_s = reinterpret_cast<C*>(smem_raw);                   // [h*h] centred n'
  C* wall_s = rock_s + h * h;                                   // [rows_in][W]
  unsigned char* live_s = reinterpret_cast<unsigned char*>(wall_s + rows_in * W);   // [h*h]
  const int tid = threadIdx.x;
  const bool scaled = level != nullptr;
  const In g = scaled ? level[e] : In(1);
  const In* wall = walls + ((size_t)e * H + i0) * W;
  for (int k = tid; k < rows_in * W; k += blockDim.x) wall_s[k] = A::norm(wall[k], g, scaled);
  const size_t ridx = (size_t)e * R + r;
  for (int k = tid; k < h * h; k += blockDim.x) {
    rock_s[k] = centred[ridx * h * h + k];
    live_s[k] = A::norm(rocks[ridx * h * h + k], g, scaled) > C(0) ? 1 : 0;
  }
  __syncthreads();
  const C n_var = stats[2 * ridx], count = stats[2 * ridx + 1];
  for (int item = tid; item < rows_out * Pw; item += blockDim.x) {
    const int i = item / Pw, j = item % Pw;
    const size_t o = ((ridx * Ph) + i0 + i) * Pw + j;
    if (n_var == C(0)) {          // baselines.py:96-98: everything is zero
      out[o] = 0.;
      continue;
    }
    const C* win = wall_s + i * W + j;
    int u = 0, v = 0;
    auto next = [&]() { if (++v == h) { v = 0; ++u; } };
    auto t_sum = [&]() {
      const C x = live_s[u * h + v] ? win[u * W + v] : C(0);
      next();
      return x;
    };
    typedef Arith<uint8_t> D;     // float64 arithmetic
    const double mean =
        __ddiv_rn((double)pairwise_sum_t<A>(prog, t_sum), (double)count);
    u = v = 0;
    auto t_var = [&]() {
      const double d = __dsub_rn((double)win[u * W + v], mean);
      const double x = live_s[u * h + v] ? __dmul_rn(d, d) : 0.;
      next();
      return x;
    };
    const double o_var = pairwise_sum_t<D>(prog, t_var);
    double f = 0.;
    if (o_var != 0.) {
      u = v = 0;
      auto t_cov = [&]() {
        const double d = __dsub_rn((double)win[u * W + v], mean);
        const double x = live_s[u * h + v] ? __dmul_rn((double)rock_s[u * h + v], d) : 0.;
        next();
        return x;
      };
      const double cov = pairwise_sum_t<D>(prog, t_cov);
      f = __ddiv_rn(cov, __dsqrt_rn(__dmul_rn((double)n_var, o_var)));
    }
    out[o] = f;
  }
}

template <typename In>
int corrcoef_localized_t(const In* walls, const In* rocks, const In* level, void* work,
                         double* out, int E, int R, int H, int W, int h,
                         cudaStream_t stream) {
  typedef typename Arith<In>::C C;
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h, SRL_E_INVALID,
              "corrcoef_localized: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && rocks && work && out, SRL_E_INVALID,
              "corrcoef_localized: null pointer");
  SRL_REQUIRE(h <= 64, SRL_E_UNSUPPORTED, "corrcoef_localized: rock side %d > 64", h);
  const int sms = sm_count();
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "corrcoef_localized: no CUDA device");
  C* centred = reinterpret_cast<C*>(work);
  C* stats = centred + (size_t)E * R * h * h;
  PwProg prog;
  prog.n_ops = 0;
  build_prog(0, h * h, prog);
  const int n = E * R;
  corrcoef_rock_kernel<In><<<(n + 127) / 128, 128, 0, stream>>>(rocks, level, centred, stats,
                                                               n, R, h, prog);
  int rc = check_launch("corrcoef_rock_kernel");
  if (rc != SRL_OK) return rc;
  const int Ph = H - h + 1;
  auto smem_for = [&](int band) {
    return sizeof(C) * ((size_t)h * h + (size_t)(band + h - 1) * W) + (size_t)h * h;
  };
  int band = Ph;
  while (band > 1 && ((size_t)E * R * ((Ph + band - 1) / band) < (size_t)2 * sms ||
                      smem_for(band) > 100 * 1024))
    band = (band + 1) / 2;
  SRL_REQUIRE(smem_for(band) <= 220 * 1024, SRL_E_UNSUPPORTED,
              "corrcoef_localized: %d-column wall rows with a %d-row rock exceed shared "
              "memory", W, h);
  const int nbands = (Ph + band - 1) / band;
  const size_t smem = smem_for(band);
  SRL_CUDA(cudaFuncSetAttribute(corrcoef_localized_kernel<In>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  corrcoef_localized_kernel<In><<<E * R * nbands, 128, smem, stream>>>(
      walls, rocks, level, centred, stats, out, R, H, W, h, band, nbands, prog);
  return check_launch("corrcoef_localized_kernel");
}

}  // namespace

int corrcoef_localized_f32(const float* walls, const float* rocks, const float* level,
                           void* work, double* out, int E, int R, int H, int W, int h,
                           cudaStream_t stream) {
  return corrcoef_localized_t<float>(walls, rocks, level, work, out, E, R, H, W, h, stream);
}
int corrcoef_localized_u8(const uint8_t* walls, const uint8_t* rocks, const uint8_t* level,
                          void* work, double* out, int E, int R, int H, int W, int h,
                          cudaStream_t stream) {
  return corrcoef_localized_t<uint8_t>(walls, rocks, level, work, out, E, R, H, W, h,
                                       stream);
}

}  // namespace srl
