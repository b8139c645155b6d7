// Goal-masked, local-minimum-preferring arg-min and the negated value map
// (reference: Baseline.call, stackrl/baselines.py:201-217) plus PyGreedy's
// batchwise pick over the N views of one environment
// (stackrl/agents/policies.py:63-80), and Observer.pose's single-position drop
// height (stackrl/envs/stack/observer.py:401-409).
//
// One CTA per environment walks its R maps; every reduction keeps numpy's
// first-index tie-break (np.argmin / np.argmax), so action indices are exact.
#include <math_constants.h>

#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace {

constexpr int kSelThreads = 256;

struct Best {          // arg-min candidate: smaller value wins, then smaller index
  double v;
  int idx;
};
__device__ __forceinline__ Best better(Best a, Best b) {
  if (b.idx < 0) return a;
  if (a.idx < 0) return b;
  if (b.v < a.v || (b.v == a.v && b.idx < a.idx)) return b;
  return a;
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.idx = __shfl_xor_sync(0xffffffffu, x.idx, o);
    x = better(x, y);
  }
  return x;
}
__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ int warp_max(int x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// One CTA per environment, one WARP per map: the R views of an environment are
// dealt round-robin to the warps, every reduction inside a map is a shuffle
// reduction (no block barrier), and the batch-wise pick over the views is one
// small shared-memory reduction at the end.
template <typename V>
__global__ void __launch_bounds__(kSelThreads)
select_kernel(const V* __restrict__ values, const int32_t* __restrict__ counts,
              int64_t* __restrict__ actions, double* __restrict__ shown,
              int64_t* __restrict__ best, int R, int Ph, int Pw, int minorder,
              double overlap_threshold) {
  constexpr int NW = kSelThreads / 32;
  __shared__ double s_score[NW];
  __shared__ int s_view[NW];
  __shared__ int s_action[NW];

  const int e = blockIdx.x;
  const int P = Ph * Pw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Running batch-wise pick of this warp (views visited in increasing order, so
  // a strict > keeps numpy.argmax's first-index rule).
  double my_score = 0.;
  int my_view = -1, my_action = 0;

  for (int r = warp; r < R; r += NW) {
    const V* v = values + ((size_t)e * R + r) * P;
    const int32_t* c = counts ? counts + ((size_t)e * R + r) * P : nullptr;
    double* sh = shown ? shown + ((size_t)e * R + r) * P : nullptr;

    // ---- pass 1: overlap-count maximum -> mask cut (baselines.py:155-156) ---- //
    double cut = 0.;
    if (c) {
      int cm = 0;
      for (int k = lane; k < P; k += 32) cm = max(cm, __ldg(c + k));
      cut = overlap_threshold * (double)warp_max(cm);      // float64, like numpy
    }

    // ---- pass 2: masked maximum, local minima, both arg-min candidates ------- //
    double vm = -CUDART_INF;
    Best bmin = {0., -1};      // over mask & local minimum
    Best bmask = {0., -1};     // over mask (or everything when goal=False)
    for (int k = lane; k < P; k += 32) {
      const bool in = c ? ((double)__ldg(c + k) >= cut) : true;
      if (!in) continue;
      const double x = (double)__ldg(v + k);
      vm = fmax(vm, x);
      Best cand = {x, k};
      bmask = better(bmask, cand);
      if (c && minorder > 0) {
        // minimum_filter(size=1+2m, mode='constant', cval=0) == values
        // (baselines.py:209): x <= every in-bounds neighbour, and x <= 0 if the
        // window leaves the map (quirk Q6).
        const int i = k / Pw, j = k - i * Pw;
        bool low = true;
        if (i < minorder || j < minorder || i + minorder >= Ph || j + minorder >= Pw)
          low = x <= 0.;
        for (int di = -minorder; low && di <= minorder; ++di) {
          const int ii = i + di;
          if (ii < 0 || ii >= Ph) continue;
          for (int dj = -minorder; dj <= minorder; ++dj) {
            const int jj = j + dj;
            if (jj < 0 || jj >= Pw) continue;
            if ((double)__ldg(v + ii * Pw + jj) < x) {
              low = false;
              break;
            }
          }
        }
        if (low) bmin = better(bmin, cand);
      }
    }
    vm = warp_max(vm);
    bmin = warp_best(bmin);
    bmask = warp_best(bmask);
    const Best pick = bmin.idx >= 0 ? bmin : bmask;
    if (lane == 0) actions[(size_t)e * R + r] = pick.idx;
    // PyGreedy batchwise: first argmax over views of shown[action] = -value.
    if (my_view < 0 || -pick.v > my_score) {
      my_score = -pick.v;
      my_view = r;
      my_action = pick.idx;
    }

    // ---- pass 3: negated value map (baselines.py:213, :215, :217) ------------- //
    if (sh) {
      const double fill = vm + 0.001;
      for (int k = lane; k < P; k += 32) {
        const bool in = c ? ((double)__ldg(c + k) >= cut) : true;
        sh[k] = -(in ? (double)__ldg(v + k) : fill);
      }
    }
  }
  if (!best) return;
  if (lane == 0) {
    s_score[warp] = my_score;
    s_view[warp] = my_view;
    s_action[warp] = my_action;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int bv = -1, ba = 0;
    double bs = 0.;
    for (int w = 0; w < NW; ++w) {
      if (s_view[w] < 0) continue;
      if (bv < 0 || s_score[w] > bs || (s_score[w] == bs && s_view[w] < bv)) {
        bs = s_score[w];
        bv = s_view[w];
        ba = s_action[w];
      }
    }
    best[2 * (size_t)e] = bv;
    best[2 * (size_t)e + 1] = ba;
  }
}

template <typename V>
int launch_select(const V* values, const int32_t* counts, int64_t* actions,
                  double* shown, int64_t* best, int E, int R, int Ph, int Pw,
                  int minorder, double overlap_threshold, cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && Ph >= 1 && Pw >= 1 && minorder >= 0, SRL_E_INVALID,
              "select: bad shape E=%d R=%d Ph=%d Pw=%d minorder=%d", E, R, Ph, Pw,
              minorder);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(values && actions, SRL_E_INVALID, "select: null pointer");
  select_kernel<V><<<E, kSelThreads, 0, stream>>>(values, counts, actions, shown, best,
                                                   R, Ph, Pw, minorder,
                                                   overlap_threshold);
  return check_launch("select_kernel");
}

// One warp per environment: max over the live cells of (window + rock).
__global__ void __launch_bounds__(128)
drop_height_kernel(const float* __restrict__ walls, const float* __restrict__ rocks,
                   const int32_t* __restrict__ picks, float* __restrict__ out, int E,
                   int R, int H, int W, int h, float threshold) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= E) return;
  const int lane = threadIdx.x & 31;
  const int r = picks[3 * e], i = picks[3 * e + 1], j = picks[3 * e + 2];
  const float* wall = walls + (size_t)e * H * W + (size_t)i * W + j;
  const float* rock = rocks + ((size_t)e * R + r) * h * h;
  float m = kNegInf;
  for (int k = lane; k < h * h; k += 32) {
    const float n = rock[k];
    if (n > threshold) m = fmaxf(m, wall[(k / h) * W + k % h] + n);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[e] = m;
}

}  // namespace

int select_f32(const float* values, const int32_t* counts, int64_t* actions,
               double* shown, int64_t* best, int E, int R, int Ph, int Pw, int minorder,
               double overlap_threshold, cudaStream_t stream) {
  return launch_select<float>(values, counts, actions, shown, best, E, R, Ph, Pw,
                              minorder, overlap_threshold, stream);
}

int select_f64(const double* values, const int32_t* counts, int64_t* actions,
               double* shown, int64_t* best, int E, int R, int Ph, int Pw, int minorder,
               double overlap_threshold, cudaStream_t stream) {
  return launch_select<double>(values, counts, actions, shown, best, E, R, Ph, Pw,
                               minorder, overlap_threshold, stream);
}

int drop_height_f32(const float* walls, const float* rocks, const int32_t* picks,
                    float* out, int E, int R, int H, int W, int h, float threshold,
                    cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h, SRL_E_INVALID,
              "drop_height: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && rocks && picks && out, SRL_E_INVALID, "drop_height: null pointer");
  const int warps = 4;
  drop_height_kernel<<<(E + warps - 1) / warps, warps * 32, 0, stream>>>(
      walls, rocks, picks, out, E, R, H, W, h, threshold);
  return check_launch("drop_height_kernel");
}

}  // namespace srl
