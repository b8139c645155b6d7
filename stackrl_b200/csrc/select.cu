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
constexpr int kMaxViews = 64;      // views per environment the packed kernel keeps results for

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

// --------------------------------------------------------------------------- //
// goal_overlap + select in one launch ("mask_select"): the overlap counts of
// baselines.py:152-155 are produced and consumed in shared memory (bit-packed
// `wall < goal` windows, several rows per 32-bit word when the rock is small,
// AND+POPC), so they never travel through HBM.  One CTA per environment; the
// bit images are built by the whole block, then every warp owns whole views.
// --------------------------------------------------------------------------- //
// Cooperative global -> shared copy with every load independent (several 16-B
// loads in flight per thread); falls back to bytes when the source is unaligned.
__device__ __forceinline__ void stage_bytes(void* dst, const void* src, size_t nbytes,
                                            int tid, int nthreads) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    const int n4 = (int)(nbytes >> 4);
#pragma unroll 4
    for (int k = tid; k < n4; k += nthreads) d4[k] = __ldg(s4 + k);
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(src);
    unsigned char* db = reinterpret_cast<unsigned char*>(dst);
    for (size_t k = ((size_t)n4 << 4) + tid; k < nbytes; k += nthreads) db[k] = sb[k];
  } else {
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(src);
    unsigned char* db = reinterpret_cast<unsigned char*>(dst);
#pragma unroll 4
    for (size_t k = tid; k < nbytes; k += nthreads) db[k] = sb[k];
  }
}

// Copy `count` elements with the destination at the source's 16-byte phase
// (dst = base16 + (src & 15)): head and tail element-wise, the body as 16-byte vectors
// whatever the alignment of src (a [E, P] map with odd P starts 4, 8 or 12 bytes off
// for three environments in four).  base16 must have 16 spare bytes.  Returns dst.
// kAsync: the body travels as cp.async copies that the caller waits for (stage_wait(), then
// a barrier) before the first read -- the copy overlaps whatever the block does until then.
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}
template <typename T, bool kAsync = false>
__device__ __forceinline__ T* stage_phased(void* base16, const T* src, size_t count, int tid,
                                           int nthreads) {
  const int phase = (int)(reinterpret_cast<uintptr_t>(src) & 15);
  T* dst = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(base16) + phase);
  size_t head = (size_t)((16 - phase) & 15) / sizeof(T);
  if (head > count) head = count;
  const size_t n4 = ((count - head) * sizeof(T)) >> 4;
  const size_t body = n4 * (16 / sizeof(T));
  for (size_t k = tid; k < head; k += nthreads) dst[k] = src[k];
  const int4* s4 = reinterpret_cast<const int4*>(src + head);
  int4* d4 = reinterpret_cast<int4*>(dst + head);
  if constexpr (kAsync) {
    for (int k = tid; k < (int)n4; k += nthreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(d4 + k)),
                   "l"(s4 + k)
                   : "memory");
  } else {
#pragma unroll 4
    for (int k = tid; k < (int)n4; k += nthreads) d4[k] = __ldg(s4 + k);
  }
  for (size_t k = head + body + tid; k < count; k += nthreads) dst[k] = src[k];
  return dst;
}

struct MaskSelectParams {
  int staged;                  // inputs + score maps of one environment fit shared memory
  int R, H, W, h, Ph, Pw, minorder;
  int nW, pf, hb, ng;          // words per bit row (+1 zero word); row packing
  double overlap_threshold;
  uint32_t mulPw;              // ceil(2^32 / Pw) for k / Pw (k * Pw < 2^32)
  int vec4;                    // packed kernel: bit images straight from 4-wide global loads
  int stage_values;            // packed kernel: score maps staged in shared memory (small maps)
  int rch;                     // packed kernel: views per chunk (power of two dividing R, <= 32)
  uint32_t mulW4, mulh4, mulhh4;   // ceil(2^32 / d) for d = W/4, h/4, h*h/4
};

// Four consecutive observation values (one 16-B or 4-B load).
__device__ __forceinline__ void load4(const float* p, float (&x)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
__device__ __forceinline__ void load4(const uint8_t* p, uint8_t (&x)[4]) {
  const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(p));
  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
__device__ __forceinline__ uint32_t udiv_mul(uint32_t k, uint32_t mul) {
  return mul == 0u ? k : __umulhi(k, mul);
}

// Arg-min candidate in the score type itself (float compares are exact; only
// the returned value map is float64).
template <typename V>
struct Cand {
  V v;
  int idx;
};
template <typename V>
__device__ __forceinline__ void take(Cand<V>& a, V v, int idx) {
  if (a.idx < 0 || v < a.v || (v == a.v && idx < a.idx)) {
    a.v = v;
    a.idx = idx;
  }
}
template <typename V>
__device__ __forceinline__ Cand<V> warp_cand(Cand<V> x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const V v = __shfl_xor_sync(0xffffffffu, x.v, o);
    const int i = __shfl_xor_sync(0xffffffffu, x.idx, o);
    if (i >= 0) take(x, v, i);
  }
  return x;
}
template <typename V>
__device__ __forceinline__ V warp_vmax(V x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const V y = __shfl_xor_sync(0xffffffffu, x, o);
    x = y > x ? y : x;
  }
  return x;
}

// Is v[i,j] a minimum of its zero-padded (2M+1)^2 window (baselines.py:209)?
template <typename V, int M>
__device__ __forceinline__ bool local_min(const V* v, V x, int i, int j, int Ph, int Pw,
                                          int m_runtime) {
  const int m = M >= 0 ? M : m_runtime;
  if (i >= m && j >= m && i + m < Ph && j + m < Pw) {
    const V* c = v + i * Pw + j;
    if (M == 1) {
      return !(c[-Pw - 1] < x) && !(c[-Pw] < x) && !(c[-Pw + 1] < x) && !(c[-1] < x) &&
             !(c[1] < x) && !(c[Pw - 1] < x) && !(c[Pw] < x) && !(c[Pw + 1] < x);
    }
    for (int di = -m; di <= m; ++di)
      for (int dj = -m; dj <= m; ++dj)
        if (c[di * Pw + dj] < x) return false;
    return true;
  }
  // window leaves the map: the padding contributes 0 (quirk Q6)
  if (!(x <= V(0))) return false;
  for (int di = -m; di <= m; ++di) {
    const int ii = i + di;
    if (ii < 0 || ii >= Ph) continue;
    for (int dj = -m; dj <= m; ++dj) {
      const int jj = j + dj;
      if (jj < 0 || jj >= Pw) continue;
      if (v[ii * Pw + jj] < x) return false;
    }
  }
  return true;
}

template <typename V, typename In, bool STAGED, int M>
__global__ void __launch_bounds__(kSelThreads)
mask_select_kernel(const V* __restrict__ values, const In* __restrict__ walls,
                   const In* __restrict__ goals, const In* __restrict__ rocks,
                   int64_t* __restrict__ actions, double* __restrict__ shown,
                   int64_t* __restrict__ best, const MaskSelectParams q) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  constexpr int NW = kSelThreads / 32;
  __shared__ V s_score[NW];
  __shared__ int s_view[NW];
  __shared__ int s_action[NW];
  __shared__ int s_cm[NW];
  __shared__ Cand<V> s_bmin[NW], s_bmask[NW];
  __shared__ V s_vm[NW];
  __shared__ bool s_any[NW];
  __shared__ double s_fill[NW];
  const int R = q.R, H = q.H, W = q.W, h = q.h, Ph = q.Ph, Pw = q.Pw, P = Ph * Pw;
  uint32_t* below = reinterpret_cast<uint32_t*>(sel_smem);          // [H][nW]
  uint32_t* foot = below + H * q.nW;                                // [R][ng] row-packed
  uint32_t* win = foot + R * q.ng;                                  // [H][Pw] row-packed
  uint16_t* cnt = reinterpret_cast<uint16_t*>(win + H * Pw);        // [NW][P]

  const int e = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t hmask = h >= 32 ? 0xffffffffu : ((1u << h) - 1u);

  const In* wall = walls + (size_t)e * H * W;
  const In* goal = goals + (size_t)e * H * W;
  const In* rock = rocks + (size_t)e * R * h * h;
  const V* vals = values + (size_t)e * R * P;
  if constexpr (STAGED) {
    // One latency round trip for everything this environment needs.
    unsigned char* at = reinterpret_cast<unsigned char*>(cnt + (size_t)NW * P);
    at += (16 - (reinterpret_cast<uintptr_t>(at) & 15)) & 15;
    V* v_s = reinterpret_cast<V*>(at);
    at += ((size_t)R * P * sizeof(V) + 15) & ~(size_t)15;
    In* wall_s = reinterpret_cast<In*>(at);
    at += ((size_t)H * W * sizeof(In) + 15) & ~(size_t)15;
    In* goal_s = reinterpret_cast<In*>(at);
    at += ((size_t)H * W * sizeof(In) + 15) & ~(size_t)15;
    In* rock_s = reinterpret_cast<In*>(at);
    stage_bytes(v_s, vals, (size_t)R * P * sizeof(V), tid, kSelThreads);
    stage_bytes(wall_s, wall, (size_t)H * W * sizeof(In), tid, kSelThreads);
    stage_bytes(goal_s, goal, (size_t)H * W * sizeof(In), tid, kSelThreads);
    stage_bytes(rock_s, rock, (size_t)R * h * h * sizeof(In), tid, kSelThreads);
    __syncthreads();
    vals = v_s;
    wall = wall_s;
    goal = goal_s;
    rock = rock_s;
  }

  // ---- bit images of the raw observation (baselines.py:153-154) -------------- //
  for (int row = warp; row < H; row += NW) {
    for (int word = 0; word < q.nW; ++word) {
      const int col = word * 32 + lane;
      bool b = false;
      if (col < W) b = wall[row * W + col] < goal[row * W + col];
      const uint32_t bits = __ballot_sync(0xffffffffu, b);
      if (lane == 0) below[row * q.nW + word] = bits;
    }
  }
  for (int r = warp; r < R; r += NW) {
    for (int grp = 0; grp < q.ng; ++grp) {
      uint32_t packed = 0;
      for (int s = 0; s < q.pf; ++s) {
        const int u = grp * q.pf + s;
        const bool b = u < h && lane < h && rock[(r * h + u) * h + lane] > In(0);
        packed |= (__ballot_sync(0xffffffffu, b) & hmask) << (s * q.hb);
      }
      if (lane == 0) foot[r * q.ng + grp] = packed;
    }
  }
  __syncthreads();
  for (int k = tid; k < H * Pw; k += kSelThreads) {
    const int row = __umulhi((uint32_t)k, q.mulPw), j = k - row * Pw;
    uint32_t packed = 0;
    for (int s = 0; s < q.pf && row + s < H; ++s) {
      const uint32_t* b = below + (row + s) * q.nW + (j >> 5);
      packed |= (__funnelshift_r(b[0], b[1], j & 31) & hmask) << (s * q.hb);
    }
    win[k] = packed;
  }
  __syncthreads();

  // ---- per view: counts -> cut -> candidates ------------------------------------- //
  // `wpv` warps cooperate on one view (all 8 when there is a single view, one
  // each when there are >= 8), so small-R batches still use the whole block.
  int wpv = 1;
  while (wpv * 2 * R <= NW) wpv *= 2;
  const int vpr = NW / wpv;                 // views per round
  const int sub = warp % wpv, slot = warp / wpv;
  const int gstride = q.pf * Pw;
  uint16_t* mine = cnt + (size_t)slot * P;
  V my_score = V(0);                        // batch-wise pick, kept by the slot leader
  int my_view = -1, my_action = 0;
  for (int r0 = 0; r0 < R; r0 += vpr) {
    const int r = r0 + slot;
    const bool active = r < R;
    const V* v = vals + (size_t)(active ? r : 0) * P;
    if (active) {
      const uint32_t* fp = foot + r * q.ng;
      int cm = 0;
      for (int k = sub * 32 + lane; k < P; k += 32 * wpv) {
        const uint32_t* wp = win + k;                       // (i*Pw + j) == k
        int c = 0;
#pragma unroll 4
        for (int g = 0; g < q.ng; ++g) c += __popc(wp[g * gstride] & fp[g]);
        mine[k] = (uint16_t)c;
        cm = max(cm, c);
      }
      cm = warp_max(cm);
      if (lane == 0) s_cm[warp] = cm;
    }
    __syncthreads();
    int cmin = 0;
    Cand<V> bmin = {V(0), -1}, bmask = {V(0), -1};
    V vm = V(0);
    bool any = false;
    if (active) {
      int cm = 0;
      for (int w = 0; w < wpv; ++w) cm = max(cm, s_cm[slot * wpv + w]);
      // count >= threshold*max  <=>  count >= ceil(threshold*max) for integer counts
      cmin = (int)ceil(q.overlap_threshold * (double)cm);
      for (int k = sub * 32 + lane; k < P; k += 32 * wpv) {
        if ((int)mine[k] < cmin) continue;
        const V x = v[k];
        vm = (!any || x > vm) ? x : vm;
        any = true;
        take(bmask, x, k);
        if (M != 0) {
          const int i = __umulhi((uint32_t)k, q.mulPw), j = k - i * Pw;
          if (local_min<V, M>(v, x, i, j, Ph, Pw, q.minorder)) take(bmin, x, k);
        }
      }
      bmin = warp_cand(bmin);
      bmask = warp_cand(bmask);
      // masked maximum over the lanes that saw a masked cell
      const unsigned who = __ballot_sync(0xffffffffu, any);
      if (who) {
        const V seed = __shfl_sync(0xffffffffu, vm, __ffs(who) - 1);
        vm = warp_vmax(any ? vm : seed);
      }
      if (lane == 0) {
        s_bmin[warp] = bmin;
        s_bmask[warp] = bmask;
        s_vm[warp] = vm;
        s_any[warp] = who != 0;
      }
    }
    __syncthreads();
    if (active && sub == 0 && lane == 0) {
      Cand<V> a = {V(0), -1}, b = {V(0), -1};
      V m2 = V(0);
      bool have = false;
      for (int w = 0; w < wpv; ++w) {
        const int ww = slot * wpv + w;
        if (s_bmin[ww].idx >= 0) take(a, s_bmin[ww].v, s_bmin[ww].idx);
        if (s_bmask[ww].idx >= 0) take(b, s_bmask[ww].v, s_bmask[ww].idx);
        if (s_any[ww]) {
          m2 = (!have || s_vm[ww] > m2) ? s_vm[ww] : m2;
          have = true;
        }
      }
      const Cand<V> pick = a.idx >= 0 ? a : b;
      actions[(size_t)e * R + r] = pick.idx;
      // PyGreedy batchwise: first argmax over views of -value (strict <: first wins)
      if (my_view < 0 || pick.v < my_score) {
        my_score = pick.v;
        my_view = r;
        my_action = pick.idx;
      }
      s_fill[slot] = (double)m2 + 0.001;
    }
    if (shown) {
      __syncthreads();
      if (active) {
        const double fill = s_fill[slot];
        double* sh = shown + ((size_t)e * R + r) * P;
        for (int k = sub * 32 + lane; k < P; k += 32 * wpv)
          sh[k] = -((int)mine[k] >= cmin ? (double)v[k] : fill);
      }
    }
    __syncthreads();
  }
  if (!best) return;
  if (sub == 0 && lane == 0) {
    s_score[slot] = my_score;
    s_view[slot] = my_view;
    s_action[slot] = my_action;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int bv = -1, ba = 0;
    V bs = V(0);
    for (int w = 0; w < vpr; ++w) {
      if (s_view[w] < 0) continue;
      if (bv < 0 || s_score[w] < bs || (s_score[w] == bs && s_view[w] < bv)) {
        bs = s_score[w];
        bv = s_view[w];
        ba = s_action[w];
      }
    }
    best[2 * (size_t)e] = bv;
    best[2 * (size_t)e + 1] = ba;
  }
}

// Same computation with lanes mapped to (view, position) pairs: R (a power of
// two <= 32) consecutive lanes share one window word and each keeps the packed
// footprint of ITS view in registers, so an overlap step is LDS + LOP3 + POPC +
// IADD with no footprint traffic, and all eight warps work on all views at once.
// GEO = (H << 20 | W << 10 | h) << 6 | R fixes the geometry at compile time (0:
// run-time geometry from the parameters): all index arithmetic folds.
// Order-preserving float -> uint key (-0 and +0 share a key) and back.
__device__ __forceinline__ uint32_t ord_key(float v) {
  const uint32_t b = __float_as_uint(__fadd_rn(v, 0.f));
  return b ^ ((b >> 31) != 0u ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ uint32_t ord_key(double) { return 0u; }      // (float maps only)
__device__ __forceinline__ float ord_val(uint32_t k) {
  return __uint_as_float(k ^ ((k >> 31) != 0u ? 0x80000000u : 0xffffffffu));
}
// Arg-min candidate of the whole warp: smallest value, then smallest index (what the
// shuffle rounds of take() give), in every lane.
template <typename V>
__device__ __forceinline__ void warp_cand_redux(Cand<V>& c) {
  const uint32_t key = c.idx >= 0 ? ord_key(c.v) : 0xffffffffu;
  const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
  const uint32_t imin = __reduce_min_sync(
      0xffffffffu, (c.idx >= 0 && key == kmin) ? (uint32_t)c.idx : 0x7fffffffu);
  c.idx = imin == 0x7fffffffu ? -1 : (int)imin;
  c.v = (V)ord_val(kmin);
}

template <typename V, typename In, int M, int NG, long long GEO = 0>
__global__ void __launch_bounds__(kSelThreads)
mask_select_packed_kernel(const V* __restrict__ values, const In* __restrict__ walls,
                          const In* __restrict__ goals, const In* __restrict__ rocks,
                          int64_t* __restrict__ actions, double* __restrict__ shown,
                          int64_t* __restrict__ best, const MaskSelectParams q) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  constexpr int NW = kSelThreads / 32;
  __shared__ int s_cmax[32];
  __shared__ Cand<V> s_bmin[NW][32], s_bmask[NW][32];
  __shared__ V s_vm[NW][32];
  __shared__ bool s_any[NW][32];
  __shared__ V s_pick[kMaxViews];
  __shared__ int s_pick_i[kMaxViews];
  __shared__ double s_fill[32];
  constexpr bool kFixed = GEO != 0;
  const int R = kFixed ? (int)(GEO & 63) : q.R, H = kFixed ? (int)(GEO >> 26) : q.H;
  const int W = kFixed ? (int)((GEO >> 16) & 1023) : q.W;
  const int h = kFixed ? (int)((GEO >> 6) & 1023) : q.h;
  const int Ph = H - h + 1, Pw = W - h + 1, P = Ph * Pw;
  const int g_nW = kFixed ? (W + 31) / 32 + 1 : q.nW;
  const int g_pf = kFixed ? (h > 16 ? 1 : (h > 8 ? 2 : 4)) : q.pf;
  const int g_hb = kFixed ? 32 / g_pf : q.hb;
  const int g_ng = kFixed ? (h + g_pf - 1) / g_pf : q.ng;
  uint32_t* below = reinterpret_cast<uint32_t*>(sel_smem);          // [H][nW]
  uint32_t* foot = below + H * g_nW;                                // [R][ng] row-packed
  uint32_t* win = foot + R * g_ng;                                  // [H][Pw] row-packed
  uint16_t* cnt = reinterpret_cast<uint16_t*>(win + H * Pw);        // [views per chunk][P]
  unsigned char* at =
      reinterpret_cast<unsigned char*>(cnt + (size_t)(kFixed ? R : q.rch) * P);
  at += (16 - (reinterpret_cast<uintptr_t>(at) & 15)) & 15;
  V* vals = reinterpret_cast<V*>(at);
  at += ((size_t)R * P * sizeof(V) + 15) & ~(size_t)15;
  In* wall = reinterpret_cast<In*>(at);
  at += ((size_t)H * W * sizeof(In) + 15) & ~(size_t)15;
  In* goal = reinterpret_cast<In*>(at);
  at += ((size_t)H * W * sizeof(In) + 15) & ~(size_t)15;
  In* rock = reinterpret_cast<In*>(at);

  const int e = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t hmask = h >= 32 ? 0xffffffffu : ((1u << h) - 1u);
  // pf (rows packed per word) is 1, 2 or 4: shifts and masks, no divisions
  const int lp = g_pf == 4 ? 2 : (g_pf == 2 ? 1 : 0);
  if (kFixed || q.vec4) {
    // ---- bit images straight from global memory (baselines.py:153-154) --------- //
    // Every thread turns 4 consecutive pixels into 4 bits and ORs them into the
    // packed word; only the score maps are staged in shared memory.
    // The thread's first wall / goal / rock quads are requested before the shared-memory
    // set-up and its barrier (one DRAM round trip less in the CTA's dependent chain).
    const In* wsrc = walls + (size_t)e * H * W;
    const In* gsrc = goals + (size_t)e * H * W;
    const In* rsrc = rocks + (size_t)e * R * h * h;
    // (fixed geometries only: in the generic instantiation the twelve extra live registers
    // spill -- the heat-map sweep's launch got 26 % slower with it)
    In a0[4], b0[4], r0[4];
    if constexpr (kFixed) {
      if ((uint32_t)tid < (uint32_t)(H * W) / 4) {
        load4(wsrc + 4 * tid, a0);
        load4(gsrc + 4 * tid, b0);
      }
      if ((uint32_t)tid < (uint32_t)(R * h * h) / 4) load4(rsrc + 4 * tid, r0);
    }
    for (int k = tid; k < H * g_nW + R * g_ng; k += kSelThreads) below[k] = 0u;
    // (asynchronously: the score maps are first read after the overlap counts)
    if (q.stage_values)
      vals = stage_phased<V, true>(vals, values + (size_t)e * R * P, (size_t)R * P, tid,
                                   kSelThreads);
    if (tid < 32) s_cmax[tid] = 0;
    __syncthreads();
    for (uint32_t k = tid; k < (uint32_t)(H * W) / 4; k += kSelThreads) {
      In a[4], b[4];
      if (kFixed && k == (uint32_t)tid) {
#pragma unroll
        for (int t = 0; t < 4; ++t) { a[t] = a0[t]; b[t] = b0[t]; }
      } else {
        load4(wsrc + 4 * k, a);
        load4(gsrc + 4 * k, b);
      }
      const uint32_t bits = (a[0] < b[0] ? 1u : 0u) | (a[1] < b[1] ? 2u : 0u) |
                            (a[2] < b[2] ? 4u : 0u) | (a[3] < b[3] ? 8u : 0u);
      const uint32_t row = kFixed ? k / (uint32_t)(W / 4) : udiv_mul(k, q.mulW4), col = 4 * (k - row * (W / 4));
      if (bits) atomicOr(below + row * g_nW + (col >> 5), bits << (col & 31));
    }
    for (uint32_t k = tid; k < (uint32_t)(R * h * h) / 4; k += kSelThreads) {
      In a[4];
      if (kFixed && k == (uint32_t)tid) {
#pragma unroll
        for (int t = 0; t < 4; ++t) a[t] = r0[t];
      } else {
        load4(rsrc + 4 * k, a);
      }
      const uint32_t bits = (a[0] > In(0) ? 1u : 0u) | (a[1] > In(0) ? 2u : 0u) |
                            (a[2] > In(0) ? 4u : 0u) | (a[3] > In(0) ? 8u : 0u);
      const uint32_t rr = kFixed ? k / (uint32_t)(h * h / 4) : udiv_mul(k, q.mulhh4), rem = k - rr * (h * h / 4);
      const uint32_t u = kFixed ? rem / (uint32_t)(h / 4) : udiv_mul(rem, q.mulh4), col = 4 * (rem - u * (h / 4));
      const uint32_t sub = u & (g_pf - 1);
      if (bits) atomicOr(foot + rr * g_ng + (u >> lp), bits << (sub * g_hb + col));
    }
    __syncthreads();
  } else {
  stage_bytes(vals, values + (size_t)e * R * P, (size_t)R * P * sizeof(V), tid, kSelThreads);
  stage_bytes(wall, walls + (size_t)e * H * W, (size_t)H * W * sizeof(In), tid, kSelThreads);
  stage_bytes(goal, goals + (size_t)e * H * W, (size_t)H * W * sizeof(In), tid, kSelThreads);
  stage_bytes(rock, rocks + (size_t)e * R * h * h, (size_t)R * h * h * sizeof(In), tid,
              kSelThreads);
  if (tid < 32) s_cmax[tid] = 0;
  __syncthreads();

  // ---- bit images of the raw observation (baselines.py:153-154) ---------------- //
  for (int row = warp; row < H; row += NW) {
    for (int word = 0; word < g_nW; ++word) {
      const int col = word * 32 + lane;
      bool b = false;
      if (col < W) b = wall[row * W + col] < goal[row * W + col];
      const uint32_t bits = __ballot_sync(0xffffffffu, b);
      if (lane == 0) below[row * g_nW + word] = bits;
    }
  }
  for (int r = warp; r < R; r += NW) {
    const In* rk = rock + (size_t)r * h * h;
    uint32_t packed = 0;
    for (int u = 0; u < h; ++u) {
      const uint32_t bits = __ballot_sync(0xffffffffu, lane < h && rk[u * h + lane] > In(0));
      const int sub = u & (g_pf - 1);
      packed |= (bits & hmask) << (sub * g_hb);
      if (sub == g_pf - 1 || u == h - 1) {
        if (lane == 0) foot[r * g_ng + (u >> lp)] = packed;
        packed = 0;
      }
    }
  }
  __syncthreads();
  }
  // Row-packed windows win[row][j] = x(row) | x(row + 1) << hb | ... (pf rows per word), where
  // x(row) is the h-bit window of wall row `row` at column j.  A thread walks down one
  // column segment and slides the packed word (x(row) is extracted once, not once per packed
  // word it is part of); the threads of a warp are consecutive columns, so the two words of a
  // window are the same (or neighbouring) words for the whole warp and the stores are dense.
  const int nseg = kSelThreads / Pw;
  if (nseg > 0) {
    const int seglen = (H + nseg - 1) / nseg;
    const int seg = tid / Pw, j = tid - seg * Pw;
    const int r_lo = seg * seglen, r_hi = min(H, r_lo + seglen);
    if (seg < nseg && r_lo < r_hi) {
      const uint32_t* b = below + (j >> 5);
      const int sh = j & 31, top = (g_pf - 1) * g_hb;
      auto x = [&](int row) -> uint32_t {
        return row < H ? __funnelshift_r(b[row * g_nW], b[row * g_nW + 1], sh) & hmask : 0u;
      };
      uint32_t acc = 0;
      for (int s2 = 0; s2 + 1 < g_pf; ++s2) acc |= x(r_lo + s2) << ((s2 + 1) * g_hb);
      for (int row = r_lo; row < r_hi; ++row) {
        const uint32_t in = x(row + g_pf - 1);
        acc = g_pf == 1 ? in : ((acc >> g_hb) | (in << top));
        win[row * Pw + j] = acc;
      }
    }
  } else {
    // (more output columns than threads: one warp per wall row, lanes along the columns)
    for (int row = warp; row < H; row += NW) {
      for (int j = lane; j < Pw; j += 32) {
        const uint32_t* b = below + row * g_nW + (j >> 5);
        uint32_t packed = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s < g_pf && row + s < H)
            packed |= (__funnelshift_r(b[s * g_nW], b[s * g_nW + 1], j & 31) & hmask) << (s * g_hb);
        }
        win[row * Pw + j] = packed;
      }
    }
  }
  __syncthreads();

  // ---- lanes = (view, position slot), RCH views at a time -------------------------- //
  // RCH is the largest power of two (<= 32) that divides R: 8 rotations go in one
  // chunk, 36 in nine chunks of 4, an odd count one view at a time.
  const int RCH = kFixed ? R : q.rch;
  const int rl = lane & (RCH - 1);
  const int ppw = 32 / RCH;                     // positions per warp iteration
  const int qpos = lane / RCH;
  const int step = NW * ppw;
  const int gstride = g_pf * Pw;
  uint16_t* mine = cnt + (size_t)rl * P;
  for (int rc0 = 0; rc0 < R; rc0 += RCH) {
    const int r = rc0 + rl;
    uint32_t f[NG > 0 ? NG : 1];
    if (NG > 0) {
#pragma unroll
      for (int g = 0; g < NG; ++g) f[g] = foot[r * NG + g];
    }
    // Big maps are not staged (they would leave one CTA per SM): the score map of
    // the lane's view is read from global memory / L1 instead.
    const V* v = (kFixed || q.vec4) && !q.stage_values ? values + ((size_t)e * R + r) * P
                                                       : vals + (size_t)r * P;
    int cm = 0;
    for (int pos = warp * ppw + qpos; pos < P; pos += step) {
      const uint32_t* wp = win + pos;
      int c = 0;
      if (NG > 0) {
#pragma unroll
        for (int g = 0; g < NG; ++g) c += __popc(wp[g * gstride] & f[g]);
      } else {
        for (int g = 0; g < g_ng; ++g) c += __popc(wp[g * gstride] & foot[r * g_ng + g]);
      }
      mine[pos] = (uint16_t)c;
      cm = max(cm, c);
    }
    for (int o = RCH; o < 32; o <<= 1) cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    if (qpos == 0) atomicMax(s_cmax + rl, cm);
    stage_wait();
    __syncthreads();

    // ---- mask cut, masked maximum, arg-min candidates ------------------------------- //
    const int cmin = (int)ceil(q.overlap_threshold * (double)s_cmax[rl]);
    Cand<V> bmin = {V(0), -1}, bmask = {V(0), -1};
    V vm = V(0);
    bool any = false;
    for (int pos = warp * ppw + qpos; pos < P; pos += step) {
      if ((int)mine[pos] < cmin) continue;
      const V x = v[pos];
      vm = (!any || x > vm) ? x : vm;
      any = true;
      // Positions come in increasing order per lane, so only a strictly smaller
      // value can replace a candidate: test that before the neighbourhood.
      if (bmask.idx < 0 || x < bmask.v) {
        bmask.v = x;
        bmask.idx = pos;
      }
      if (M != 0 && (bmin.idx < 0 || x < bmin.v)) {
        const int i = kFixed ? pos / Pw : (int)__umulhi((uint32_t)pos, q.mulPw),
                  j = pos - i * Pw;
        if (local_min<V, M>(v, x, i, j, Ph, Pw, q.minorder)) {
          bmin.v = x;
          bmin.idx = pos;
        }
      }
    }
    if constexpr (sizeof(V) == 4) {
      // one view per warp pass (every single-view geometry): the three reductions over the
      // lanes are REDUX on order-preserving integer keys instead of five shuffle rounds each
      if (RCH == 1) {
        warp_cand_redux(bmin);
        warp_cand_redux(bmask);
        const uint32_t kmax = __reduce_max_sync(0xffffffffu, any ? ord_key(vm) : 0u);
        any = kmax != 0u;
        vm = any ? ord_val(kmax) : V(0);
      }
    }
    for (int o = (sizeof(V) == 4 && RCH == 1) ? 32 : RCH; o < 32; o <<= 1) {
      const V xv = __shfl_xor_sync(0xffffffffu, bmin.v, o);
      const int xi = __shfl_xor_sync(0xffffffffu, bmin.idx, o);
      if (xi >= 0) take(bmin, xv, xi);
      const V yv = __shfl_xor_sync(0xffffffffu, bmask.v, o);
      const int yi = __shfl_xor_sync(0xffffffffu, bmask.idx, o);
      if (yi >= 0) take(bmask, yv, yi);
      const V zv = __shfl_xor_sync(0xffffffffu, vm, o);
      const bool za = __shfl_xor_sync(0xffffffffu, (int)any, o) != 0;
      if (za) {
        vm = (!any || zv > vm) ? zv : vm;
        any = true;
      }
    }
    if (qpos == 0) {
      s_bmin[warp][rl] = bmin;
      s_bmask[warp][rl] = bmask;
      s_vm[warp][rl] = vm;
      s_any[warp][rl] = any;
    }
    __syncthreads();
    if (tid < RCH) {
      Cand<V> a = {V(0), -1}, b = {V(0), -1};
      V m2 = V(0);
      bool have = false;
      for (int w = 0; w < NW; ++w) {
        if (s_bmin[w][tid].idx >= 0) take(a, s_bmin[w][tid].v, s_bmin[w][tid].idx);
        if (s_bmask[w][tid].idx >= 0) take(b, s_bmask[w][tid].v, s_bmask[w][tid].idx);
        if (s_any[w][tid]) {
          m2 = (!have || s_vm[w][tid] > m2) ? s_vm[w][tid] : m2;
          have = true;
        }
      }
      const Cand<V> pick = a.idx >= 0 ? a : b;
      actions[(size_t)e * R + rc0 + tid] = pick.idx;
      s_pick[rc0 + tid] = pick.v;
      s_pick_i[rc0 + tid] = pick.idx;
      s_fill[tid] = (double)m2 + 0.001;
      s_cmax[tid] = 0;                            // for the next chunk of views
    }
    __syncthreads();
    if (shown) {
      const double fill = s_fill[rl];
      double* sh = shown + ((size_t)e * R + r) * P;
      for (int pos = warp * ppw + qpos; pos < P; pos += step)
        sh[pos] = -((int)mine[pos] >= cmin ? (double)v[pos] : fill);
    }
    // the next chunk overwrites the counts and the per-warp partials
    if (rc0 + RCH < R) __syncthreads();
  }
  if (best && tid == 0) {
    // PyGreedy batchwise: first argmax over views of -value (policies.py:78-80)
    int bv = 0;
    for (int k = 1; k < R; ++k)
      if (s_pick[k] < s_pick[bv]) bv = k;
    best[2 * (size_t)e] = bv;
    best[2 * (size_t)e + 1] = s_pick_i[bv];
  }
}

template <typename V, typename In>
int launch_mask_select(const V* values, const In* walls, const In* goals, const In* rocks,
                       int64_t* actions, double* shown, int64_t* best, int E, int R, int H,
                       int W, int h, int minorder, double overlap_threshold,
                       cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h && minorder >= 0, SRL_E_INVALID,
              "mask_select: bad shape E=%d R=%d H=%d W=%d h=%d minorder=%d", E, R, H, W, h,
              minorder);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(values && walls && goals && rocks && actions, SRL_E_INVALID,
              "mask_select: null pointer");
  SRL_REQUIRE(h <= 32, SRL_E_UNSUPPORTED, "mask_select: rock side %d > 32", h);
  MaskSelectParams q;
  q.R = R; q.H = H; q.W = W; q.h = h; q.Ph = H - h + 1; q.Pw = W - h + 1;
  q.minorder = minorder; q.overlap_threshold = overlap_threshold;
  q.nW = (W + 31) / 32 + 1;
  q.pf = h > 16 ? 1 : (h > 8 ? 2 : 4);
  q.hb = 32 / q.pf;
  q.ng = (h + q.pf - 1) / q.pf;
  q.mulPw = q.Pw == 1 ? 0u : (uint32_t)(((1ull << 32) + q.Pw - 1) / q.Pw);
  SRL_REQUIRE(q.Pw > 1, SRL_E_UNSUPPORTED, "mask_select: single-column maps");
  const size_t P = (size_t)q.Ph * q.Pw;
  SRL_REQUIRE((size_t)H * q.Pw * q.Pw < (1ull << 32) && P * q.Pw < (1ull << 32),
              SRL_E_UNSUPPORTED, "mask_select: map too large");
  size_t smem = 4 * ((size_t)H * q.nW + (size_t)R * q.ng + (size_t)H * q.Pw) +
                2 * (size_t)(kSelThreads / 32) * P;
  SRL_REQUIRE(smem <= 220 * 1024, SRL_E_UNSUPPORTED,
              "mask_select: %dx%d wall exceeds shared memory", H, W);
  auto pad16 = [](size_t n) { return (n + 15) & ~(size_t)15; };
  const size_t staged = 16 + pad16((size_t)R * P * sizeof(V)) +
                        2 * pad16((size_t)H * W * sizeof(In)) +
                        pad16((size_t)R * h * h * sizeof(In));
  q.staged = smem + staged <= 56 * 1024;        // keep >= 4 CTAs per SM
  if (q.staged) smem += staged;
#define SRL_MS_LAUNCH(ST, MM)                                                              \
  do {                                                                                     \
    auto k = mask_select_kernel<V, In, ST, MM>;                                            \
    SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                  (int)smem));                                             \
    k<<<E, kSelThreads, smem, stream>>>(values, walls, goals, rocks, actions, shown, best, \
                                        q);                                                \
  } while (0)
  int rch = 1;
  while (rch < 32 && R % (rch * 2) == 0) rch *= 2;
  q.rch = rch;
  const bool pow2 = R <= kMaxViews;      // (any view count: rch views at a time)
  auto mulc = [](uint32_t d) { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d); };
  const size_t al = sizeof(In) * 4;
  q.vec4 = W % 4 == 0 && h % 4 == 0 && ((uintptr_t)walls % al) == 0 &&
           ((uintptr_t)goals % al) == 0 && ((uintptr_t)rocks % al) == 0 &&
           (size_t)R * h * h * (h * h / 4) < (1ull << 32);
  q.mulW4 = mulc(W / 4); q.mulh4 = mulc(h / 4); q.mulhh4 = mulc(h * h / 4);
  const size_t packed_base =
      4 * ((size_t)H * q.nW + (size_t)R * q.ng + (size_t)H * q.Pw) + 2 * (size_t)rch * P;
  // stage the score maps while that keeps >= 4 CTAs per SM
  q.stage_values = !q.vec4 || packed_base + 32 + pad16((size_t)R * P * sizeof(V)) <= 52 * 1024;
  const size_t packed_smem =
      packed_base + (!q.vec4 ? staged
                             : (q.stage_values ? 32 + pad16((size_t)R * P * sizeof(V)) : 32));
  if (pow2 && packed_smem <= 200 * 1024 && minorder <= 1) {     // one CTA per SM at worst
#define SRL_MSP_LAUNCH(MM, NGG)                                                            \
  do {                                                                                     \
    auto k = mask_select_packed_kernel<V, In, MM, NGG>;                                    \
    SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                  (int)packed_smem));                                      \
    k<<<E, kSelThreads, packed_smem, stream>>>(values, walls, goals, rocks, actions,       \
                                               shown, best, q);                            \
  } while (0)
    // Reference geometries with every index folded at compile time.
#define SRL_GEO(HH, WW, hh, RR) (((((long long)(HH) << 20) | ((WW) << 10) | (hh)) << 6) | (RR))
#define SRL_MSP_FIXED(HH, WW, hh, RR, NGG)                                                 \
  if (q.vec4 && minorder == 1 && H == HH && W == WW && h == hh && R == RR) {               \
    auto k = mask_select_packed_kernel<V, In, 1, NGG, SRL_GEO(HH, WW, hh, RR)>;            \
    SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                  (int)packed_smem));                                      \
    k<<<E, kSelThreads, packed_smem, stream>>>(values, walls, goals, rocks, actions,       \
                                               shown, best, q);                            \
    return check_launch("mask_select_packed_kernel");                                      \
  }
    SRL_MSP_FIXED(32, 32, 16, 8, 8)
    SRL_MSP_FIXED(64, 64, 16, 8, 8)
    SRL_MSP_FIXED(64, 64, 16, 1, 8)
#undef SRL_MSP_FIXED
#undef SRL_GEO
    if (minorder == 0) {
      if (q.ng == 8) SRL_MSP_LAUNCH(0, 8);
      else if (q.ng == 32) SRL_MSP_LAUNCH(0, 32);
      else SRL_MSP_LAUNCH(0, 0);
    } else {
      if (q.ng == 8) SRL_MSP_LAUNCH(1, 8);
      else if (q.ng == 32) SRL_MSP_LAUNCH(1, 32);
      else SRL_MSP_LAUNCH(1, 0);
    }
#undef SRL_MSP_LAUNCH
    return check_launch("mask_select_packed_kernel");
  }
  if (q.staged) {
    if (minorder == 0) SRL_MS_LAUNCH(true, 0);
    else if (minorder == 1) SRL_MS_LAUNCH(true, 1);
    else SRL_MS_LAUNCH(true, -1);
  } else {
    if (minorder == 0) SRL_MS_LAUNCH(false, 0);
    else if (minorder == 1) SRL_MS_LAUNCH(false, 1);
    else SRL_MS_LAUNCH(false, -1);
  }
#undef SRL_MS_LAUNCH
  return check_launch("mask_select_kernel");
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

int mask_select_f32(const float* values, const float* walls, const float* goals,
                    const float* rocks, int64_t* actions, double* shown, int64_t* best,
                    int E, int R, int H, int W, int h, int minorder, double overlap_threshold,
                    cudaStream_t stream) {
  return launch_mask_select<float, float>(values, walls, goals, rocks, actions, shown, best, E,
                                          R, H, W, h, minorder, overlap_threshold, stream);
}

int mask_select_f64(const double* values, const float* walls, const float* goals,
                    const float* rocks, int64_t* actions, double* shown, int64_t* best,
                    int E, int R, int H, int W, int h, int minorder, double overlap_threshold,
                    cudaStream_t stream) {
  return launch_mask_select<double, float>(values, walls, goals, rocks, actions, shown, best,
                                           E, R, H, W, h, minorder, overlap_threshold, stream);
}

int mask_select_f64_u8(const double* values, const uint8_t* walls, const uint8_t* goals,
                       const uint8_t* rocks, int64_t* actions, double* shown, int64_t* best,
                       int E, int R, int H, int W, int h, int minorder,
                       double overlap_threshold, cudaStream_t stream) {
  return launch_mask_select<double, uint8_t>(values, walls, goals, rocks, actions, shown, best,
                                             E, R, H, W, h, minorder, overlap_threshold,
                                             stream);
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
