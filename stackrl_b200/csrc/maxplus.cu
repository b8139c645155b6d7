// Batched max-plus "drop" search (reference: stackrl/baselines.py:21-43,
// get_inputs + height; the Python double loop over positions at :34-41).
//
//   out[e,r,i,j] = max_{u,v}( n[e,r,u,v] > thr ? o[e,i+u,j+v] + n[e,r,u,v] : 0 )
//   o = wall/level, n = rock/level  (IEEE float32 division, then float32 add)
//
// Design (see DESIGN.md section "maxplus_f32"):
//   * One CTA owns G whole environments (wall + RC rock rotations each), staged
//     into shared memory with 1-D bulk TMA copies (cp.async.bulk, one mbarrier),
//     one copy per wall row so rows land on a padded stride (Ws/4 odd => the
//     per-row LDS.128 of 8 consecutive lanes hit 8 distinct 16-B bank groups).
//   * A normalise pass divides by the goal level in place and folds the
//     `n > thr` mask into the rock as -inf, so the inner loop has no select:
//     o + (-inf) = -inf never wins the max.  The reference's "masked cell
//     contributes 0" (quirk Q2) becomes one max(acc, 0) at the end, applied
//     only when the rock really has a masked cell.
//   * Each thread owns T consecutive outputs of one output row.  Per rock row
//     it loads T+VC-1 wall values and VC rock values with LDS.128 and runs the
//     fully unrolled T x VC block of (add, max) cells out of registers.
//   * sm_100 instruction mix: cells are paired so that two adds issue as one
//     FADD2 (add.rn.f32x2, FMA pipe) and fold into the accumulator with one
//     3-input FMNMX3 (ALU pipe): ~1.03 issue slots per cell instead of 2.
//     Max is exact and order independent, each add is a single IEEE rn add, so
//     the result is bit-identical to numpy's.
//   * No tensor cores: max-plus is not a dense contraction.
#include <stdlib.h>

#include "common.cuh"

namespace srl {

struct MaxPlusParams {
  const float* walls;
  const float* rocks;
  const float* level;
  float* out;
  int E, R, H, W, h;
  int Ph, Pw;
  int hp;           // rock columns padded to a multiple of VC
  int Ws;           // smem wall row stride (floats), Ws % 4 == 0, (Ws/4) odd
  int wall_stride;  // smem floats per wall  (H * Ws)
  int rock_stride;  // smem floats per rock copy (h * hp + 4)
  int G;            // environments per CTA
  int RC;           // rotations per CTA
  int rchunks;      // ceil(R / RC)
  int strips;       // strips per output row; strip k starts at column k*(T-1)
  float threshold;
  int tma_wall;     // wall rows can be bulk-copied (W % 4 == 0, 16-B aligned base)
  int tma_rock;     // rock rows can be bulk-copied (h % 4 == 0, 16-B aligned base)
};

// Block of T x VC (add, max) cells: acc[t] = max(acc[t], row[t+v] + nv[v]).
// PAIRED: nvs[v] = nv[v+1] is the one-column-shifted rock row, loaded from its
// own smem copy so that (nvs[v], nvs[v+1]) for even v is an aligned register
// pair holding (nv[v+1], nv[v+2]).
template <int T, int VC, bool PAIRED>
__device__ __forceinline__ void cell_block(float (&acc)[T],
                                           const float (&row)[4 * ((T + VC + 2) / 4)],
                                           const float (&nv)[VC],
                                           const float (&nvs)[VC]) {
  if constexpr (!PAIRED) {
#pragma unroll
    for (int v = 0; v < VC; ++v) {
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = fmaxf(acc[t], row[t + v] + nv[v]);
    }
  } else {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if ((t & 1) == 0) {
        // even output column: wall index k = t+v is even for even v.
#pragma unroll
        for (int v = 0; v < VC; v += 2) {
          float s0, s1;
          fadd2(s0, s1, row[t + v], row[t + v + 1], nv[v], nv[v + 1]);
          acc[t] = fmax3(acc[t], s0, s1);
        }
      } else {
        // odd output column: pair odd v with v+1 (k = t+v even), using the
        // shifted rock row; v = 0 and v = VC-1 stay single.
        float e0 = row[t] + nv[0];
        float e1 = row[t + VC - 1] + nv[VC - 1];
        acc[t] = fmax3(acc[t], e0, e1);
#pragma unroll
        for (int v = 1; v + 1 < VC; v += 2) {
          float s0, s1;
          fadd2(s0, s1, row[t + v], row[t + v + 1], nvs[v - 1], nvs[v]);
          acc[t] = fmax3(acc[t], s0, s1);
        }
      }
    }
  }
}

template <int T, int VC, bool PAIRED>
__global__ void __launch_bounds__(288, 2)
maxplus_f32_kernel(const MaxPlusParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  int* masked = reinterpret_cast<int*>(smem_raw + 16);
  const int flag_bytes = round_up(p.G * p.RC * 4, 16);
  float* wall_s = reinterpret_cast<float*>(smem_raw + 16 + flag_bytes);
  float* rock_s = wall_s + p.G * p.wall_stride;
  // PAIRED keeps a second, one-column-shifted copy of every rock behind the first.
  float* rock_sh = rock_s + p.G * p.RC * p.rock_stride;

  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const int group = blockIdx.x / p.rchunks;
  const int rchunk = blockIdx.x % p.rchunks;
  const int e0 = group * p.G;
  const int r0 = rchunk * p.RC;
  const int Gv = min(p.G, p.E - e0);
  const int RCv = min(p.RC, p.R - r0);
  const int H = p.H, W = p.W, h = p.h, hp = p.hp, Ws = p.Ws;

  // ---- stage 0: barrier init ------------------------------------------------ //
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  for (int k = tid; k < p.G * p.RC; k += nthreads) masked[k] = 0;
  __syncthreads();

  // ---- stage 1: bulk TMA loads (warp 0) + padding fills (everyone) ---------- //
  const bool any_tma = p.tma_wall || p.tma_rock;
  if (tid < 32 && any_tma) {
    if (tid == 0) {
      uint32_t bytes = 0;
      if (p.tma_wall) bytes += (uint32_t)Gv * H * W * 4;
      if (p.tma_rock) bytes += (uint32_t)Gv * RCv * h * h * 4;
      mbar_arrive_expect_tx(bar, bytes);
    }
    __syncwarp();
    if (p.tma_wall) {
      for (int k = tid; k < Gv * H; k += 32) {
        const int el = k / H, row = k % H;
        tma_load_1d(wall_s + el * p.wall_stride + row * Ws,
                    p.walls + ((size_t)(e0 + el) * H + row) * W, W * 4, bar);
      }
    }
    if (p.tma_rock) {
      if (hp == h) {
        for (int k = tid; k < Gv * RCv; k += 32) {
          const int el = k / RCv, r = k % RCv;
          tma_load_1d(rock_s + (el * p.RC + r) * p.rock_stride,
                      p.rocks + ((size_t)(e0 + el) * p.R + r0 + r) * h * h,
                      h * h * 4, bar);
        }
      } else {
        for (int k = tid; k < Gv * RCv * h; k += 32) {
          const int u = k % h, er = k / h;
          const int el = er / RCv, r = er % RCv;
          tma_load_1d(rock_s + (el * p.RC + r) * p.rock_stride + u * hp,
                      p.rocks + (((size_t)(e0 + el) * p.R + r0 + r) * h + u) * h,
                      h * 4, bar);
        }
      }
    }
  }
  if (!p.tma_wall) {
    for (int k = tid; k < Gv * H * W; k += nthreads) {
      const int el = k / (H * W), rem = k % (H * W);
      wall_s[el * p.wall_stride + (rem / W) * Ws + rem % W] =
          __ldg(p.walls + (size_t)(e0 + el) * H * W + rem);
    }
  }
  if (!p.tma_rock) {
    for (int k = tid; k < Gv * RCv * h * h; k += nthreads) {
      const int er = k / (h * h), rem = k % (h * h);
      const int el = er / RCv, r = er % RCv;
      rock_s[(el * p.RC + r) * p.rock_stride + (rem / h) * hp + rem % h] =
          __ldg(p.rocks + ((size_t)(e0 + el) * p.R + r0 + r) * h * h + rem);
    }
  }
  // Wall pad columns [W, Ws) must be finite; rock pad columns [h, hp) are -inf
  // (never win, and do not count as "masked" cells).
  {
    const int padw = Ws - W;
    for (int k = tid; k < Gv * H * padw; k += nthreads) {
      const int el = k / (H * padw), rem = k % (H * padw);
      wall_s[el * p.wall_stride + (rem / padw) * Ws + W + rem % padw] = 0.f;
    }
    const int padr = hp - h;
    for (int k = tid; k < Gv * RCv * h * padr; k += nthreads) {
      const int er = k / (h * padr), rem = k % (h * padr);
      const int el = er / RCv, r = er % RCv;
      rock_s[(el * p.RC + r) * p.rock_stride + (rem / padr) * hp + h + rem % padr] =
          kNegInf;
    }
  }
  if (any_tma) mbar_wait(bar, 0);
  __syncthreads();

  // ---- stage 2: normalise in place, fold the mask into the rock ------------- //
  if (p.level != nullptr) {
    for (int k = tid; k < Gv * H * W; k += nthreads) {
      const int el = k / (H * W), rem = k % (H * W);
      float* q = wall_s + el * p.wall_stride + (rem / W) * Ws + rem % W;
      *q = __fdiv_rn(*q, __ldg(p.level + e0 + el));
    }
  }
  for (int k = tid; k < Gv * RCv * h * h; k += nthreads) {
    const int er = k / (h * h), rem = k % (h * h);
    const int el = er / RCv, r = er % RCv;
    const int slot = el * p.RC + r;
    float* q = rock_s + slot * p.rock_stride + (rem / h) * hp + rem % h;
    float n = *q;
    if (p.level != nullptr) n = __fdiv_rn(n, __ldg(p.level + e0 + el));
    const bool live = n > p.threshold;
    if (!live) masked[slot] = 1;
    *q = live ? n : kNegInf;
  }
  __syncthreads();
  if constexpr (PAIRED) {
    // shifted copy: rock_sh[u][v] = rock_s[u][v+1] (last column -inf)
    for (int k = tid; k < Gv * RCv * h * hp; k += nthreads) {
      const int er = k / (h * hp), rem = k % (h * hp);
      const int el = er / RCv, r = er % RCv;
      const int slot = el * p.RC + r;
      const int v = rem % hp;
      rock_sh[slot * p.rock_stride + rem] =
          (v + 1 < hp) ? rock_s[slot * p.rock_stride + rem + 1] : kNegInf;
    }
    __syncthreads();
  }

  // ---- stage 3: register-tiled (add, max) sweep ----------------------------- //
  const int Ph = p.Ph, Pw = p.Pw;
  const int items = Gv * RCv * p.strips * Ph;
  constexpr int NR4 = (T + VC + 2) / 4;   // float4 loads per wall row chunk
  constexpr int S = T - 1;                // strip pitch (multiple of 4)
  static_assert(S % 4 == 0 && 4 * NR4 >= T + VC - 1, "tile shape");
  for (int item = tid; item < items; item += nthreads) {
    const int i = item % Ph;
    int rest = item / Ph;
    const int strip = rest % p.strips;
    rest /= p.strips;
    const int r = rest % RCv;
    const int el = rest / RCv;
    const int slot = el * p.RC + r;
    const float* wbase = wall_s + el * p.wall_stride + i * Ws + strip * S;
    const float* rbase = rock_s + slot * p.rock_stride;
    const float* sbase = rock_sh + slot * p.rock_stride;

    float acc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = kNegInf;

    for (int u = 0; u < h; ++u) {
      for (int vc = 0; vc < hp; vc += VC) {
        float row[4 * NR4];
        float nv[VC];
        float nvs[VC];
#pragma unroll
        for (int k = 0; k < NR4; ++k) {
          const float4 x = lds128(wbase + u * Ws + vc + 4 * k);
          row[4 * k + 0] = x.x;
          row[4 * k + 1] = x.y;
          row[4 * k + 2] = x.z;
          row[4 * k + 3] = x.w;
        }
#pragma unroll
        for (int k = 0; k < VC / 4; ++k) {
          const float4 x = lds128(rbase + u * hp + vc + 4 * k);
          nv[4 * k + 0] = x.x;
          nv[4 * k + 1] = x.y;
          nv[4 * k + 2] = x.z;
          nv[4 * k + 3] = x.w;
          if constexpr (PAIRED) {
            const float4 y = lds128(sbase + u * hp + vc + 4 * k);
            nvs[4 * k + 0] = y.x;
            nvs[4 * k + 1] = y.y;
            nvs[4 * k + 2] = y.z;
            nvs[4 * k + 3] = y.w;
          }
        }
        cell_block<T, VC, PAIRED>(acc, row, nv, nvs);
      }
    }

    // ---- stage 4: the reference's zero floor, then store --------------------- //
    const bool floor0 = masked[slot] != 0;
    float* orow = p.out +
                  (((size_t)(e0 + el) * p.R + r0 + r) * Ph + i) * (size_t)Pw +
                  strip * S;
    // The last column of a strip is the first of the next one; only the last
    // strip stores it.
    const int ncols = (strip == p.strips - 1) ? min(T, Pw - strip * S) : S;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float v = acc[t];
      if (floor0) v = fmaxf(v, 0.f);
      if (t < ncols) __stcs(orow + t, v);
    }
  }
}

// --------------------------------------------------------------------------- //
// Host-side dispatch.
// --------------------------------------------------------------------------- //
namespace {

struct Choice {
  int T, VC;
};

// Per-thread tile widths with an instantiation: T = 4m+1 outputs, strips pitched
// S = 4m apart (16-B aligned starts).  The reference geometries have
// Pw = 2^a - 2^b + 1 == 1 (mod 4), which these cover with no wasted column.
const int kTs[] = {5, 9, 13, 17, 21, 25};

int strips_for(int Pw, int T) {
  return Pw <= T ? 1 : (Pw - T + (T - 1) - 1) / (T - 1) + 1;
}

Choice choose_tile(int Pw, int h) {
  Choice c;
  c.VC = h >= 13 ? 16 : (h >= 5 ? 8 : 4);
  int best = kTs[0];
  double best_cost = 1e30;
  for (int T : kTs) {
    if (c.VC == 16 && T > 21) continue;   // register budget (112/thread)
    const int strips = strips_for(Pw, T);
    const double waste = (double)strips * T / Pw;
    const double loads = ((T + c.VC + 2) / 4 + c.VC / 2) / (double)(T * c.VC);
    const double cost = waste * (1.03 + loads);
    if (cost < best_cost - 1e-12) {
      best_cost = cost;
      best = T;
    }
  }
  c.T = best;
  return c;
}

template <int T, int VC>
int launch(const MaxPlusParams& p, int blocks, int threads, size_t smem,
           bool paired, cudaStream_t stream) {
  if (paired) {
    auto k = maxplus_f32_kernel<T, VC, true>;
    SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    k<<<blocks, threads, smem, stream>>>(p);
  } else {
    auto k = maxplus_f32_kernel<T, VC, false>;
    SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    k<<<blocks, threads, smem, stream>>>(p);
  }
  return check_launch("maxplus_f32_kernel");
}

}  // namespace

int maxplus_f32(const float* walls, const float* rocks, const float* level,
                float* out, int E, int R, int H, int W, int h, float threshold,
                int variant, cudaStream_t stream) {
  SRL_REQUIRE(walls && rocks && out, SRL_E_INVALID, "maxplus_f32: null pointer");
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h, SRL_E_INVALID,
              "maxplus_f32: bad shape E=%d R=%d H=%d W=%d h=%d", E, R, H, W, h);
  if (E == 0) return SRL_OK;

  MaxPlusParams p;
  p.walls = walls; p.rocks = rocks; p.level = level; p.out = out;
  p.E = E; p.R = R; p.H = H; p.W = W; p.h = h;
  p.Ph = H - h + 1; p.Pw = W - h + 1;
  p.threshold = threshold;

  const Choice c = choose_tile(p.Pw, h);
  const int T = c.T, VC = c.VC;
  const bool paired = (variant != 0) && VC >= 4;
  p.hp = round_up(h, VC);
  p.strips = strips_for(p.Pw, T);
  // Columns a thread may touch: strip start + (hp - VC) + 4*NR4 floats.
  const int nr4 = (T + VC + 2) / 4;
  int need = (p.strips - 1) * (T - 1) + (p.hp - VC) + 4 * nr4;
  p.Ws = round_up(need > W ? need : W, 4);
  if ((p.Ws / 4) % 2 == 0) p.Ws += 4;
  p.wall_stride = H * p.Ws;
  p.rock_stride = h * p.hp + 4;
  p.tma_wall = (W % 4 == 0) && (((uintptr_t)walls) % 16 == 0);
  p.tma_rock = (h % 4 == 0) && (((uintptr_t)rocks) % 16 == 0);

  // Shared-memory budget: aim for >= 2 CTAs per SM (one CTA's load/normalise
  // phases overlap the other's sweep).
  const size_t kBudget = 100 * 1024;
  const size_t wall_bytes = (size_t)p.wall_stride * 4;
  const size_t rock_bytes = (size_t)p.rock_stride * 4 * (paired ? 2 : 1);
  auto smem_for = [&](int G, int RC) {
    return 16 + (size_t)round_up(G * RC * 4, 16) + G * wall_bytes +
           (size_t)G * RC * rock_bytes;
  };
  int RC = R, G = 1;
  while (RC > 1 && smem_for(1, RC) > kBudget) RC = (RC + 1) / 2;
  SRL_REQUIRE(smem_for(1, RC) <= 220 * 1024, SRL_E_UNSUPPORTED,
              "maxplus_f32: one wall (%dx%d) + one rock (%d) exceed shared memory",
              H, W, h);
  // Target <= 288 threads per CTA so two CTAs (<= 112 registers/thread) are
  // resident per SM: one CTA's load/normalise phases hide under the other's
  // sweep.  SRL_MP_G / SRL_MP_THREADS override for experiments.
  const int items_per_env = RC * p.strips * p.Ph;
  const int kThreads = 288;
  if (RC == R) {
    while (G < 32 && G < E && smem_for(G + 1, RC) <= kBudget &&
           (G + 1) * items_per_env <= kThreads)
      ++G;
  }
  if (const char* s = getenv("SRL_MP_G")) {
    const int g = atoi(s);
    if (g >= 1 && RC == R && smem_for(g, RC) <= 220 * 1024) G = g;
  }
  p.G = G; p.RC = RC; p.rchunks = (R + RC - 1) / RC;

  const int items = G * items_per_env;
  int threads;
  if (items <= kThreads) {
    threads = round_up(items, 32);
  } else {
    // several passes: pick the warp count that wastes the fewest lanes
    int best_t = 256; double best_w = 1e9;
    for (int t = 192; t <= kThreads; t += 32) {
      const int passes = (items + t - 1) / t;
      const double w = (double)passes * t / items;
      if (w < best_w - 1e-9) { best_w = w; best_t = t; }
    }
    threads = best_t;
  }
  if (const char* s = getenv("SRL_MP_THREADS")) {
    const int t = atoi(s);
    if (t >= 32 && t <= 288 && t % 32 == 0) threads = t;
  }
  const int blocks = ((E + G - 1) / G) * p.rchunks;
  const size_t smem = smem_for(G, RC);

#define SRL_MP_CASE(TT, VV)                                          \
  if (T == TT && VC == VV)                                           \
    return launch<TT, VV>(p, blocks, threads, smem, paired, stream);
#define SRL_MP_ROW(VV)                                                        \
  SRL_MP_CASE(5, VV) SRL_MP_CASE(9, VV) SRL_MP_CASE(13, VV) SRL_MP_CASE(17, VV) \
  SRL_MP_CASE(21, VV) SRL_MP_CASE(25, VV)
  SRL_MP_ROW(4)
  SRL_MP_ROW(8)
  SRL_MP_ROW(16)
#undef SRL_MP_ROW
#undef SRL_MP_CASE
  return fail(SRL_E_UNSUPPORTED, "maxplus_f32: no kernel for T=%d VC=%d", T, VC);
}

}  // namespace srl
