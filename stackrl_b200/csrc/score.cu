// Fused placement scoring: ONE kernel for what the reference's
// Baseline('height', batched, batchwise) does per observation
// (stackrl/baselines.py:201-217 + agents/policies.py:57-91):
//
//   values = height(obs)                      baselines.py:21-43   (max-plus sweep)
//   mask   = goal_overlap(obs, threshold)     baselines.py:152-156 (bit-packed counts)
//   action = arg-min over masked local minima baselines.py:207-215
//   best   = first arg-max over the R views   policies.py:78-80
//
// It is maxplus_staged_kernel (maxplus.cu: persistent CTAs, bulk-TMA prefetch of
// the next group of environments, prep pass, register-tiled FADD2/FMNMX3 sweep,
// score maps staged in shared memory) with an epilogue that never leaves shared
// memory: the goal maps ride along in the same TMA transaction, `wall < goal`
// and `rock > 0` are ballot-packed during the prep pass, and after the sweep
// each thread -- still owning the same (view, output row, strip) item -- counts
// overlaps with AND+POPC on row-packed windows, applies the integer mask cut,
// tests zero-padded local minima on the staged score map and keeps first-index
// arg-min candidates; one warp per view reduces them.  Counts never touch HBM
// and the score maps are written only if the caller asks for them.
//
// Results are identical to the three separate kernels (srl_maxplus_f32,
// srl_goal_overlap_f32, srl_select_f32): same IEEE adds, exact integer counts,
// float compares are exact, ties go to the smaller index.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "maxplus_core.cuh"

namespace srl {

namespace {

struct ScoreParams {
  MaxPlusParams mp;
  const float* goals;       // [E,H,W] or nullptr (goal=False: plain arg-min)
  int64_t* actions;         // [E,R]
  int64_t* best;            // [E,2] or nullptr
  int minorder;
  double overlap_threshold;
  int level_mode;           // 0 none, 1 mp.level[e], 2 max(goals[e]) in-kernel
  int nW;                   // 32-bit words per bit-packed wall row (+1 zero word)
  int pf, hb, ng;           // rows packed per word, bits per packed row, groups
  FastDiv dPw, dNW;
  int region_bytes;         // compute layout, reused by the epilogue (win + counts)
};

__device__ __forceinline__ void keep_min(float& bv, int& bi, float v, int i) {
  if (bi < 0 || v < bv || (v == bv && i < bi)) {
    bv = v;
    bi = i;
  }
}

template <int T, int VC>
__global__ void __launch_bounds__(288, 2)
score_fused_kernel(const ScoreParams sp) {
  const MaxPlusParams& p = sp.mp;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = p.H, W = p.W, h = p.h, hp = p.hp, Ws = p.Ws, R = p.R, G = p.G;
  const int Ph = p.Ph, Pw = p.Pw, P = Ph * Pw;
  constexpr int S = T - 1;
  const int slots = G * R;
  const int nthreads = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const bool with_goal = sp.goals != nullptr;

  // ---- shared-memory carve-up ---------------------------------------------------- //
  unsigned char* at = smem_raw;
  uint64_t* bar = reinterpret_cast<uint64_t*>(at);                at += 16;
  int* masked = reinterpret_cast<int*>(at);                       at += round_up(2 * slots * 4, 16);
  unsigned* level_bits = reinterpret_cast<unsigned*>(at);         at += round_up(G * 4, 16);
  int* cmax = reinterpret_cast<int*>(at);                         at += round_up(slots * 4, 16);
  float* slot_v = reinterpret_cast<float*>(at);                   at += round_up(slots * 4, 16);
  int* slot_i = reinterpret_cast<int*>(at);                       at += round_up(slots * 4, 16);
  float* cand_v = reinterpret_cast<float*>(at);                   at += round_up(2 * nthreads * 4, 16);
  int* cand_i = reinterpret_cast<int*>(at);                       at += round_up(2 * nthreads * 4, 16);
  uint32_t* below = reinterpret_cast<uint32_t*>(at);              at += round_up(G * H * sp.nW * 4, 16);
  uint32_t* foot = reinterpret_cast<uint32_t*>(at);               at += round_up(slots * sp.ng * 4, 16);
  float* raw_goal = reinterpret_cast<float*>(at);                 at += with_goal ? G * H * W * 4 : 0;
  float* raw_wall = reinterpret_cast<float*>(at);                 at += G * H * W * 4;
  float* raw_rock = reinterpret_cast<float*>(at);                 at += slots * h * h * 4;
  float* wall_s = reinterpret_cast<float*>(at);
  float* rock_s = wall_s + G * p.wall_stride;
  float* rock_sh = rock_s + slots * p.rock_stride;
  at += sp.region_bytes;          // >= compute layout and >= the epilogue's win + counts
  float* out_s = reinterpret_cast<float*>(at);
  // The epilogue reuses the compute layout (dead after the sweep) for the
  // row-packed windows and the 16-bit counts.
  uint32_t* win = reinterpret_cast<uint32_t*>(wall_s);            // [G*H*Pw]
  uint16_t* counts = reinterpret_cast<uint16_t*>(win + G * H * Pw);   // [slots*P]

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  for (int k = tid; k < G * p.wall_stride; k += nthreads) wall_s[k] = 0.f;
  for (int k = tid; k < 2 * slots * p.rock_stride; k += nthreads) rock_s[k] = kNegInf;
  for (int k = tid; k < 2 * slots; k += nthreads) masked[k] = 0;
  __syncthreads();

  auto issue_loads = [&](int g) {
    const int e0 = g * G;
    const int Gv = min(G, p.E - e0);
    const uint32_t wb = (uint32_t)Gv * H * W * 4, rb = (uint32_t)Gv * R * h * h * 4;
    mbar_arrive_expect_tx(bar, wb + rb + (with_goal ? wb : 0u));
    tma_load_1d(raw_wall, p.walls + (size_t)e0 * H * W, wb, bar);
    tma_load_1d(raw_rock, p.rocks + (size_t)e0 * R * h * h, rb, bar);
    if (with_goal) tma_load_1d(raw_goal, sp.goals + (size_t)e0 * H * W, wb, bar);
  };
  if (tid == 0 && (int)blockIdx.x < p.ngroups) issue_loads(blockIdx.x);

  const int W4 = W / 4, h4 = h / 4;
  const bool scaled = sp.level_mode != 0;
  const uint32_t hmask = h >= 32 ? 0xffffffffu : ((1u << h) - 1u);
  int it = 0;
  for (int g = blockIdx.x; g < p.ngroups; g += gridDim.x, ++it) {
    const int e0 = g * G;
    const int Gv = min(G, p.E - e0);
    int* flags = masked + (it & 1) * slots;
    int* flags_next = masked + ((it + 1) & 1) * slots;

    mbar_wait(bar, it & 1);

    // ---- goal level + bit images (raw observation, baselines.py:23, :153-154) --- //
    if (tid < G) level_bits[tid] = 0u;
    for (int k = tid; k < slots; k += nthreads) {
      flags_next[k] = 0;
      cmax[k] = 0;
    }
    if (with_goal) {
      __syncthreads();
      for (int k = warp; k < Gv * H * sp.nW; k += nwarps) {
        uint32_t row, word;
        fdivmod((uint32_t)k, sp.dNW, row, word);
        const int col = word * 32 + lane;
        bool b = false;
        float gv = 0.f;
        if (col < W) {
          gv = raw_goal[row * W + col];
          b = raw_wall[row * W + col] < gv;
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, b);
        if (lane == 0) below[k] = bits;
        if (sp.level_mode == 2) {
          // goal heights are >= 0, so their bit patterns order like unsigned ints
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) gv = fmaxf(gv, __shfl_xor_sync(0xffffffffu, gv, o));
          if (lane == 0) atomicMax(level_bits + fdiv(row, p.dH), __float_as_uint(gv));
        }
      }
      for (int k = warp; k < Gv * R * sp.ng; k += nwarps) {
        const int slot = k / sp.ng, grp = k % sp.ng;
        uint32_t packed = 0;
        for (int q = 0; q < sp.pf; ++q) {
          const int u = grp * sp.pf + q;
          const bool b = u < h && lane < h && raw_rock[(slot * h + u) * h + lane] > 0.f;
          packed |= (__ballot_sync(0xffffffffu, b) & hmask) << (q * sp.hb);
        }
        if (lane == 0) foot[k] = packed;
      }
      __syncthreads();
    }

    // ---- prep: raw -> compute layout --------------------------------------------- //
    for (uint32_t q = tid; q < (uint32_t)(Gv * H * W4); q += nthreads) {
      uint32_t row, c4;
      fdivmod(q, p.dW4, row, c4);
      float4 x = lds128(raw_wall + 4 * q);
      if (scaled) {
        const uint32_t el = fdiv(row, p.dH);
        const float lv = sp.level_mode == 1 ? __ldg(p.level + e0 + el)
                                            : __uint_as_float(level_bits[el]);
        const float inv = pow2_inverse(lv);
        x.x = div_level(x.x, lv, inv);
        x.y = div_level(x.y, lv, inv);
        x.z = div_level(x.z, lv, inv);
        x.w = div_level(x.w, lv, inv);
      }
      *reinterpret_cast<float4*>(wall_s + row * Ws + 4 * c4) = x;
    }
    for (uint32_t q = tid; q < (uint32_t)(Gv * R * h * h4); q += nthreads) {
      uint32_t rrow, c4, slot, u;
      fdivmod(q, p.dh4, rrow, c4);
      fdivmod(rrow, p.dh, slot, u);
      float lv = 1.f;
      if (scaled) {
        const uint32_t el = fdiv(slot, p.dRC);
        lv = sp.level_mode == 1 ? __ldg(p.level + e0 + el) : __uint_as_float(level_bits[el]);
      }
      float4 x = lds128(raw_rock + 4 * q);
      bool dead = false;
      const float inv = scaled ? pow2_inverse(lv) : 0.f;
      x.x = prep_rock(x.x, scaled, lv, inv, p.threshold, dead);
      x.y = prep_rock(x.y, scaled, lv, inv, p.threshold, dead);
      x.z = prep_rock(x.z, scaled, lv, inv, p.threshold, dead);
      x.w = prep_rock(x.w, scaled, lv, inv, p.threshold, dead);
      if (dead) flags[slot] = 1;
      *reinterpret_cast<float4*>(rock_s + slot * p.rock_stride + u * hp + 4 * c4) = x;
      float* sh = rock_sh + slot * p.rock_stride + u * hp + 4 * c4;
      if (c4 != 0) sh[-1] = x.x;
      sh[0] = x.y;
      sh[1] = x.z;
      sh[2] = x.w;
      if ((int)c4 == h4 - 1) sh[3] = kNegInf;        // shifted copy: column h-1
    }
    // The epilogue of the previous group reused the compute layout, so the pad
    // cells the sweep may touch are rewritten every time: wall columns [W, Ws)
    // finite, rock columns [h, hp) of both copies -inf.
    {
      const int padw = Ws - W;
      for (int k = tid; k < Gv * H * padw; k += nthreads)
        wall_s[(k / padw) * Ws + W + k % padw] = 0.f;
      const int padr = hp - h;
      for (int k = tid; k < Gv * R * h * padr; k += nthreads) {
        const int rrow = k / padr, c = h + k % padr;
        const int off = (rrow / h) * p.rock_stride + (rrow % h) * hp + c;
        rock_s[off] = kNegInf;
        rock_sh[off] = kNegInf;
      }
    }
    if (tid == 0 && p.stage_out) tma_store_wait_read();
    __syncthreads();

    if (tid == 0 && g + (int)gridDim.x < p.ngroups) {
      fence_proxy_async();
      issue_loads(g + gridDim.x);
    }

    // ---- sweep: this thread's item = (view slot, strip, output row) ---------------- //
    const int items = Gv * R * p.strips * Ph;
    const bool active = tid < items;
    uint32_t i = 0, strip = 0, slot = 0;
    int ncols = 0;
    if (active) {
      uint32_t rest;
      fdivmod((uint32_t)tid, p.dPh, rest, i);
      fdivmod(rest, p.dStrips, slot, strip);
      const uint32_t el = fdiv(slot, p.dRC);
      float acc[T];
      sweep_item<T, VC, 1>(acc, wall_s + el * p.wall_stride + i * Ws + strip * S,
                              rock_s + slot * p.rock_stride, rock_sh + slot * p.rock_stride,
                              h, hp, Ws);
      const bool floor0 = flags[slot] != 0;
      ncols = ((int)strip == p.strips - 1) ? min(T, Pw - (int)strip * S) : S;
      float* o = out_s + (size_t)slot * P + i * Pw + strip * S;
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (t < ncols) o[t] = floor0 ? fmaxf(acc[t], 0.f) : acc[t];
    }
    if (p.stage_out) fence_proxy_async();
    __syncthreads();                                   // score maps complete in out_s
    if (p.out != nullptr) {
      if (p.stage_out) {
        if (tid == 0) {
          tma_store_1d(p.out + (size_t)e0 * R * P, out_s, (uint32_t)Gv * R * P * 4);
          tma_store_commit();
        }
      } else {
        float* o = p.out + (size_t)e0 * R * P;
        for (int k = tid; k < Gv * R * P; k += nthreads) __stcs(o + k, out_s[k]);
      }
    }

    // ---- epilogue 1: row-packed windows of `wall < goal` --------------------------- //
    if (with_goal) {
      for (uint32_t k = tid; k < (uint32_t)(Gv * H * Pw); k += nthreads) {
        uint32_t row, j;
        fdivmod(k, sp.dPw, row, j);
        const uint32_t rin = row - fdiv(row, p.dH) * H;      // row inside its wall
        uint32_t packed = 0;
        for (int q = 0; q < sp.pf; ++q) {
          if ((int)rin + q >= H) break;
          const uint32_t* b = below + (row + q) * sp.nW + (j >> 5);
          packed |= (__funnelshift_r(b[0], b[1], j & 31) & hmask) << (q * sp.hb);
        }
        win[k] = packed;
      }
      __syncthreads();
      // ---- epilogue 2: overlap counts of this item's outputs ----------------------- //
      if (active) {
        const uint32_t el = fdiv(slot, p.dRC);
        const uint32_t* wbase = win + (el * H + i) * Pw + strip * S;
        const uint32_t* fp = foot + slot * sp.ng;
        uint16_t* cp = counts + (size_t)slot * P + i * Pw + strip * S;
        int local = 0;
        for (int t = 0; t < ncols; ++t) {
          int c = 0;
          for (int q = 0; q < sp.ng; ++q) c += __popc(wbase[q * sp.pf * Pw + t] & fp[q]);
          cp[t] = (uint16_t)c;
          local = max(local, c);
        }
        atomicMax(cmax + slot, local);
      }
      __syncthreads();
    }

    // ---- epilogue 3: masked local-minimum / masked arg-min candidates -------------- //
    {
      float bmin_v = 0.f, bmask_v = 0.f;
      int bmin_i = -1, bmask_i = -1;
      if (active) {
        // count >= threshold*max  <=>  count >= ceil(threshold*max) for integers
        const int cmin = with_goal
            ? (int)ceil(sp.overlap_threshold * (double)cmax[slot]) : 0;
        const float* map = out_s + (size_t)slot * P;
        const uint16_t* cp = counts + (size_t)slot * P;
        const int m = sp.minorder;
        for (int t = 0; t < ncols; ++t) {
          const int j = strip * S + t, idx = i * Pw + j;
          if (with_goal && (int)cp[idx] < cmin) continue;
          const float x = map[idx];
          keep_min(bmask_v, bmask_i, x, idx);
          if (with_goal && m > 0) {
            bool low = true;
            if ((int)i < m || j < m || (int)i + m >= Ph || j + m >= Pw) low = x <= 0.f;
            for (int di = -m; low && di <= m; ++di) {
              const int ii = (int)i + di;
              if (ii < 0 || ii >= Ph) continue;
              for (int dj = -m; dj <= m; ++dj) {
                const int jj = j + dj;
                if (jj < 0 || jj >= Pw) continue;
                if (map[ii * Pw + jj] < x) {
                  low = false;
                  break;
                }
              }
            }
            if (low) keep_min(bmin_v, bmin_i, x, idx);
          }
        }
      }
      cand_v[tid] = bmin_v;
      cand_i[tid] = bmin_i;
      cand_v[nthreads + tid] = bmask_v;
      cand_i[nthreads + tid] = bmask_i;
    }
    __syncthreads();
    // one warp per view reduces the candidates of its items (contiguous threads)
    const int per_slot = p.strips * Ph;
    for (int s = warp; s < Gv * R; s += nwarps) {
      float av = 0.f, bv = 0.f;
      int ai = -1, bi = -1;
      for (int k = lane; k < per_slot; k += 32) {
        const int t = s * per_slot + k;
        if (cand_i[t] >= 0) keep_min(av, ai, cand_v[t], cand_i[t]);
        if (cand_i[nthreads + t] >= 0) keep_min(bv, bi, cand_v[nthreads + t], cand_i[nthreads + t]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float av2 = __shfl_xor_sync(0xffffffffu, av, o);
        const int ai2 = __shfl_xor_sync(0xffffffffu, ai, o);
        const float bv2 = __shfl_xor_sync(0xffffffffu, bv, o);
        const int bi2 = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ai2 >= 0) keep_min(av, ai, av2, ai2);
        if (bi2 >= 0) keep_min(bv, bi, bv2, bi2);
      }
      if (lane == 0) {
        const float v = ai >= 0 ? av : bv;
        const int idx = ai >= 0 ? ai : bi;
        sp.actions[(size_t)e0 * R + s] = idx;
        slot_v[s] = v;
        slot_i[s] = idx;
      }
    }
    __syncthreads();
    // PyGreedy batchwise: first arg-max over views of -value (policies.py:78-80)
    if (sp.best != nullptr && tid < Gv) {
      int br = 0;
      float bv = slot_v[tid * R];
      for (int r = 1; r < R; ++r) {
        const float v = slot_v[tid * R + r];
        if (v < bv) {
          bv = v;
          br = r;
        }
      }
      sp.best[2 * (size_t)(e0 + tid)] = br;
      sp.best[2 * (size_t)(e0 + tid) + 1] = slot_i[tid * R + br];
    }
    // (the next iteration's prep rewrites the compute layout aliased by win/counts
    //  only after its own barriers; cand/slot arrays are rewritten after them too)
    __syncthreads();
  }
  if (tid == 0 && p.stage_out && p.out != nullptr) tma_store_wait_all();
}

template <int T, int VC>
int launch(const ScoreParams& sp, int blocks, int threads, size_t smem, cudaStream_t stream) {
  auto k = score_fused_kernel<T, VC>;
  SRL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<blocks, threads, smem, stream>>>(sp);
  return check_launch("score_fused_kernel");
}

}  // namespace

int score_f32(const float* walls, const float* goals, const float* rocks, const float* level,
              float* values, int64_t* actions, int64_t* best, int E, int R, int H, int W,
              int h, int level_mode, int minorder, double overlap_threshold,
              cudaStream_t stream) {
  SRL_REQUIRE(E >= 0 && R >= 1 && h >= 1 && H >= h && W >= h && minorder >= 0, SRL_E_INVALID,
              "score_f32: bad shape E=%d R=%d H=%d W=%d h=%d minorder=%d", E, R, H, W, h,
              minorder);
  SRL_REQUIRE(level_mode >= 0 && level_mode <= 2, SRL_E_INVALID, "score_f32: level_mode %d",
              level_mode);
  if (E == 0) return SRL_OK;
  SRL_REQUIRE(walls && rocks && actions, SRL_E_INVALID, "score_f32: null pointer");
  SRL_REQUIRE(level_mode != 1 || level, SRL_E_INVALID, "score_f32: level_mode 1 needs level");
  SRL_REQUIRE(level_mode != 2 || goals, SRL_E_INVALID, "score_f32: level_mode 2 needs goals");
  SRL_REQUIRE(h <= 32, SRL_E_UNSUPPORTED, "score_f32: rock side %d > 32", h);

  ScoreParams sp;
  MaxPlusParams& p = sp.mp;
  p.walls = walls; p.rocks = rocks; p.level = level; p.out = values;
  p.E = E; p.R = R; p.H = H; p.W = W; p.h = h;
  p.Ph = H - h + 1; p.Pw = W - h + 1;
  p.threshold = 0.f;
  const int P = p.Ph * p.Pw;
  sp.goals = goals; sp.actions = actions; sp.best = best;
  sp.minorder = minorder; sp.overlap_threshold = overlap_threshold;
  sp.level_mode = level_mode;
  sp.nW = (W + 31) / 32 + 1;
  sp.pf = h > 16 ? 1 : (h > 8 ? 2 : 4);
  sp.hb = 32 / sp.pf;
  sp.ng = (h + sp.pf - 1) / sp.pf;

  const Choice c = choose_tile(p.Pw, h);
  const int T = c.T, VC = c.VC;
  p.hp = round_up(h, VC);
  p.strips = strips_for(p.Pw, T);
  const int nr4 = (T + VC + 2) / 4;
  const int need = (p.strips - 1) * (T - 1) + (p.hp - VC) + 4 * nr4;
  p.Ws = round_up(need > W ? need : W, 4);
  if ((p.Ws / 4) % 2 == 0) p.Ws += 4;
  p.wall_stride = H * p.Ws;
  p.rock_stride = h * p.hp + 4;
  p.tma_wall = (W % 4 == 0) && (((uintptr_t)walls) % 16 == 0) &&
               (!goals || ((uintptr_t)goals) % 16 == 0);
  p.tma_rock = (h % 4 == 0) && (((uintptr_t)rocks) % 16 == 0);
  SRL_REQUIRE(p.tma_wall && p.tma_rock, SRL_E_UNSUPPORTED,
              "score_f32: rows must be multiples of 16 bytes and 16-byte aligned");

  const int sms = sm_count();
  SRL_REQUIRE(sms > 0, SRL_E_CUDA, "score_f32: no CUDA device");
  const int kThreads = 288;
  const size_t kBudget = 112 * 1024;
  const bool stage_out = values != nullptr && ((size_t)R * P) % 4 == 0 &&
                         ((uintptr_t)values) % 16 == 0;
  auto region_for = [&](int G) {
    const int slots = G * R;
    const size_t alias = (size_t)G * H * p.Pw * 4 + (size_t)slots * P * 2;
    const size_t layout = (size_t)G * p.wall_stride * 4 + 2 * (size_t)slots * p.rock_stride * 4;
    return (size_t)round_up((int)(alias > layout ? alias : layout), 16);
  };
  auto smem_for = [&](int G, int threads) {
    const int slots = G * R;
    size_t s = 16 + round_up(2 * slots * 4, 16) + round_up(G * 4, 16) +
               3 * (size_t)round_up(slots * 4, 16) + 2 * (size_t)round_up(2 * threads * 4, 16) +
               round_up(G * H * sp.nW * 4, 16) + round_up(slots * sp.ng * 4, 16);
    s += (goals ? (size_t)G * H * W * 4 : 0) + (size_t)G * H * W * 4 + (size_t)slots * h * h * 4;
    s += region_for(G);
    s += (size_t)slots * P * 4;
    return s;
  };
  auto fits = [&](int G) {
    const int slots = G * R;
    return smem_for(G, kThreads) <= kBudget &&
           G * R * p.strips * p.Ph <= kThreads && (size_t)G * H * W < 65536 &&
           (size_t)slots * h * p.hp < 65536 && (size_t)slots * P < 65536 * 4;
  };
  SRL_REQUIRE(fits(1), SRL_E_UNSUPPORTED,
              "score_f32: shape outside the fused kernel (use the separate kernels)");
  int G = 1;
  while (G < 32 && G < E && fits(G + 1)) ++G;
  if (const char* s = getenv("SRL_MP_G")) {
    const int g = atoi(s);
    if (g >= 1 && fits(g)) G = g;
  }
  p.G = G; p.RC = R; p.rchunks = 1;
  sp.region_bytes = (int)region_for(G);
  p.stage_out = stage_out;
  p.ngroups = (E + G - 1) / G;
  const int threads = round_up(G * R * p.strips * p.Ph, 32);
  const size_t smem = smem_for(G, threads);
  const int blocks = p.ngroups < 2 * sms ? p.ngroups : 2 * sms;
  p.dPh = make_fastdiv(p.Ph); p.dStrips = make_fastdiv(p.strips);
  p.dRC = make_fastdiv(R); p.dW = make_fastdiv(W); p.dH = make_fastdiv(H);
  p.dh = make_fastdiv(h); p.dhp = make_fastdiv(p.hp);
  p.dW4 = make_fastdiv(W / 4); p.dh4 = make_fastdiv(h / 4);
  sp.dPw = make_fastdiv(p.Pw); sp.dNW = make_fastdiv(sp.nW);

#define SRL_SC_CASE(TT, VV) \
  if (T == TT && VC == VV) return launch<TT, VV>(sp, blocks, threads, smem, stream);
#define SRL_SC_ROW(VV)                                                          \
  SRL_SC_CASE(5, VV) SRL_SC_CASE(9, VV) SRL_SC_CASE(13, VV) SRL_SC_CASE(17, VV) \
  SRL_SC_CASE(21, VV) SRL_SC_CASE(25, VV)
  SRL_SC_ROW(4)
  SRL_SC_ROW(8)
  SRL_SC_ROW(16)
#undef SRL_SC_ROW
#undef SRL_SC_CASE
  return fail(SRL_E_UNSUPPORTED, "score_f32: no kernel for T=%d VC=%d", T, VC);
}

}  // namespace srl
