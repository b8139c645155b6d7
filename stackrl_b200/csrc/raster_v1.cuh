// Round-1 rasteriser (one 256-thread CTA per image, per-pass shuffle broadcast of the
// triangle set-up).  Kept behind SRL_RASTER_MODE=1 as the A/B reference of raster.cu:
// both must produce the same bits.  Included by raster.cu only.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace srl {

namespace v1 {

constexpr int kRasterThreads = 256;
constexpr int kVertCap = 2048;        // cached screen-space vertices per instance
constexpr int kInstCap = 32;          // instances rasterised as one batch (wall images)

// clip = M * (x, y, z, 1) in float64 (left to right), then the viewport
// transform; M is the instance's combined matrix in shared memory (row-major).
__device__ __forceinline__ float3 project(const float* __restrict__ v, const double* M,
                                          int rows, int cols) {
  const double x = v[0], y = v[1], z = v[2];
  const double cx = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[0], x), __dmul_rn(M[1], y)),
                                        __dmul_rn(M[2], z)), M[3]);
  const double cy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[4], x), __dmul_rn(M[5], y)),
                                        __dmul_rn(M[6], z)), M[7]);
  const double cz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[8], x), __dmul_rn(M[9], y)),
                                        __dmul_rn(M[10], z)), M[11]);
  const double cw = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[12], x), __dmul_rn(M[13], y)),
                                        __dmul_rn(M[14], z)), M[15]);
  float3 s;
  s.x = (float)__dmul_rn(__dadd_rn(__dmul_rn(__ddiv_rn(cx, cw), 0.5), 0.5), (double)cols);
  s.y = (float)__dmul_rn(__dadd_rn(0.5, -__dmul_rn(__ddiv_rn(cy, cw), 0.5)), (double)rows);
  s.z = (float)__dadd_rn(__dmul_rn(__ddiv_rn(cz, cw), 0.5), 0.5);
  return s;
}

// M = proj * (view * [rot pos; 0 1]), every entry summed left to right over
// k = 0..3, computed by 16 threads (one entry each) in two steps.
// Entry `e` (0..15) of view * [rot pos; 0 1] and of proj * that.
__device__ __forceinline__ double view_model_entry(const srl_raster_instance& in,
                                                   const srl_raster_job& job, int e) {
  const int r = e >> 2, c = e & 3;
  double a = 0.;
  for (int k = 0; k < 4; ++k) {
    const double t = k < 3 ? (c < 3 ? in.rot[3 * k + c] : in.pos[k]) : (c < 3 ? 0. : 1.);
    const double term = __dmul_rn(job.view[k * 4 + r], t);
    a = k == 0 ? term : __dadd_rn(a, term);
  }
  return a;
}
__device__ __forceinline__ double proj_entry(const double* VT, const srl_raster_job& job,
                                             int e) {
  const int r = e >> 2, c = e & 3;
  double a = 0.;
  for (int k = 0; k < 4; ++k) {
    const double term = __dmul_rn(job.proj[k * 4 + r], VT[4 * k + c]);
    a = k == 0 ? term : __dadd_rn(a, term);
  }
  return a;
}
__device__ __forceinline__ void combine_matrices(double* VT, double* M,
                                                 const srl_raster_instance& in,
                                                 const srl_raster_job& job, int tid) {
  if (tid < 16) VT[tid] = view_model_entry(in, job, tid);
  __syncthreads();
  if (tid < 16) M[tid] = proj_entry(VT, job, tid);
  __syncthreads();
}

struct Tri {
  float x0, y0, d0, x1, y1, d1, x2, y2, d2, area;
  int ilo, ihi, jlo, jhi;
};

__device__ __forceinline__ bool owns_tie(float dx, float dy) {
  return dy > 0.f || (dy == 0.f && dx < 0.f);
}

// Set-up shared with oracle_raster_depth(): returns false for culled triangles.
__device__ __forceinline__ bool setup(Tri& t, int rows, int cols) {
  float area = __fsub_rn(__fmul_rn(__fsub_rn(t.x1, t.x0), __fsub_rn(t.y2, t.y0)),
                         __fmul_rn(__fsub_rn(t.x2, t.x0), __fsub_rn(t.y1, t.y0)));
  if (!(area == area) || area == 0.f) return false;
  if (area < 0.f) {
    float s;
    s = t.x1; t.x1 = t.x2; t.x2 = s;
    s = t.y1; t.y1 = t.y2; t.y2 = s;
    s = t.d1; t.d1 = t.d2; t.d2 = s;
    area = -area;
  }
  t.area = area;
  const float minx = fminf(t.x0, fminf(t.x1, t.x2)), maxx = fmaxf(t.x0, fmaxf(t.x1, t.x2));
  const float miny = fminf(t.y0, fminf(t.y1, t.y2)), maxy = fmaxf(t.y0, fmaxf(t.y1, t.y2));
  if (!(maxx >= 0.f) || !(maxy >= 0.f) || !(minx <= (float)cols) || !(miny <= (float)rows))
    return false;
  // candidates: pixels whose centre lies inside the float32 bounding box
  t.jlo = max((int)ceilf(__fsub_rn(fmaxf(minx, 0.f), 0.5f)), 0);
  t.jhi = min((int)floorf(__fsub_rn(fminf(maxx, (float)cols), 0.5f)), cols - 1);
  t.ilo = max((int)ceilf(__fsub_rn(fmaxf(miny, 0.f), 0.5f)), 0);
  t.ihi = min((int)floorf(__fsub_rn(fminf(maxy, (float)rows), 0.5f)), rows - 1);
  return t.jlo <= t.jhi && t.ilo <= t.ihi;
}

__device__ __forceinline__ void shade(const Tri& t, int i, int j, uint32_t* depth, int cols) {
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const float e01x = __fsub_rn(t.x1, t.x0), e01y = __fsub_rn(t.y1, t.y0);
  const float e12x = __fsub_rn(t.x2, t.x1), e12y = __fsub_rn(t.y2, t.y1);
  const float e20x = __fsub_rn(t.x0, t.x2), e20y = __fsub_rn(t.y0, t.y2);
  const float w2 = __fsub_rn(__fmul_rn(e01x, __fsub_rn(py, t.y0)),
                             __fmul_rn(e01y, __fsub_rn(px, t.x0)));
  const float w0 = __fsub_rn(__fmul_rn(e12x, __fsub_rn(py, t.y1)),
                             __fmul_rn(e12y, __fsub_rn(px, t.x1)));
  const float w1 = __fsub_rn(__fmul_rn(e20x, __fsub_rn(py, t.y2)),
                             __fmul_rn(e20y, __fsub_rn(px, t.x2)));
  if (w0 < 0.f || w1 < 0.f || w2 < 0.f) return;
  if (w2 == 0.f && !owns_tie(e01x, e01y)) return;
  if (w0 == 0.f && !owns_tie(e12x, e12y)) return;
  if (w1 == 0.f && !owns_tie(e20x, e20y)) return;
  float acc = __fmul_rn(w0, t.d0);
  acc = __fadd_rn(acc, __fmul_rn(w1, t.d1));
  acc = __fadd_rn(acc, __fmul_rn(w2, t.d2));
  float d = __fdiv_rn(acc, t.area);
  if (!(d >= 0.f) || d > 1.f) return;
  if (d == 0.f) d = 0.f;
  atomicMin(depth + i * cols + j, __float_as_uint(d));
}


// Rasterise the (up to) 32 triangles held one per lane.  The candidate pixels of
// the 32 bounding boxes form ONE flat work list per warp: an inclusive scan of
// the box sizes gives every triangle its slice, the warp walks the list 32
// entries at a time, each lane finds the owner of its entry by a binary search
// over the scan (shuffles) and fetches that triangle's set-up from the owner's
// registers.  A pass therefore shades 32 pixels whatever the mix of box sizes
// (1-pixel slivers of a 2k-triangle rock next to a box face that covers the whole
// wall image) instead of max-over-lanes box loops with half the lanes idle.
__device__ __forceinline__ void raster_warp_triangles(const Tri& tri, bool valid,
                                                      uint32_t* depth, int cols) {
  constexpr uint32_t kAll = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int bw = 1, npx = 0;
  if (valid) {
    bw = tri.jhi - tri.jlo + 1;
    npx = bw * (tri.ihi - tri.ilo + 1);
  }
  int incl = npx;                                 // inclusive scan over the lanes
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(kAll, incl, d);
    if (lane >= d) incl += up;
  }
  const int total = __shfl_sync(kAll, incl, 31);
  const int excl = incl - npx;
  // row = k / bw through the float reciprocal: exact for k < 2^20 (images are at
  // most 220 KB of shared memory), since (k + 0.5) / bw is >= 0.5 / bw away from an integer.
  const float inv_bw = __frcp_rn((float)bw);
  for (int k0 = 0; k0 < total; k0 += 32) {
    const int k = k0 + lane;
    // owner = first lane whose inclusive scan exceeds k (lanes past `total` idle)
    int lo = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const int probe = __shfl_sync(kAll, incl, lo + step - 1);
      if (probe <= k) lo += step;
    }
    const int src = min(lo, 31);
    Tri b;
    b.x0 = __shfl_sync(kAll, tri.x0, src); b.y0 = __shfl_sync(kAll, tri.y0, src);
    b.d0 = __shfl_sync(kAll, tri.d0, src); b.x1 = __shfl_sync(kAll, tri.x1, src);
    b.y1 = __shfl_sync(kAll, tri.y1, src); b.d1 = __shfl_sync(kAll, tri.d1, src);
    b.x2 = __shfl_sync(kAll, tri.x2, src); b.y2 = __shfl_sync(kAll, tri.y2, src);
    b.d2 = __shfl_sync(kAll, tri.d2, src); b.area = __shfl_sync(kAll, tri.area, src);
    const int ilo = __shfl_sync(kAll, tri.ilo, src);
    const int jlo = __shfl_sync(kAll, tri.jlo, src);
    const int w = __shfl_sync(kAll, bw, src);
    const int first = __shfl_sync(kAll, excl, src);
    const float inv = __shfl_sync(kAll, inv_bw, src);
    if (k < total) {
      const int local = k - first;
      const int row = __float2int_rz(__fmul_rn((float)local + 0.5f, inv));
      shade(b, ilo + row, jlo + (local - row * w), depth, cols);
    }
  }
}

__global__ void __launch_bounds__(kRasterThreads, 4)
raster_kernel(const float* __restrict__ verts, const int32_t* __restrict__ tris,
              const srl_raster_instance* __restrict__ insts,
              const srl_raster_job* __restrict__ jobs,
              const int32_t* __restrict__ inst_counts, float* __restrict__ out, int rows,
              int cols, int mode, double far_plane, int tri_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* VT = reinterpret_cast<double*>(smem_raw);                  // [kInstCap][16]
  double* M = VT + 16 * kInstCap;                                   // [kInstCap][16]
  int* vbase = reinterpret_cast<int*>(M + 16 * kInstCap);           // [kInstCap+1] vertex prefix
  int* tbase = vbase + kInstCap + 1;                                // [kInstCap+1] triangle prefix
  uint32_t* depth = reinterpret_cast<uint32_t*>(tbase + kInstCap + 1 + 2);   // [rows*cols]
  float* sv = reinterpret_cast<float*>(depth + rows * cols);        // [kVertCap*3]
  // triangle indices of the cached vertices (< kVertCap, so 16 bits each), copied
  // with coalesced loads while the vertices are projected: the triangle loop then
  // starts from shared memory instead of a dependent global load per warp pass
  uint16_t* st = reinterpret_cast<uint16_t*>(sv + 3 * kVertCap);    // [tri_cap*3]

  const srl_raster_job& job = jobs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = kRasterThreads / 32;
  const uint32_t one = __float_as_uint(1.0f);
  for (int k = tid; k < rows * cols; k += kRasterThreads) depth[k] = one;

  // Batched path (wall images: a handful of small meshes): all instances share one
  // pass for the matrices, one for the vertices and one flat loop over the
  // triangles, instead of five block barriers per instance.
  const int ninst = inst_counts ? inst_counts[blockIdx.x] : job.inst_count;
  bool batched = ninst >= 2 && ninst <= kInstCap;
  if (batched) {
    if (tid == 0) {
      int nv = 0, nt = 0;
      for (int q = 0; q < ninst; ++q) {
        vbase[q] = nv;
        tbase[q] = nt;
        nv += insts[job.inst_begin + q].vert_count;
        nt += insts[job.inst_begin + q].tri_count;
      }
      vbase[ninst] = nv;
      tbase[ninst] = nt;
    }
    __syncthreads();
    batched = vbase[ninst] <= kVertCap;
  }
  if (batched) {
    for (int k = tid; k < ninst * 16; k += kRasterThreads)
      VT[k] = view_model_entry(insts[job.inst_begin + (k >> 4)], job, k & 15);
    __syncthreads();
    for (int k = tid; k < ninst * 16; k += kRasterThreads)
      M[k] = proj_entry(VT + (k & ~15), job, k & 15);
    __syncthreads();
    const int nv = vbase[ninst], nt = tbase[ninst];
    for (int g = tid; g < nv; g += kRasterThreads) {
      int q = 0;
      while (g >= vbase[q + 1]) ++q;
      const srl_raster_instance& in = insts[job.inst_begin + q];
      const float3 sc = project(verts + 3 * (size_t)(in.vert_begin + g - vbase[q]), M + 16 * q,
                                rows, cols);
      sv[3 * g] = sc.x;
      sv[3 * g + 1] = sc.y;
      sv[3 * g + 2] = sc.z;
    }
    for (int c0 = 0; c0 < nt; c0 += tri_cap) {
      const int cn = min(tri_cap, nt - c0);
      if (c0 > 0) __syncthreads();                 // previous chunk's indices are done with
      for (int tl = tid; tl < cn; tl += kRasterThreads) {
        const int t = c0 + tl;
        int q = 0;
        while (t >= tbase[q + 1]) ++q;
        const srl_raster_instance& in = insts[job.inst_begin + q];
        const int32_t* idx = tris + 3 * (size_t)(in.tri_begin + t - tbase[q]);
        st[3 * tl] = (uint16_t)(vbase[q] + idx[0]);
        st[3 * tl + 1] = (uint16_t)(vbase[q] + idx[1]);
        st[3 * tl + 2] = (uint16_t)(vbase[q] + idx[2]);
      }
      __syncthreads();
      for (int base = warp * 32; base < cn; base += nwarps * 32) {
        const int t = base + lane;
        Tri tri;
        bool valid = t < cn;
        if (valid) {
          const int i0 = st[3 * t], i1 = st[3 * t + 1], i2 = st[3 * t + 2];
          tri.x0 = sv[3 * i0]; tri.y0 = sv[3 * i0 + 1]; tri.d0 = sv[3 * i0 + 2];
          tri.x1 = sv[3 * i1]; tri.y1 = sv[3 * i1 + 1]; tri.d1 = sv[3 * i1 + 2];
          tri.x2 = sv[3 * i2]; tri.y2 = sv[3 * i2 + 1]; tri.d2 = sv[3 * i2 + 2];
          valid = setup(tri, rows, cols);
        }
        raster_warp_triangles(tri, valid, depth, cols);
      }
    }
  } else {
    for (int q = 0; q < ninst; ++q) {
      const srl_raster_instance& in = insts[job.inst_begin + q];
      const bool cached = in.vert_count <= kVertCap;
      __syncthreads();                       // previous instance done with `sv`, M
      combine_matrices(VT, M, in, job, tid);
      if (cached) {
        for (int k = tid; k < in.vert_count; k += kRasterThreads) {
          const float3 s = project(verts + 3 * (size_t)(in.vert_begin + k), M, rows, cols);
          sv[3 * k] = s.x;
          sv[3 * k + 1] = s.y;
          sv[3 * k + 2] = s.z;
        }
        const int32_t* tflat = tris + 3 * (size_t)in.tri_begin;
        for (int c0 = 0; c0 < in.tri_count; c0 += tri_cap) {
          const int cn = min(tri_cap, in.tri_count - c0);
          if (c0 > 0) __syncthreads();
          for (int k = tid; k < 3 * cn; k += kRasterThreads) st[k] = (uint16_t)tflat[3 * (size_t)c0 + k];
          __syncthreads();
          for (int base = warp * 32; base < cn; base += nwarps * 32) {
            const int t = base + lane;
            Tri tri;
            bool valid = t < cn;
            if (valid) {
              const int i0 = st[3 * t], i1 = st[3 * t + 1], i2 = st[3 * t + 2];
              tri.x0 = sv[3 * i0]; tri.y0 = sv[3 * i0 + 1]; tri.d0 = sv[3 * i0 + 2];
              tri.x1 = sv[3 * i1]; tri.y1 = sv[3 * i1 + 1]; tri.d1 = sv[3 * i1 + 2];
              tri.x2 = sv[3 * i2]; tri.y2 = sv[3 * i2 + 1]; tri.d2 = sv[3 * i2 + 2];
              valid = setup(tri, rows, cols);
            }
            raster_warp_triangles(tri, valid, depth, cols);
          }
        }
      } else {
        // mesh too big for the vertex cache: project the three corners per triangle
        __syncthreads();
        for (int base = warp * 32; base < in.tri_count; base += nwarps * 32) {
          const int t = base + lane;
          Tri tri;
          bool valid = t < in.tri_count;
          if (valid) {
            const int32_t* idx = tris + 3 * (size_t)(in.tri_begin + t);
            const float3 a = project(verts + 3 * (size_t)(in.vert_begin + idx[0]), M, rows, cols);
            const float3 b = project(verts + 3 * (size_t)(in.vert_begin + idx[1]), M, rows, cols);
            const float3 c = project(verts + 3 * (size_t)(in.vert_begin + idx[2]), M, rows, cols);
            tri.x0 = a.x; tri.y0 = a.y; tri.d0 = a.z;
            tri.x1 = b.x; tri.y1 = b.y; tri.d1 = b.z;
            tri.x2 = c.x; tri.y2 = c.y; tri.d2 = c.z;
            valid = setup(tri, rows, cols);
          }
          raster_warp_triangles(tri, valid, depth, cols);
        }
      }
    }
  }
  __syncthreads();

  // ---- fused depth -> elevation conversion (float32, numpy's op order) -------- //
  const double far_d = far_plane, oz = job.zrange;
  const float far_f = (float)far_d, oz_f = (float)oz;
  const float c_wall = (float)(far_d * (far_d - oz));                       // observer.py:260
  const float a_rock = (float)(far_d + oz / 2);                             // observer.py:274
  const float b_rock = (float)(far_d * far_d - (oz / 2) * (oz / 2));        // observer.py:275
  float* o = out + (size_t)blockIdx.x * rows * cols;
  for (int k = tid; k < rows * cols; k += kRasterThreads) {
    const float d = __uint_as_float(depth[k]);
    if (mode == SRL_RASTER_DEPTH) {
      o[k] = d;
    } else if (mode == SRL_RASTER_WALL) {
      const float den = __fsub_rn(far_f, __fmul_rn(oz_f, d));
      o[k] = __fsub_rn(far_f, __fdiv_rn(c_wall, den));
    } else {
      const float den = __fadd_rn(far_f, __fmul_rn(oz_f, __fsub_rn(0.5f, d)));
      const float val = __fsub_rn(a_rock, __fdiv_rn(b_rock, den));
      const int i = k / cols, j = k % cols;
      o[i * cols + (cols - 1 - j)] = val;                                   // observer.py:277
    }
  }
}

int raster(const float* verts, const int32_t* tris, const srl_raster_instance* insts,
           const srl_raster_job* jobs, const int32_t* inst_counts, float* out, int njobs,
           int rows, int cols, int mode, double far_plane, cudaStream_t stream) {
  SRL_REQUIRE(njobs >= 0 && rows >= 1 && cols >= 1, SRL_E_INVALID,
              "raster: bad shape njobs=%d rows=%d cols=%d", njobs, rows, cols);
  SRL_REQUIRE(mode >= SRL_RASTER_DEPTH && mode <= SRL_RASTER_ROCK, SRL_E_INVALID,
              "raster: bad mode %d", mode);
  if (njobs == 0) return SRL_OK;
  SRL_REQUIRE(verts && tris && insts && jobs && out, SRL_E_INVALID, "raster: null pointer");
  const size_t base = (size_t)kInstCap * 256 + (2 * (kInstCap + 1) + 2) * 4 +
                      (size_t)rows * cols * 4 + (size_t)kVertCap * 12;
  SRL_REQUIRE(base + 512 * 6 <= 220 * 1024, SRL_E_UNSUPPORTED,
              "raster: %dx%d image exceeds the shared-memory depth tile", rows, cols);
  // index staging: as many triangles per chunk as leave the CTAs per SM unchanged
  auto per_sm = [](size_t bytes) { return (int)std::min<size_t>(4, (227 * 1024) / (bytes + 1024)); };
  int tri_cap = 2048;
  while (tri_cap > 512 && (base + (size_t)tri_cap * 6 > 220 * 1024 ||
                           per_sm(base + (size_t)tri_cap * 6) < per_sm(base + 512 * 6)))
    tri_cap >>= 1;
  const size_t smem = base + (size_t)tri_cap * 6;
  SRL_CUDA(cudaFuncSetAttribute(raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  SRL_CUDA(cudaFuncSetAttribute(raster_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared));
  raster_kernel<<<njobs, kRasterThreads, smem, stream>>>(verts, tris, insts, jobs, inst_counts,
                                                         out, rows, cols, mode, far_plane,
                                                         tri_cap);
  return check_launch("raster_kernel");
}

}  // namespace v1
}  // namespace srl
