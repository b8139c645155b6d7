/*
 * Plain-C restatement of the stackrl observation + placement-scoring arithmetic.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for the CUDA path
 * at sizes the numpy restatement is too slow for, and the software z-buffer
 * that stands in for pybullet's TinyRenderer (which is NOT part of the reference
 * tree: raster parity is UNPINNED).  Never linked into the
 * product library.  Build: oracle/csrc/build.py (gcc -O2 -ffp-contract=off).
 *
 * Citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- stackrl/baselines.py:21-43 (get_inputs + height), float32 path -------- */
/* out[r,i,j] = max_{u,v}( n > thr ? o[i+u,j+v] + n : 0 ), o = wall/level,
 * n = rock/level (level <= 0 means "no normalisation").  One environment. */
void oracle_maxplus_f32(const float* wall, const float* rocks, float level,
                        float thr, float* out, int R, int H, int W, int h) {
  const int Ph = H - h + 1, Pw = W - h + 1;
  float* o = (float*)malloc(sizeof(float) * H * W);
  float* n = (float*)malloc(sizeof(float) * h * h);
  for (int k = 0; k < H * W; ++k) o[k] = level > 0 ? wall[k] / level : wall[k];
  for (int r = 0; r < R; ++r) {
    for (int k = 0; k < h * h; ++k)
      n[k] = level > 0 ? rocks[r * h * h + k] / level : rocks[r * h * h + k];
    for (int i = 0; i < Ph; ++i)
      for (int j = 0; j < Pw; ++j) {
        float best = -INFINITY;
        int any_dead = 0;
        for (int u = 0; u < h; ++u)
          for (int v = 0; v < h; ++v) {
            const float nv = n[u * h + v];
            if (nv > thr) {
              const float s = o[(i + u) * W + j + v] + nv;
              if (s > best) best = s;
            } else {
              any_dead = 1;
            }
          }
        if (any_dead && !(best > 0.f)) best = 0.f;
        out[((size_t)r * Ph + i) * Pw + j] = best;
      }
  }
  free(o);
  free(n);
}

/* ---- software z-buffer (stand-in for pybullet.getCameraImage) -------------- */
/* Geometry contract shared with stackrl_b200/csrc/raster.cu (DESIGN.md "raster"):
 *  - per instance, in float64: M = proj * (view * [rot pos; 0 1]), each entry
 *    summed left to right; per vertex, in float64, left to right:
 *      clip = M * (x, y, z, 1)
 *      sx = (clip.x / clip.w * 0.5 + 0.5) * cols        (column coordinate)
 *      sy = (0.5 - clip.y / clip.w * 0.5) * rows        (row coordinate)
 *      d  = clip.z / clip.w * 0.5 + 0.5                 (GL depth in [0,1])
 *    then each rounded to float32;
 *  - per triangle, in float32: signed area; clockwise/counter-clockwise both
 *    drawn (vertices 1,2 swapped when the area is negative);
 *  - a pixel (i, j) is sampled at its centre (j + 0.5, i + 0.5); candidates are
 *    the pixels whose centre lies in the triangle's float32 bounding box
 *    (ceil(min - 0.5) .. floor(max - 0.5)); a candidate is covered when the
 *    three edge functions are >= 0 with a top-left style tie rule;
 *  - its depth is (w0*d0 + w1*d1 + w2*d2) / area, evaluated left to right with
 *    separate float32 multiplies and adds; fragments outside [0, 1] are clipped;
 *  - the image keeps the minimum depth per pixel; background 1.0. */
typedef struct {
  double rot[9];
  double pos[3];
  int32_t vert_begin, vert_count, tri_begin, tri_count;
} oracle_instance;

typedef struct {
  double view[16]; /* column-major */
  double proj[16]; /* column-major */
  int32_t inst_begin, inst_count;
  double zrange;
} oracle_job;

static int edge_owns_tie(float dx, float dy) { return dy > 0.f || (dy == 0.f && dx < 0.f); }

void oracle_raster_depth(const float* verts, const int32_t* tris,
                         const oracle_instance* insts, const oracle_job* job,
                         float* depth, int rows, int cols) {
  for (int k = 0; k < rows * cols; ++k) depth[k] = 1.0f;
  for (int q = 0; q < job->inst_count; ++q) {
    const oracle_instance* in = insts + job->inst_begin + q;
    float* s = (float*)malloc(sizeof(float) * 3 * (in->vert_count > 0 ? in->vert_count : 1));
    /* M = proj * (view * [rot pos; 0 1]) in float64, every entry summed left to
     * right over k = 0..3 (matrices: view/proj column-major, rot row-major). */
    double Tm[4][4], VT[4][4], M[4][4];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) Tm[r][c] = in->rot[3 * r + c];
      Tm[r][3] = in->pos[r];
    }
    Tm[3][0] = Tm[3][1] = Tm[3][2] = 0.0;
    Tm[3][3] = 1.0;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double a = job->view[0 * 4 + r] * Tm[0][c];
        a = a + job->view[1 * 4 + r] * Tm[1][c];
        a = a + job->view[2 * 4 + r] * Tm[2][c];
        a = a + job->view[3 * 4 + r] * Tm[3][c];
        VT[r][c] = a;
      }
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double a = job->proj[0 * 4 + r] * VT[0][c];
        a = a + job->proj[1 * 4 + r] * VT[1][c];
        a = a + job->proj[2 * 4 + r] * VT[2][c];
        a = a + job->proj[3 * 4 + r] * VT[3][c];
        M[r][c] = a;
      }
    for (int k = 0; k < in->vert_count; ++k) {
      const float* v = verts + 3 * (size_t)(in->vert_begin + k);
      const double x = v[0], y = v[1], z = v[2];
      const double cx = M[0][0] * x + M[0][1] * y + M[0][2] * z + M[0][3];
      const double cy = M[1][0] * x + M[1][1] * y + M[1][2] * z + M[1][3];
      const double cz = M[2][0] * x + M[2][1] * y + M[2][2] * z + M[2][3];
      const double cw = M[3][0] * x + M[3][1] * y + M[3][2] * z + M[3][3];
      s[3 * k + 0] = (float)((cx / cw * 0.5 + 0.5) * cols);
      s[3 * k + 1] = (float)((0.5 - cy / cw * 0.5) * rows);
      s[3 * k + 2] = (float)(cz / cw * 0.5 + 0.5);
    }
    for (int t = 0; t < in->tri_count; ++t) {
      const int32_t* tri = tris + 3 * (size_t)(in->tri_begin + t);
      int i0 = tri[0], i1 = tri[1], i2 = tri[2];
      float x0 = s[3 * i0], y0 = s[3 * i0 + 1], d0 = s[3 * i0 + 2];
      float x1 = s[3 * i1], y1 = s[3 * i1 + 1], d1 = s[3 * i1 + 2];
      float x2 = s[3 * i2], y2 = s[3 * i2 + 1], d2 = s[3 * i2 + 2];
      float area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
      if (!(area == area) || area == 0.f) continue;
      if (area < 0.f) {
        float t_;
        t_ = x1; x1 = x2; x2 = t_;
        t_ = y1; y1 = y2; y2 = t_;
        t_ = d1; d1 = d2; d2 = t_;
        area = -area;
      }
      float minx = fminf(x0, fminf(x1, x2)), maxx = fmaxf(x0, fmaxf(x1, x2));
      float miny = fminf(y0, fminf(y1, y2)), maxy = fmaxf(y0, fmaxf(y1, y2));
      if (!(maxx >= 0.f) || !(maxy >= 0.f) || !(minx <= (float)cols) ||
          !(miny <= (float)rows))
        continue;
      /* pixels whose centre (j + 0.5, i + 0.5) lies inside the bounding box */
      int jlo = (int)ceilf(fmaxf(minx, 0.f) - 0.5f), jhi = (int)floorf(fminf(maxx, (float)cols) - 0.5f);
      int ilo = (int)ceilf(fmaxf(miny, 0.f) - 0.5f), ihi = (int)floorf(fminf(maxy, (float)rows) - 0.5f);
      if (jlo < 0) jlo = 0;
      if (ilo < 0) ilo = 0;
      if (jhi > cols - 1) jhi = cols - 1;
      if (ihi > rows - 1) ihi = rows - 1;
      const float e01x = x1 - x0, e01y = y1 - y0;
      const float e12x = x2 - x1, e12y = y2 - y1;
      const float e20x = x0 - x2, e20y = y0 - y2;
      for (int i = ilo; i <= ihi; ++i)
        for (int j = jlo; j <= jhi; ++j) {
          const float px = (float)j + 0.5f, py = (float)i + 0.5f;
          const float w2 = e01x * (py - y0) - e01y * (px - x0);
          const float w0 = e12x * (py - y1) - e12y * (px - x1);
          const float w1 = e20x * (py - y2) - e20y * (px - x2);
          if (w0 < 0.f || w1 < 0.f || w2 < 0.f) continue;
          if (w2 == 0.f && !edge_owns_tie(e01x, e01y)) continue;
          if (w0 == 0.f && !edge_owns_tie(e12x, e12y)) continue;
          if (w1 == 0.f && !edge_owns_tie(e20x, e20y)) continue;
          float acc = w0 * d0;
          acc = acc + w1 * d1;
          acc = acc + w2 * d2;
          float d = acc / area;
          if (!(d >= 0.f) || d > 1.f) continue;
          if (d == 0.f) d = 0.f; /* +0: the device keeps depths as ordered uint bits */
          if (d < depth[i * cols + j]) depth[i * cols + j] = d;
        }
    }
    free(s);
  }
}
