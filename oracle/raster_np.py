"""Software z-buffer standing in for pybullet's TinyRenderer + GL camera maths.

Test infrastructure (see oracle/__init__.py).  The reference obtains its depth
images from ``pybullet.getCameraImage`` (observer.py:252-257, 267-272), i.e.
Bullet's CPU TinyRenderer -- a third-party dependency that is not under
/root/reference, is unpinned (setup.py:20) and is not installed here.  PARITY
UNPINNED: this module restates the published GL conventions the reference's own
inverse formulas imply (observer.py:259-260) and fixes the free choices
(pixel-centre sampling, tie rule, float32 interpolation order) in
oracle/csrc/oracle.c; the CUDA rasteriser is checked against THIS, and any
raster parity claim must say so.
"""
import ctypes
import os

import numpy as np

from oracle.csrc import build as _build


class Instance(ctypes.Structure):
  _fields_ = [('rot', ctypes.c_double * 9), ('pos', ctypes.c_double * 3),
              ('vert_begin', ctypes.c_int32), ('vert_count', ctypes.c_int32),
              ('tri_begin', ctypes.c_int32), ('tri_count', ctypes.c_int32)]


class Job(ctypes.Structure):
  _fields_ = [('view', ctypes.c_double * 16), ('proj', ctypes.c_double * 16),
              ('inst_begin', ctypes.c_int32), ('inst_count', ctypes.c_int32),
              ('zrange', ctypes.c_double)]


_lib = None


def lib():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(_build.build())
  return _lib


# ---- GL camera matrices, column-major 16-tuples like pybullet returns -------- #
def look_at(eye, target, up):
  """computeViewMatrix: right-handed GL look-at."""
  eye, target, up = (np.asarray(v, dtype='float64') for v in (eye, target, up))
  f = target - eye
  f = f / np.sqrt(f.dot(f))
  u = up / np.sqrt(up.dot(up))
  s = np.cross(f, u)
  s = s / np.sqrt(s.dot(s))
  u = np.cross(s, f)
  m = np.identity(4)
  m[0, :3], m[1, :3], m[2, :3] = s, u, -f
  m[0, 3], m[1, 3], m[2, 3] = -s.dot(eye), -u.dot(eye), f.dot(eye)
  return tuple(m.T.ravel())


def frustum(left, right, bottom, top, near, far):
  """computeProjectionMatrix(left, right, bottom, top, nearVal, farVal): glFrustum."""
  m = np.zeros((4, 4))
  m[0, 0] = 2 * near / (right - left)
  m[1, 1] = 2 * near / (top - bottom)
  m[0, 2] = (right + left) / (right - left)
  m[1, 2] = (top + bottom) / (top - bottom)
  m[2, 2] = -(far + near) / (far - near)
  m[2, 3] = -2 * far * near / (far - near)
  m[3, 2] = -1
  return tuple(m.T.ravel())


# ---- quaternions [x, y, z, w] ---------------------------------------------- #
def quat_from_euler(rpy):
  r, p, y = (0.5 * a for a in rpy)
  cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
  return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
          cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def quat_mul(a, b):
  ax, ay, az, aw = a
  bx, by, bz, bw = b
  return (aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
          aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz)


def quat_conj(q):
  return (-q[0], -q[1], -q[2], q[3])


def quat_matrix(q):
  x, y, z, w = q
  n = x * x + y * y + z * z + w * w
  s = 2.0 / n
  return np.array([
    [1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
    [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
    [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


def quat_rotate(q, v):
  return tuple(quat_matrix(q).dot(np.asarray(v, dtype='float64')))


# ---- rasterisation --------------------------------------------------------- #
def render_depth(view, proj, rows, cols, bodies):
  """bodies: iterable of (verts [V,3] f32 local, tris [T,3] i32, rot [3,3], pos [3]).
  Returns the GL depth image [rows, cols] float32 (background 1.0)."""
  bodies = list(bodies)
  verts = np.concatenate([np.asarray(b[0], dtype='float32').reshape(-1, 3) for b in bodies]) \
    if bodies else np.zeros((0, 3), 'float32')
  tris = np.concatenate([np.asarray(b[1], dtype='int32').reshape(-1, 3) for b in bodies]) \
    if bodies else np.zeros((0, 3), 'int32')
  insts = (Instance * max(1, len(bodies)))()
  vb = tb = 0
  for k, (v, t, rot, pos) in enumerate(bodies):
    nv, nt = len(np.asarray(v).reshape(-1, 3)), len(np.asarray(t).reshape(-1, 3))
    insts[k].rot[:] = list(np.asarray(rot, dtype='float64').ravel())
    insts[k].pos[:] = list(np.asarray(pos, dtype='float64').ravel())
    insts[k].vert_begin, insts[k].vert_count = vb, nv
    insts[k].tri_begin, insts[k].tri_count = tb, nt
    vb += nv
    tb += nt
  job = Job()
  job.view[:] = list(view)
  job.proj[:] = list(proj)
  job.inst_begin, job.inst_count = 0, len(bodies)
  depth = np.empty((rows, cols), dtype='float32')
  verts = np.ascontiguousarray(verts)
  tris = np.ascontiguousarray(tris)
  lib().oracle_raster_depth(
    verts.ctypes.data_as(ctypes.c_void_p), tris.ctypes.data_as(ctypes.c_void_p),
    ctypes.byref(insts), ctypes.byref(job), depth.ctypes.data_as(ctypes.c_void_p),
    ctypes.c_int(rows), ctypes.c_int(cols))
  return depth


def maxplus_f32(wall, rocks, level, threshold=0.):
  """C restatement of baselines.height over R rocks of one environment."""
  wall = np.ascontiguousarray(wall, dtype='float32')
  rocks = np.ascontiguousarray(rocks, dtype='float32')
  R, h = rocks.shape[0], rocks.shape[1]
  H, W = wall.shape
  out = np.empty((R, H - h + 1, W - h + 1), dtype='float32')
  lib().oracle_maxplus_f32(
    wall.ctypes.data_as(ctypes.c_void_p), rocks.ctypes.data_as(ctypes.c_void_p),
    ctypes.c_float(level if level is not None else -1.), ctypes.c_float(threshold),
    out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(R), ctypes.c_int(H),
    ctypes.c_int(W), ctypes.c_int(h))
  return out
