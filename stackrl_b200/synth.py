"""Deterministic synthetic walls, rock-underside maps, goals and rock meshes.

Shared by the tests, the golden-vector generator and bench.py (SURVEY section
8d fixes the shapes and seeds of the BASELINE configs).  numpy only; nothing
here is on the device path.

Geometry follows the reference defaults: a rock image is
``h`` pixels across ``object_max_dimension = 0.125 m``, the wall image is
``H x W`` of the same pixels, ``max_z = 0.375`` and the goal level is
``max_z - object_max_dimension = 0.25``.
"""
import numpy as np

MAX_Z = 0.375
OBJECT_MAX_DIMENSION = 0.125
GOAL_LEVEL = MAX_Z - OBJECT_MAX_DIMENSION


def _box_blur(x):
  """3x3 box blur with edge replication over the last two axes."""
  p = np.pad(x, [(0, 0)] * (x.ndim - 2) + [(1, 1), (1, 1)], mode='edge')
  acc = np.zeros_like(x)
  for di in range(3):
    for dj in range(3):
      acc += p[..., di:di + x.shape[-2], dj:dj + x.shape[-1]]
  return acc / 9


def walls(seed, count, H, W, top=0.3):
  """[count, H, W] float32 smooth-ish wall heightmaps in [0, top)."""
  rng = np.random.default_rng(seed)
  x = rng.uniform(0, top, (count, H, W)).astype('float32')
  return _box_blur(x).astype('float32')


def rocks(seed, count, rotations, side, zero_fraction=0.3):
  """[count, rotations, side, side] float32 underside maps.

  Each rock is the lower half of a random super-ellipsoid seen from below,
  re-sampled (not image-rotated) at ``rotations`` orientations 2*pi*k/R about
  z, like the reference re-renders per orientation (observer.py:128-141).
  Value = distance from the plane ``side`` pixels' half-height above the rock
  centre down to the underside (<= OBJECT_MAX_DIMENSION); background exactly 0.
  ``zero_fraction`` sets the mean background share."""
  rng = np.random.default_rng(seed)
  # Footprint area share = 1 - zero_fraction for a super-ellipse |x/a|^p+|y/b|^p<=1
  share = max(1e-3, 1. - zero_fraction)
  ratio = rng.uniform(0.6, 1.0, count)                 # b / a
  power = rng.uniform(2.0, 4.0, count)
  # area of the super-ellipse relative to the unit square [-.5,.5]^2 (approx.
  # with the p=2 constant; exact share is not needed)
  a = np.sqrt(share / (np.pi * ratio)) * rng.uniform(0.95, 1.05, count)
  b = a * ratio
  depth = rng.uniform(0.3, 0.5, count) * OBJECT_MAX_DIMENSION
  phase = rng.uniform(0, 2 * np.pi, count)
  c = (np.arange(side) + 0.5) / side - 0.5
  yy, xx = np.meshgrid(c, c)                           # row i <-> x, column j <-> y
  out = np.zeros((count, rotations, side, side), dtype='float32')
  for k in range(rotations):
    ang = phase + 2 * np.pi * k / rotations
    ca, sa = np.cos(ang)[:, None, None], np.sin(ang)[:, None, None]
    xr = ca * xx + sa * yy
    yr = -sa * xx + ca * yy
    q = np.abs(xr / a[:, None, None]) ** power[:, None, None] + \
      np.abs(yr / b[:, None, None]) ** power[:, None, None]
    inside = q < 1
    bulge = np.sqrt(np.clip(1 - q, 0, 1)) * depth[:, None, None]
    val = np.where(inside, OBJECT_MAX_DIMENSION / 2 + bulge, 0)
    out[:, k] = val.astype('float32')
  if zero_fraction >= 1:
    out[:] = 0
  return out


def goal_rects(seed, count, H, W, ratio=0.25):
  """[count, 4] int32 goal limits (u0, v0, u1, v1) = Rewarder._goal_lims flattened:
  one rectangle of ~ratio*H*W pixels per map, kept 1/8 away from the borders like
  rewarder.py:239-249."""
  rng = np.random.default_rng(seed)
  rects = np.zeros((count, 4), dtype='int32')
  area = int(ratio * H * W)
  for e in range(count):
    gh = int(rng.integers(max(2, area // W), min(H, max(3, area // 2)) + 1))
    gh = min(gh, H)
    gw = min(max(2, area // gh), W)
    u = int(rng.integers((H - gh) // 8, 7 * (H - gh) // 8 + 1))
    v = int(rng.integers((W - gw) // 8, 7 * (W - gw) // 8 + 1))
    rects[e] = (u, v, u + gh, v + gw)
  return rects


def goals(seed, count, H, W, ratio=0.25, level=GOAL_LEVEL):
  """[count, H, W] float32 goal maps: the rectangles of ``goal_rects`` at ``level``."""
  g = np.zeros((count, H, W), dtype='float32')
  for e, (u0, v0, u1, v1) in enumerate(goal_rects(seed, count, H, W, ratio)):
    g[e, u0:u1, v0:v1] = level
  return g


def to_dtype(x, dtype):
  """StackEnv._return (env.py:171-180)."""
  dtype = np.dtype(dtype)
  if dtype.kind == 'u':
    return np.array(x * (2 ** (8 * dtype.itemsize) - 1) /
                    max(MAX_Z, OBJECT_MAX_DIMENSION), dtype=dtype)
  return np.array(x, dtype=dtype)


def observation(seed, H, W, side, zero_fraction=0.3, quantum=None, flat=False,
                dtype='float32'):
  """One reference-layout observation ([H, W, 2], [side, side, 1])."""
  wall = walls(seed, 1, H, W)[0]
  if flat:
    wall[:] = 0
  rock = rocks(seed + 1000, 1, 1, side, zero_fraction)[0, 0]
  if zero_fraction <= 0:
    rock = np.maximum(rock, np.float32(OBJECT_MAX_DIMENSION / 4))
    rock += np.random.default_rng(seed + 3000).uniform(
      0, 0.02, rock.shape).astype('float32')
  if quantum:
    wall = (np.round(wall / quantum) * quantum).astype('float32')
    rock = (np.round(rock / quantum) * quantum).astype('float32')
  goal = goals(seed + 2000, 1, H, W)[0]
  return (to_dtype(np.stack([wall, goal], axis=-1), dtype),
          to_dtype(rock[:, :, None], dtype))


def batched_observation(seed, H, W, side, rotations, dtype='float32'):
  """TestStackEnv layout ([N, H, W, 2], [N, side, side, 1]): one wall/goal
  repeated for N = rotations views of one rock (env.py:472-480)."""
  wall = walls(seed, 1, H, W)[0]
  goal = goals(seed + 2000, 1, H, W)[0]
  rock = rocks(seed + 1000, 1, rotations, side)[0]
  wg = np.stack([wall, goal], axis=-1)
  return (to_dtype(np.array([wg] * rotations), dtype),
          to_dtype(rock[..., None], dtype))


def placement_batch(seed, envs, rotations, H, W, side, zero_fraction=0.3):
  """Device-layout batch for the max-plus search (BASELINE configs 2/4/5):
  walls [E, H, W], rocks [E, R, side, side], goal level [E] (float32)."""
  return (walls(seed, envs, H, W),
          rocks(seed + 1, envs, rotations, side, zero_fraction),
          np.full((envs,), GOAL_LEVEL, dtype='float32'))
