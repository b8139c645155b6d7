"""numpy restatement of depth->elevation, observation packing and IoU/OR.

Test infrastructure (see oracle/__init__.py).  Citations are relative to
/root/reference.  Pinned against the reference's own Observer / Rewarder /
StackEnv code by tests/golden (the depth IMAGE itself comes from pybullet in
the reference and is not pinned; the arithmetic applied to it is).
"""
import numpy as np

FAR = 10 ** 3   # Observer.far, observer.py:6


# ---- observer.py:259-260 ---------------------------------------------------- #
def wall_elevation(depth, max_z, far=FAR):
  """Overhead depth image (float32, GL [0,1]) -> elevation above the ground.
  The scalars are Python numbers (weak), so every array op rounds to float32
  at magnitude ~1000: the result is quantised to 2^-14 m."""
  return far - far * (far - max_z) / (far - max_z * depth)


# ---- observer.py:274-277 ---------------------------------------------------- #
def rock_elevation(depth, object_z, far=FAR):
  """Underside depth image -> distance from the plane object_z/2 above the
  rock centre down to the rock's underside; background (depth 1) -> 0.0.
  Columns are mirrored (the camera looks up)."""
  d = far + object_z / 2 - \
    (far ** 2 - (object_z / 2) ** 2) / (far + object_z * (1 / 2 - depth))
  return d[:, ::-1]


# ---- env.py:171-180 (_return) ----------------------------------------------- #
def cast_obs(x, dtype, max_z, object_max_dimension):
  """Float dtypes: plain cast.  uintK: x*(2^K-1)/max(max_z, omd), evaluated in
  the array's float32, then C truncation toward zero (quirk Q11)."""
  dtype = np.dtype(dtype)
  if dtype.kind == 'u':
    levels = 2 ** (8 * dtype.itemsize) - 1
    return np.array(x * levels / max(max_z, object_max_dimension), dtype=dtype)
  return np.array(x, dtype=dtype)


# ---- env.py:226-231 (StackEnv.observation) ---------------------------------- #
def pack_obs(wall, goal, rock, dtype, max_z, object_max_dimension):
  return (
    cast_obs(np.stack([wall, goal], axis=-1), dtype, max_z, object_max_dimension),
    cast_obs(rock[:, :, np.newaxis], dtype, max_z, object_max_dimension),
  )


# ---- env.py:472-480 (TestStackEnv.observation) ------------------------------ #
def pack_obs_batched(wall, goal, rocks, dtype, max_z, object_max_dimension):
  n = len(rocks)
  stacked = np.array([np.stack([wall, goal], axis=-1)] * n)
  shape = (n,) + np.shape(rocks[0]) + (1,)
  return (
    cast_obs(stacked, dtype, max_z, object_max_dimension),
    cast_obs(np.array(rocks).reshape(shape), dtype, max_z, object_max_dimension),
  )


# ---- rewarder.py:297-307 ---------------------------------------------------- #
def intersection(wall, goal, goal_z):
  """sum(min(wall[goal != 0], goal_z)) -- float32, numpy pairwise order over
  the compacted 1-D selection."""
  return np.sum(np.minimum(wall[goal != 0], goal_z))


def union(wall, goal):
  return np.sum(np.maximum(wall, goal))


# ---- rewarder.py:162-179 (call, heightmap metrics only) --------------------- #
def reward(wall, goal, goal_z, metric, memory=0., scale=1.):
  """metric 'iou' -> I/U; 'or' -> I/goal_volume; returns (delta*scale, new
  memory) like Rewarder.call's bookkeeping."""
  if metric == 'iou':
    r = intersection(wall, goal, goal_z) / union(wall, goal)
  elif metric == 'or':
    r = intersection(wall, goal, goal_z) / np.sum(goal)
  else:
    raise ValueError(metric)
  return (r - memory) * scale, r


def contact_precheck(wall, rock, pixel, eps=2. ** -13, threshold=1e-4):
  """Restatement of the heightmap contact pre-check (srl_contact_precheck_f32): the
  residual field h0 - (wall_window + rock) of baselines.py:64-72 at ``pixel``, cells
  within ``eps`` of the contact height counted and sorted into the 8 octants around
  the map centre; supported = at least 3 contacts and no 4 consecutive empty octants
  (the heightmap analogue of Simulator._drop's contact-point count,
  simulator.py:337-341).  -> (contacts, octant mask, supported)."""
  wall = np.asarray(wall, dtype='float32')
  rock = np.asarray(rock, dtype='float32')
  h = rock.shape[0]
  i, j = pixel
  lift = wall[i:i + h, j:j + h] + rock
  live = rock > np.float32(threshold)
  top = lift[live].max()
  touch = live & ((top - lift) <= np.float32(eps))
  mask = 0
  for u, v in zip(*np.nonzero(touch)):
    dx, dy = 2 * int(u) + 1 - h, 2 * int(v) + 1 - h
    ax, ay = abs(dx), abs(dy)
    if dx >= 0 and dy >= 0:
      o = 0 if ax >= ay else 1
    elif dx < 0 and dy >= 0:
      o = 2 if ay > ax else 3
    elif dx < 0 and dy < 0:
      o = 4 if ax >= ay else 5
    else:
      o = 6 if ay > ax else 7
    mask |= 1 << o
  m2 = mask | (mask << 8)
  gap = any(((m2 >> s) & 0xf) == 0 for s in range(8))
  count = int(touch.sum())
  return count, mask, bool(count >= 3 and not gap)
