"""numpy restatement of the reference's placement-scoring functions.

Test infrastructure (see oracle/__init__.py) -- the checker for the CUDA path
and the ``cpu_baseline`` of bench.py, never a product code path.

Every function names the reference lines (relative to /root/reference) whose
arithmetic it restates.  The restatement is pinned against the reference's own
code by tests/golden/make_golden.py (run in the build container) and
tests/test_oracle_golden.py (run anywhere).

Conventions shared with the reference:
  obs = (wall_goal [H, W, 2], rock [h, w, 1]) in the env dtype; wall_goal[..., 0]
  is the wall heightmap, wall_goal[..., 1] the goal map; P = (H-h+1)(W-w+1)
  candidate positions, row-major.
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view
from scipy import ndimage
from scipy import signal

try:
  import cv2
except ImportError:  # pragma: no cover
  cv2 = None


# ---- stackrl/baselines.py:21-26 (get_inputs) -------------------------------- #
def normalise(obs):
  """wall and rock divided by the goal level.  float32 obs stay float32;
  integer obs become float64 (numpy true division), which is why the uint8
  path needs float64 arithmetic downstream."""
  level = obs[0][..., 1].max()
  return obs[0][..., 0] / level, obs[1][..., 0] / level


def _windows(wall, rock):
  """All [h, h] windows of wall (the reference slices BOTH axes with
  rock.shape[0], baselines.py:39 -- quirk Q3 -- so only square rocks are in
  contract)."""
  side = rock.shape[0]
  if rock.shape[0] != rock.shape[1]:
    raise ValueError('reference semantics are only defined for square rocks')
  return sliding_window_view(wall, (side, side))


# ---- stackrl/baselines.py:28-43 (height) ------------------------------------ #
def height(obs, mask=None, **_):
  """Max-plus map: f[i,j] = max_{u,v}( n[u,v] > 0 ? o[i+u,j+v] + n[u,v] : 0 ).
  Returned in a float64 container like the reference's np.zeros default."""
  o, n = normalise(obs)
  win = _windows(o, n)
  cells = np.where(n > 0, win + n, 0)
  f = cells.max(axis=(-2, -1)).astype('float64')
  if mask is not None:
    f = np.where(mask, f, 0.)
  return f


def height_loop(obs):
  """Same result as ``height`` with the reference's cost structure (one numpy
  add/select/max per candidate position inside a Python double loop,
  baselines.py:34-41).  This is the form bench.py times as the CPU baseline."""
  o, n = normalise(obs)
  side = n.shape[0]
  f = np.zeros((o.shape[0] - side + 1, o.shape[1] - side + 1))
  live = n > 0
  for i, j in np.ndindex(*f.shape):
    f[i, j] = np.where(live, o[i:i + side, j:j + side] + n, 0).max()
  return f


# ---- stackrl/baselines.py:45-77 (difference) -------------------------------- #
def difference_weights(n, weights_exponent=2):
  """Radial weights of baselines.py:54-62 (float64, zero off the rock,
  normalised with numpy's pairwise sum over the contiguous [h, w] array)."""
  live = n > 0
  if weights_exponent > 0:
    di = (np.arange(n.shape[0], dtype='float64') - n.shape[0] / 2) ** 2
    dj = (np.arange(n.shape[1], dtype='float64') - n.shape[1] / 2) ** 2
    w = (di[:, None] + dj[None, :]) ** (weights_exponent / 2)
    w = np.where(live, w, 0)
  else:
    w = live.astype('float64')
  return w / w.sum()


def difference(obs, mask=None, difference_exponent=2, weights_exponent=2,
               return_height=False, **_):
  """f[i,j] = sum( w * |h0 - (o_win + n)| ** p ), h0 the max-plus value.
  The per-position np.sum call is kept (not a multi-axis reduction) because
  its pairwise summation order is part of the result's bits."""
  o, n = normalise(obs)
  win = _windows(o, n)
  live = n > 0
  w = difference_weights(n, weights_exponent)
  f = np.zeros(win.shape[:2])
  top = np.zeros_like(f)
  for i, j in np.ndindex(*f.shape):
    if mask is not None and not mask[i, j]:
      continue
    lifted = win[i, j] + n
    h0 = np.where(live, lifted, 0).max()
    f[i, j] = np.sum(w * np.abs(h0 - lifted) ** difference_exponent)
    top[i, j] = h0
  return (f, top) if return_height else f


# ---- stackrl/baselines.py:141-143 (correlate) ------------------------------- #
def correlate(obs, **_):
  o, n = normalise(obs)
  return signal.correlate2d(o, n, mode='valid') / n.sum()


# ---- stackrl/baselines.py:79-114 (corrcoef) --------------------------------- #
def corrcoef(obs, mask=None, localized=False, **_):
  """Normalised cross-correlation coefficient.  Non-localized with OpenCV
  present is cv2.matchTemplate(TM_CCOEFF_NORMED) (baselines.py:84-85);
  otherwise the masked per-position formula of baselines.py:88-112."""
  o, n = normalise(obs)
  if not localized and cv2 is not None:
    return cv2.matchTemplate(o.astype('float32'), n.astype('float32'),
                             cv2.TM_CCOEFF_NORMED)
  live = (n > 0) if localized else np.ones(n.shape, dtype=bool)
  win = _windows(o, n)
  f = np.zeros(win.shape[:2])
  count = np.count_nonzero(live)
  # In-place on purpose: count is a numpy integer, so the mean is float64 and
  # the reference's ``n -= mean`` rounds the float64 difference back to n's
  # own dtype (baselines.py:95), unlike ``n = n - mean``.
  n -= np.sum(np.where(live, n, 0)) / count
  n_var = np.sum(np.where(live, n ** 2, 0))
  if n_var == 0:
    return f
  for i, j in np.ndindex(*f.shape):
    if mask is not None and not mask[i, j]:
      continue
    centred = win[i, j] - np.sum(np.where(live, win[i, j], 0)) / count
    o_var = np.sum(np.where(live, centred ** 2, 0))
    if o_var != 0:
      f[i, j] = np.sum(np.where(live, n * centred, 0)) / np.sqrt(n_var * o_var)
  return f


# ---- stackrl/baselines.py:145-150 (random) ---------------------------------- #
def random(obs, seed=None, **_):
  shape = np.subtract(obs[0].shape, obs[1].shape)[:-1] + 1
  return np.random.default_rng(seed).random(shape)


# ---- stackrl/baselines.py:152-156 (goal_overlap) ---------------------------- #
def overlap_counts(obs):
  """Integer sliding sum of (wall < goal) under the rock footprint (rock > 0),
  on the RAW observation (no normalisation)."""
  below = (obs[0][..., 0] < obs[0][..., 1]).astype('int64')
  foot = (obs[1][..., 0] > 0).astype('int64')
  return signal.correlate2d(below, foot, mode='valid')


def goal_overlap(obs, threshold=0.75, **_):
  counts = overlap_counts(obs)
  return counts >= threshold * counts.max()


METHODS = {
  'random': random,
  'correlate': correlate,
  'height': height,
  'difference': difference,
  'corrcoef': corrcoef,
}


# ---- stackrl/baselines.py:201-217 (Baseline.call) --------------------------- #
def select(values, mask=None, minorder=1):
  """Goal-masked, local-minimum-preferring arg-min and the negated value map.

  mask None is the reference's ``goal=False`` branch (baselines.py:216-217).
  Local minima use a (1+2*minorder)^2 minimum filter padded with ZEROS
  (scipy 'constant' mode, cval 0 -- quirk Q6); np.argmin's first-index
  tie-break is part of the contract."""
  if mask is None:
    return int(np.argmin(values)), -values
  shown = -np.where(mask, values, values[mask].max() + 0.001)
  if minorder:
    size = 1 + 2 * minorder
    lowest = ndimage.minimum_filter(values, size=size, mode='constant') == values
    minima = mask & lowest
    if minima.any():
      return int(np.argmin(np.where(minima, values, np.inf))), shown
  return int(np.argmin(np.where(mask, values, np.inf))), shown


def baseline_call(obs, method='height', goal=True, minorder=1, **kwargs):
  """Baseline.call (baselines.py:201-217): score, goal mask, select."""
  fn = METHODS[method] if isinstance(method, str) else method
  values = fn(obs, **kwargs)
  mask = goal_overlap(obs, **kwargs) if goal else None
  return select(values, mask, minorder)


def baseline_call_loop(obs, goal=True, minorder=1, **kwargs):
  """Baseline('height').call with the reference's cost structure (the Python
  double loop of ``height_loop``); the form bench.py times on the CPU."""
  values = height_loop(obs)
  mask = goal_overlap(obs, **kwargs) if goal else None
  return select(values, mask, minorder)


# ---- stackrl/agents/policies.py:57-91 (PyGreedy.__call__) ------------------- #
def greedy(obs, call, value=False, unravel=False, batched=False,
           batchwise=False):
  """Return conventions of PyGreedy around ``call(obs) -> (argmax, values)``."""
  if batched:
    picks, maps, best = [], [], []
    for item in zip(*obs):
      a, v = call(tuple(item))
      if unravel:
        a = np.unravel_index(a, v.shape)
      else:
        v = v.ravel()
      picks.append(a)
      maps.append(v)
      best.append(v[a])
    out = np.array(picks)
    values = np.array(maps)
    if batchwise:
      k = np.argmax(best)
      out = (k, out[k])
  else:
    out, values = call(obs)
    if unravel:
      out = np.array(np.unravel_index(out, values.shape))
    elif value:
      values = values.ravel()
  return (out, values) if value else out


# ---- stackrl/envs/stack/observer.py:392-421 (Observer.pose) ----------------- #
def drop_height(wall, rock, pixel, threshold=1e-4):
  """Single-position max-plus with the observer's mask threshold (quirk Q4):
  max((wall_window + rock)[rock > 1e-4]) in float32."""
  i, j = pixel
  lifted = wall[i:i + rock.shape[0], j:j + rock.shape[1]] + rock
  return lifted[rock > threshold].max()


def pose(wall, rock, pixel, pixel_size, object_size):
  """(x, y, z) of observer.py:398-413.  pixel_size = (pixel_h, pixel_w);
  object_size = (object_x, object_y, object_z)."""
  x = pixel[0] * pixel_size[0] + object_size[0] / 2
  y = pixel[1] * pixel_size[1] + object_size[1] / 2
  z = drop_height(wall, rock, pixel) - object_size[2] / 2
  return x, y, z
