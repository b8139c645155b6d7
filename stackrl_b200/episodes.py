"""Host-side randomness of the environments: rock order and goal rectangle.

The reference keeps TWO numpy RandomState streams per environment: the env's own
(``StackEnv._random``: the episode's rock list, env.py:105, 268-272) and the
rewarder's (``Rewarder._random``: the goal rectangle, rewarder.py:211-259),
seeded with ``randint(2**32)`` drawn from the first (env.py:166 at construction,
env.py:340-346 in ``seed()``).  ``ParallelEnv`` seeds environment i with
``seed + i`` (utils.py:433, 530-532).  ``EpisodeSampler`` reproduces those draws
for E environments (pure numpy, no device), or -- ``vector=True`` -- replaces the
2E streams by one vectorised stream with the same distributions for very large
batches (not the reference's draw sequence).
"""
import numpy as np

MARGIN_FACTOR = 8          # Rewarder.margin_factor (rewarder.py:16)


def goal_rectangle(rng, shape, object_shape, goal_size_ratio):
  """One ``Rewarder._reset_goal`` draw (rewarder.py:211-250) from ``rng``:
  returns (u, v, h, w), the goal is rows u:u+h, columns v:v+w."""
  H, W = shape
  min_h, min_w = object_shape
  max_h, max_w = H, W
  ratio = goal_size_ratio
  if not ratio:
    b = 1 + rng.randint(2) * 2
    h = int(min_h + rng.beta(b, 4 - b) * (min_h - min_h))       # quirk Q13
    w = int(min_w + rng.beta(4 - b, b) * (max_w - min_w))
  elif np.isscalar(ratio):
    size = int(ratio * H * W)
    min_h = max(min_h, size // max_w)                          # rewarder.py:76-78
    max_h = min(max_h, size // min_w)
    b = 1 + rng.randint(2) * 2
    h = int(min_h + rng.beta(b, 4 - b) * (max_h - min_h))
    w = min(max(min_w, size // h), max_w)
  else:
    size = tuple(int(g * s) for g, s in zip(ratio, (H, W)))     # rewarder.py:83-85
    i = rng.randint(2)
    h, w = min(size[i], max_h), min(size[1 - i], max_w)
  u_max, v_max = H - h, W - w
  m = MARGIN_FACTOR
  u = rng.randint(u_max // m, (m - 1) * u_max // m + 1)
  v = rng.randint(v_max // m, (m - 1) * v_max // m + 1)
  return u, v, h, w


class EpisodeSampler(object):
  def __init__(self, envs, n_meshes, length, shape, object_shape, goal_size_ratio=.25,
               vector=False):
    self.E, self.M, self.L = int(envs), int(n_meshes), int(length)
    self.shape, self.object_shape = tuple(shape), tuple(object_shape)
    self.ratio = goal_size_ratio
    self.replace = self.M < self.L                              # env.py:103
    self.vector = bool(vector)

  def seed(self, seed=None):
    if seed is None:
      seed = int(np.random.SeedSequence().generate_state(1)[0])
    if self.vector:
      self._rng = np.random.RandomState(seed % 2 ** 32)
      self.rngs = self.goal_rngs = None
    else:
      self.rngs = [np.random.RandomState((seed + i) % 2 ** 32) for i in range(self.E)]
      # StackEnv.seed -> Rewarder.seed(self._random.randint(2**32)) (env.py:340-346)
      self.goal_seeds = [int(r.randint(2 ** 32)) for r in self.rngs]
      self.goal_rngs = [np.random.RandomState(g) for g in self.goal_seeds]
    return seed

  # -- rock orders (env.py:268-272): [n, L] in the order ``choice`` returned them ------- #
  def orders(self, ids):
    if self.vector:
      return self._orders_vector(len(ids))
    return np.stack([self.rngs[e].choice(self.M, size=self.L, replace=self.replace)
                     for e in ids]).astype('int64')

  def _orders_vector(self, n):
    rng, M, L = self._rng, self.M, self.L
    if self.replace:
      return rng.randint(M, size=(n, L)).astype('int64')
    # without replacement: the first L steps of a Fisher-Yates shuffle, all rows at once
    deck = np.tile(np.arange(M, dtype='int64'), (n, 1))
    rows = np.arange(n)
    u = rng.random_sample((L, n))
    for k in range(L):
      j = k + (u[k] * (M - k)).astype('int64')
      picked = deck[rows, j]
      deck[rows, j] = deck[rows, k]
      deck[rows, k] = picked
    return deck[:, :L].copy()

  # -- goal rectangles: [n, 2, 2] = ((u, v), (u + h, v + w)) = Rewarder._goal_lims ------- #
  def goals(self, ids):
    if self.vector:
      return self._goals_vector(len(ids))
    lims = np.empty((len(ids), 2, 2), dtype='int64')
    for k, e in enumerate(ids):
      u, v, h, w = goal_rectangle(self.goal_rngs[e], self.shape, self.object_shape, self.ratio)
      lims[k] = ((u, v), (u + h, v + w))
    return lims

  def _goals_vector(self, n):
    rng = self._rng
    H, W = self.shape
    min_h, min_w = self.object_shape
    max_h, max_w = H, W
    ratio = self.ratio
    if not ratio:
      b = 1 + rng.randint(2, size=n) * 2
      h = np.full(n, min_h, dtype='int64')
      w = (min_w + rng.beta(4 - b, b) * (max_w - min_w)).astype('int64')
    elif np.isscalar(ratio):
      size = int(ratio * H * W)
      min_h = max(min_h, size // max_w)
      max_h = min(max_h, size // min_w)
      b = 1 + rng.randint(2, size=n) * 2
      h = (min_h + rng.beta(b, 4 - b) * (max_h - min_h)).astype('int64')
      w = np.minimum(np.maximum(min_w, size // h), max_w)
    else:
      size = tuple(int(g * s) for g, s in zip(ratio, (H, W)))
      i = rng.randint(2, size=n)
      h = np.minimum(np.where(i == 0, size[0], size[1]), max_h)
      w = np.minimum(np.where(i == 0, size[1], size[0]), max_w)
    m = MARGIN_FACTOR
    u_max, v_max = H - h, W - w
    lo_u, hi_u = u_max // m, (m - 1) * u_max // m + 1
    lo_v, hi_v = v_max // m, (m - 1) * v_max // m + 1
    u = lo_u + (rng.random_sample(n) * (hi_u - lo_u)).astype('int64')
    v = lo_v + (rng.random_sample(n) * (hi_v - lo_v)).astype('int64')
    return np.stack([np.stack([u, v], -1), np.stack([u + h, v + w], -1)], 1).astype('int64')
