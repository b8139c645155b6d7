"""Batched stacking environment with device-resident observations.

The reference runs one ``StackEnv`` (stackrl/envs/stack/env.py) per process and
batches N of them over ``multiprocessing.Pipe`` (``ParallelEnv``,
stackrl/envs/utils.py:302-576).  ``BatchedStackEnv`` keeps that contract --
``reset() / step(action) -> (observation, reward, terminal)`` with a leading
batch axis, ``batch_size``, ``observation_spec``, ``action_spec``, ``sample()``,
``seed()`` -- for E environments on one GPU: observation capture, placement
pose, rewards and packing are the sm_100a kernels; the rigid-body settle step is
NOT part of this package (BASELINE north_star: it stays the reference's pybullet
code) and is plugged in as ``settle``: a callable that receives the placement
poses and returns where the rocks came to rest.  The default ``settle=None``
leaves every rock where it was placed, which is also what the golden episodes
(reference env on the static fake backend) do.

Host-side randomness (rock order, goal rectangle) uses numpy RandomState like
the reference (env.py:105, rewarder.py:211-259); the streams are per
environment, seeded ``seed + i`` (utils.py:433).
"""
import numpy as np
import torch

from stackrl_b200 import capi
from stackrl_b200.baselines import PlacementScorer
from stackrl_b200.camera import HEIGHT_QUANTUM_LOG2
from stackrl_b200.observer import BatchedObserver


class BatchedStackEnv(object):
  metadata = {'dtypes': ['uint8', 'float32']}

  def __init__(self, bank, batch_size, episode_length=30, object_max_dimension=0.125,
               observable_size_ratio=4, resolution_factor=5, max_z=0.375,
               rewarder=None, goal_size_ratio=.25, reward_scale=1., reward_params=None,
               orientation_freedom=0, dtype='float32', settle=None, seed=None,
               device=None):
    """Arguments follow StackEnv (env.py:28-50); ``bank`` is the MeshBank of
    candidate rocks (the reference's ``urdfs`` list), ``orientation_freedom``
    the TestStackEnv option (env.py:443-463), ``settle`` the physics hook."""
    if dtype not in self.metadata['dtypes']:
      raise ValueError('Invalid value {} for argument dtype.'.format(dtype))
    if len(bank) == 0:
      raise AssertionError('List of object descriptor files is empty.')
    self.bank = bank
    self.E = int(batch_size)
    self._length = int(episode_length)
    self._replace = len(bank) < self._length
    self._dtype = dtype
    self._omd = object_max_dimension
    self._max_z = max_z
    object_resolution = 2 ** resolution_factor
    overhead_resolution = object_resolution * observable_size_ratio \
      if np.isscalar(observable_size_ratio) else \
      [object_resolution * r for r in observable_size_ratio[:2]]
    self.obs = BatchedObserver(
      bank, self.E, self._length, overhead_resolution, object_resolution,
      object_max_dimension / object_resolution, max_z, orientation_freedom,
      spawn_pose=((0., 0., max_z + object_max_dimension), (0., 0., 0., 1.)), device=device)
    self.dev = self.obs.dev
    g = self.obs.geo
    self.R = g.n_orientations
    self._settle = settle
    # -- rewards (rewarder.py:17-142) --------------------------------------------- #
    metric = 'iou' if rewarder is None else str(rewarder).lower()
    if metric not in ('iou', 'or', 'dor', 'diou'):
      raise ValueError('Invalid value {} for argument metric'.format(rewarder))
    self.metric = metric
    self.scale = float(reward_scale) if reward_scale is not None else float(episode_length)
    if reward_params is None:
      self._pexp = self._oexp = None
    elif np.isscalar(reward_params):
      self._pexp = self._oexp = reward_params
    else:
      self._pexp, self._oexp = (list(reward_params) * 2)[:2]
    self._pmax = max(g.object_h * g.pixel_h, g.object_w * g.pixel_w)
    self._goal_z = g.max_z
    self._goal_size_ratio = goal_size_ratio
    H, W = g.overhead_h, g.overhead_w
    self.goals = torch.zeros((self.E, H, W), dtype=torch.float32, device=self.dev)
    self._goal_z_d = torch.full((self.E,), self._goal_z, dtype=torch.float32, device=self.dev)
    self.goal_lims = np.zeros((self.E, 2, 2), dtype='int64')
    self._memory = np.zeros(self.E, dtype='float64')
    # Per-environment episode state as arrays (no Python loop over E per step):
    # rest pose / placement pose of every placed rock, rock order and cursor.
    self._rest = np.zeros((self.E, self._length, 3), dtype='float64')
    self._placed_at = np.zeros((self.E, self._length, 3), dtype='float64')
    self._quats = np.zeros((self.E, self._length, 4), dtype='float64')
    self._placed_quats = np.zeros((self.E, self._length, 4), dtype='float64')
    self._n_placed = np.zeros(self.E, dtype='int64')
    self._Ph, self._Pw = H - g.object_h + 1, W - g.object_w + 1
    self.seed(seed)
    self._done = np.ones(self.E, dtype=bool)
    self._order = np.zeros((self.E, self._length), dtype='int64')
    self._cursor = np.zeros(self.E, dtype='int64')     # rocks consumed so far
    self._current = np.zeros(self.E, dtype='int64')

  # -- ParallelEnv-style metadata (utils.py:185-300) --------------------------------- #
  @property
  def batch_size(self):
    return self.E

  @property
  def multiprocessing(self):
    return False

  @property
  def observation_spec(self):
    g = self.obs.geo
    lead = (self.R,) if self.R > 1 else ()
    return ((lead + (g.overhead_h, g.overhead_w, 2), self._dtype),
            (lead + (g.object_h, g.object_w, 1), self._dtype))

  @property
  def action_spec(self):
    n = self._Ph * self._Pw
    return ((self.R, n), 'int64') if self.R > 1 else (n, 'int64')

  def seed(self, seed=None):
    """Per-environment streams seeded seed + i (utils.py:433, 530-532)."""
    if seed is None:
      seed = int(np.random.SeedSequence().generate_state(1)[0])
    self._rngs = [np.random.RandomState((seed + i) % 2 ** 32) for i in range(self.E)]
    self._action_rng = np.random.RandomState(seed % 2 ** 32)
    return [seed]

  def sample(self):
    n = self._Ph * self._Pw
    flat = torch.from_numpy(self._action_rng.randint(n, size=self.E))
    if self.R > 1:
      return torch.from_numpy(self._action_rng.randint(self.R, size=self.E)), flat
    return flat

  # -- goal (rewarder.py:211-259) ------------------------------------------------------ #
  def _new_goal(self, rng):
    g = self.obs.geo
    H, W = g.overhead_h, g.overhead_w
    min_h, min_w, max_h, max_w = g.object_h, g.object_w, H, W
    ratio = self._goal_size_ratio
    if not ratio:
      b = 1 + rng.randint(2) * 2
      h = int(min_h + rng.beta(b, 4 - b) * (min_h - min_h))       # quirk Q13
      w = int(min_w + rng.beta(4 - b, b) * (max_w - min_w))
    elif np.isscalar(ratio):
      size = int(ratio * H * W)
      min_h = max(min_h, size // max_w)
      max_h = min(max_h, size // min_w)
      b = 1 + rng.randint(2) * 2
      h = int(min_h + rng.beta(b, 4 - b) * (max_h - min_h))
      w = min(max(min_w, size // h), max_w)
    else:
      size = tuple(int(s * r) for s, r in zip(ratio, (H, W)))
      i = rng.randint(2)
      h, w = min(size[i], max_h), min(size[1 - i], max_w)
    u_max, v_max = H - h, W - w
    u = rng.randint(u_max // 8, 7 * u_max // 8 + 1)
    v = rng.randint(v_max // 8, 7 * v_max // 8 + 1)
    return u, v, h, w

  def set_goals(self, lims, env_ids=None):
    """Install goal rectangles ((u, v), (u+h, v+w)) (rewarder.py:252-258)."""
    env_ids = range(self.E) if env_ids is None else env_ids
    goals = np.zeros((len(lims),) + tuple(self.goals.shape[1:]), dtype='float32')
    for k, ((u0, v0), (u1, v1)) in enumerate(lims):
      goals[k, u0:u1, v0:v1] = self._goal_z
    ids = torch.as_tensor(list(env_ids), device=self.dev, dtype=torch.long)
    self.goals[ids] = torch.from_numpy(goals).to(self.dev)
    self.goal_lims[list(env_ids)] = np.asarray(lims, dtype='int64')

  # -- episode control ----------------------------------------------------------------- #
  def reset(self, env_ids=None, rock_orders=None, goal_lims=None):
    """Start new episodes (env.py:266-293).  ``rock_orders`` / ``goal_lims``
    override the random draws (used to replay recorded episodes)."""
    ids = list(range(self.E)) if env_ids is None else list(env_ids)
    lims = []
    for k, e in enumerate(ids):
      rng = self._rngs[e]
      if rock_orders is not None:
        order = list(rock_orders[k])
      else:
        order = list(rng.choice(len(self.bank), size=self._length, replace=self._replace))
      self._order[e] = order[::-1]       # the reference pops from the end (env.py:245)
      self._cursor[e] = 1
      self._current[e] = self._order[e, 0]
      if goal_lims is not None:
        lims.append(goal_lims[k])
      else:
        u, v, h, w = self._new_goal(rng)
        lims.append(((u, v), (u + h, v + w)))
      self._n_placed[e] = 0
      self._memory[e] = 0.
      self._done[e] = False
    self.set_goals(lims, ids)
    self.obs.reset(None if env_ids is None else ids)
    self.obs.observe_walls()
    self.obs.observe_rocks(self._current)
    return self.observation, torch.zeros(self.E, device=self.dev), \
      torch.zeros(self.E, dtype=torch.bool, device=self.dev)

  @property
  def observation(self):
    """Packed observation in the env dtype (env.py:226-231; the TestStackEnv
    layout of env.py:472-480 when orientation_freedom > 0)."""
    scale = max(self._max_z, self._omd)
    wall_goal, rock = capi.pack_obs(self.obs.walls, self.goals, self.obs.rocks,
                                    dtype=self._dtype, scale=scale,
                                    repeat_wall=self.R > 1)
    if self.R == 1:
      rock = rock[:, 0]
    return wall_goal, rock

  def planes(self):
    """Planar float32 maps for device-side scoring: (walls, goals, rocks)."""
    return self.obs.walls, self.goals, self.obs.rocks

  def step(self, action):
    """action: [E] flat indices, or (views [E], flat indices [E]) when
    orientation_freedom > 0 (env.py:233-264, 482-520)."""
    if self._done.any():
      raise RuntimeError('reset() the finished environments before stepping them')
    if self.R > 1:
      views, flat = action
    else:
      views, flat = np.zeros(self.E, dtype='int64'), action
    positions, quats = self.obs.poses(views, flat)
    place_positions, place_quats = positions.copy(), quats.copy()
    if self._settle is not None:
      positions, quats = self._settle(self._current.copy(), positions, quats)
    self.obs.place(self._current, positions, quats)
    rows = np.arange(self.E)
    self._rest[rows, self._n_placed] = positions
    self._placed_at[rows, self._n_placed] = place_positions
    self._quats[rows, self._n_placed] = quats
    self._placed_quats[rows, self._n_placed] = place_quats
    self._n_placed += 1
    more = self._cursor < self._length
    self._current = np.where(more, self._order[rows, np.minimum(self._cursor, self._length - 1)],
                             self._current)
    self._cursor += more
    self._done = ~more
    self.obs.observe_walls()
    self.obs.observe_rocks(self._current)
    reward = self._reward()
    terminal = torch.from_numpy(self._done.copy()).to(self.dev)
    return self.observation, reward, terminal

  # -- rewards (rewarder.py:144-179, 261-307) ------------------------------------------ #
  def reward_terms(self):
    """(intersection, union, goal volume) per environment, device tensors."""
    return capi.reward_sums(self.obs.walls, self.goals, self._goal_z_d)

  def _reward(self):
    if self.metric in ('iou', 'or'):
      inter, uni, vol = self.reward_terms()
      value = inter / uni if self.metric == 'iou' else inter / vol
      value = value.double().cpu().numpy()
    else:
      # DOR / DIoU (rewarder.py:261-295): per placed rock, inside-goal test on its
      # rest position and the distance discount from where it was placed.
      g = self.obs.geo
      live = np.arange(self._length)[None, :] < self._n_placed[:, None]
      u = self._rest[..., 0] // g.pixel_h
      v = self._rest[..., 1] // g.pixel_w
      lims = self.goal_lims
      inside = live & (u >= lims[:, 0, 0, None]) & (v >= lims[:, 0, 1, None]) & \
        (u < lims[:, 1, 0, None]) & (v < lims[:, 1, 1, None])
      perr = np.linalg.norm(self._placed_at - self._rest, axis=-1)
      disc = np.ones_like(perr)
      if self._pexp is not None:
        disc = disc * np.maximum(0., 1 - (perr / self._pmax) ** self._pexp)
      if self._oexp is not None:
        # rotation distance 2*acos(min(w, 1)) of the difference quaternion
        # (simulator.py:116-117); w = <q_placed, q_rest> for unit quaternions
        w = np.minimum((self._placed_quats * self._quats).sum(axis=-1), 1.)
        oerr = 2 * np.arccos(np.where(live, w, 1.))
        disc = disc * np.maximum(0., 1 - (oerr / np.pi) ** self._oexp)
      total = np.where(inside, disc, 0.).sum(axis=1)
      n_out = (live & ~inside).sum(axis=1)
      value = total / self._length if self.metric == 'dor' else total / (self._length + n_out)
    out = (value - self._memory) * self.scale
    self._memory = value
    return torch.from_numpy(out.astype('float32')).to(self.dev)


class HeightPolicy(object):
  """Device-side ``Baseline('height', batched=True, batchwise=True)`` for a
  BatchedStackEnv: returns the action tensor(s) ``step`` expects."""

  def __init__(self, goal=True, minorder=1, threshold=0.75):
    # float32 maps come straight from the rasteriser: multiples of 2^-14 m
    self._scorer = PlacementScorer('height', goal, minorder, threshold,
                                   quantum_log2=HEIGHT_QUANTUM_LOG2)

  def __call__(self, env):
    walls, goals, rocks = env.planes()
    if env._dtype == 'uint8':
      # The reference policy scores the observation it is given: for the
      # registered uint8 envs that is the quantised maps (float64 arithmetic).
      wall_goal, rock = env.observation
      if env.R > 1:
        wall_goal = wall_goal[:, 0]
      walls = wall_goal[..., 0].contiguous()
      goals = wall_goal[..., 1].contiguous()
      rocks = rock[..., 0].contiguous()
      if env.R == 1:
        rocks = rocks[:, None].contiguous()
    # a goal rectangle is never empty, so goal.max() is the goal height
    level = env._goal_z_d if walls.dtype == torch.float32 else None
    best = self._scorer(walls, goals, rocks, level=level)['best']
    return (best[:, 0], best[:, 1]) if env.R > 1 else best[:, 1]
