"""Batched stacking environment with device-resident observations.

The reference runs one ``StackEnv`` (stackrl/envs/stack/env.py) per process and
batches N of them over ``multiprocessing.Pipe`` (``ParallelEnv``,
stackrl/envs/utils.py:302-576).  ``BatchedStackEnv`` keeps that contract --
``reset() / step(action) -> (observation, reward, terminal)`` with a leading
batch axis, ``batch_size``, ``observation_spec``, ``action_spec``, ``sample()``,
``seed()`` -- for E environments on one GPU.  A step is a chain of sm_100a
kernels with NO device->host round trip: placement pose (a4) -> instance append
and next rock (a15) -> wall and rock rasterisation (a2/a3) -> reward (a11/a12)
-> packing (a14); ``capture()`` records the chain (optionally with the policy in
front) as one CUDA graph.  The rigid-body settle step is NOT part of this
package (BASELINE north_star: it stays the reference's pybullet code) and is
plugged in as ``settle``: a host callable that receives the placement poses and
returns where the rocks came to rest -- the only point where a step
synchronises.  The default ``settle=None`` leaves every rock where it was placed,
which is also what the golden episodes (reference env on the static fake
backend) do.

Host-side randomness follows the reference: per environment one RandomState for
the rock order (env.py:105, 268-272) and a second one for the goal rectangle,
seeded with ``randint(2**32)`` of the first (env.py:166, 346; rewarder.py:
211-259); environment i is seeded ``seed + i`` (utils.py:433, 530-532).
``vector_rng=True`` replaces the 2E streams by one vectorised stream (same
distributions, not the reference's draw sequence) for very large batches.
"""
import collections

import numpy as np
import torch

from stackrl_b200 import capi
from stackrl_b200.baselines import PlacementScorer
from stackrl_b200.camera import HEIGHT_QUANTUM_LOG2
from stackrl_b200.episodes import EpisodeSampler
from stackrl_b200.observer import BatchedObserver

METRIC_NAMES = ('IoU', 'OR', 'DIoU', 'DOR')      # Rewarder.metrics (rewarder.py:7)

# What get_space_spec (utils.py:19-48) returns is a tf.TensorSpec; TensorFlow is not part of
# this package, so a spec is its two fields the callers read: ``.shape`` and ``.dtype``.
Spec = collections.namedtuple('Spec', 'shape dtype')


class BatchedStackEnv(object):
  metadata = {'dtypes': ['uint8', 'float32']}

  def __init__(self, bank, batch_size, episode_length=30, object_max_dimension=0.125,
               observable_size_ratio=4, resolution_factor=5, max_z=0.375,
               rewarder=None, goal_size_ratio=.25, reward_scale=1., reward_params=None,
               orientation_freedom=0, dtype='float32', settle=None, seed=None,
               device=None, vector_rng=False, rock_cache_bytes=1 << 30, block=True,
               persistent_observation=False):
    """Arguments follow StackEnv (env.py:28-50); ``bank`` is the MeshBank of
    candidate rocks (the reference's ``urdfs`` list), ``orientation_freedom``
    the TestStackEnv option (env.py:443-463), ``settle`` the physics hook:
    ``settle(mesh_ids [E], positions [E,3], quaternions [E,4]) -> (positions,
    quaternions)`` of the new rocks at rest, optionally followed by a third item,
    the rest poses [E, n_placed, 7] of ALL placed rocks including the new one
    (Simulator.positions re-reads every body, simulator.py:86-92).  ``rock_cache_bytes``:
    see BatchedObserver (the spawned rocks' images are fetched from a per-bank table
    rasterised once, when it fits).  ``block``: ParallelEnv's default for ``step`` /
    ``reset`` (utils.py:393-428): False makes them return a callable that waits for the
    kernels of the call and returns the time step.  ``persistent_observation``: ``step``
    returns the SAME observation tensors every time and rewrites only the wall rows the
    placed rock changed (a quarter of the bytes of a step's packed observation); a caller
    that keeps an observation across steps must copy it.  ``capture()`` always works that
    way (a CUDA graph replays into fixed buffers)."""
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
      spawn_pose=((0., 0., max_z + object_max_dimension), (0., 0., 0., 1.)), device=device,
      episode_length=self._length, rock_cache_bytes=rock_cache_bytes)
    self.dev = self.obs.dev
    g = self.obs.geo
    self.R = g.n_orientations
    self._settle = settle
    # -- rewards (rewarder.py:17-142) --------------------------------------------- #
    metric = 'iou' if rewarder is None else str(rewarder).lower()
    if metric not in capi.METRICS:
      raise ValueError('Invalid value {} for argument metric'.format(rewarder))
    self.metric = metric
    self.scale = float(reward_scale) if reward_scale is not None else float(episode_length)
    if reward_params is None:
      self._pexp = self._oexp = None
    elif np.isscalar(reward_params):
      if reward_params < 0:
        raise ValueError('Invalid value {} for argument params. Must be non negative.'.format(
          reward_params))
      self._pexp = self._oexp = reward_params
    else:
      self._pexp, self._oexp = (list(reward_params) * 2)[:2]
    self._pmax = max(g.object_h * g.pixel_h, g.object_w * g.pixel_w)
    self._goal_z = g.max_z
    self._goal_size_ratio = goal_size_ratio
    H, W = g.overhead_h, g.overhead_w
    E = self.E
    self.goals = torch.zeros((E, H, W), dtype=torch.float32, device=self.dev)
    self._goal_z_d = torch.full((E,), self._goal_z, dtype=torch.float32, device=self.dev)
    self._goal_lims = np.zeros((E, 2, 2), dtype='int64')
    self._rects_d = torch.zeros((E, 4), dtype=torch.int32, device=self.dev)
    self._zero_reward = torch.zeros(E, dtype=torch.float32, device=self.dev)
    self._Ph, self._Pw = H - g.object_h + 1, W - g.object_w + 1
    # uint8 observations: StackEnv._return's scale and the level the quantised goal
    # plane reaches (env.py:171-178): trunc(goal_z * 255 / scale) in float32.
    self._scale = max(self._max_z, self._omd)
    level8 = np.array(np.float32(self._goal_z) * np.float32(255) / np.float32(self._scale))
    self._level8_d = torch.full((E,), int(level8.astype('uint8')), dtype=torch.uint8,
                                device=self.dev)
    # host mirror of the (deterministic) episode cursors: no device read-back
    self._order_h = np.zeros((E, self._length), dtype='int32')
    self._host_stale = False      # device-side draws not yet mirrored on the host
    self._episode = 0
    self._cursor = np.zeros(E, dtype='int64')
    self._done = np.ones(E, dtype=bool)
    self._sampler = EpisodeSampler(E, len(bank), self._length, (H, W), (g.object_h, g.object_w),
                                   goal_size_ratio, vector=vector_rng)
    self._graph = None
    self._block = bool(block)
    self._obs_buf = self._obs_full = None
    if persistent_observation:
      self._make_persistent()
    self.seed(seed)

  # -- ParallelEnv-style metadata (utils.py:185-300) --------------------------------- #
  @property
  def batch_size(self):
    return self.E

  @property
  def multiprocessing(self):
    return False

  @property
  def observation_spec(self):
    """get_space_spec of the observation space (utils.py:24-39): at most the last three
    dimensions are kept, so the view axis of TestStackEnv is not part of the spec."""
    g = self.obs.geo
    return (Spec((g.overhead_h, g.overhead_w, 2), self._dtype),
            Spec((g.object_h, g.object_w, 1), self._dtype))

  @property
  def action_spec(self):
    n = self._Ph * self._Pw
    # Discrete(n) -> a scalar int64 per environment (TestStackEnv: Tuple(Discrete(R),
    # Discrete(n)), env.py:463); the bounds are not part of a TensorSpec
    return (Spec((), 'int64'), Spec((), 'int64')) if self.R > 1 else Spec((), 'int64')

  def seed(self, seed=None):
    """Per-environment streams seeded seed + i (utils.py:433, 530-532); the goal
    stream of each environment is seeded from its rock stream like
    StackEnv.seed -> Rewarder.seed (env.py:340-346, rewarder.py:196-200).  Returns what
    ParallelEnv.seed returns: one ``[seed_i, goal_seed_i]`` per environment
    (``vector_rng``: the single ``[[seed]]`` of the batch-wide stream)."""
    seed = self._sampler.seed(seed)
    self._seed = seed
    self._action_rng = np.random.RandomState(seed % 2 ** 32)
    if self._sampler.vector:
      return [[seed]]
    return [[(seed + i) % 2 ** 32, g] for i, g in enumerate(self._sampler.goal_seeds)]

  def sample(self):
    n = self._Ph * self._Pw
    flat = torch.from_numpy(self._action_rng.randint(n, size=self.E))
    if self.R > 1:
      return torch.from_numpy(self._action_rng.randint(self.R, size=self.E)), flat
    return flat

  # -- host mirrors of what may have been drawn on the device (vector_rng) --------------- #
  def _sync_host(self):
    if self._host_stale:
      self._order_h[...] = self.obs.state.order.cpu().numpy()
      self._goal_lims[...] = self._rects_d.cpu().numpy().reshape(self.E, 2, 2)
      self._host_stale = False

  @property
  def goal_lims(self):
    """Goal limits ((u, v), (u + h, v + w)) per environment, Rewarder._goal_lims."""
    self._sync_host()
    return self._goal_lims

  @property
  def _order(self):
    self._sync_host()
    return self._order_h

  def set_goals(self, lims, env_ids=None):
    """Install goal rectangles ((u, v), (u+h, v+w)) (rewarder.py:252-258): the
    limits go to the device, the maps are filled there."""
    ids = np.arange(self.E) if env_ids is None else np.asarray(list(env_ids), dtype='int64')
    self._mark_full(None if env_ids is None else ids)
    self._sync_host()
    self._goal_lims[ids] = np.asarray(lims, dtype='int64').reshape(len(ids), 2, 2)
    self._rects_d.copy_(torch.from_numpy(
      np.ascontiguousarray(self._goal_lims.reshape(self.E, 4), dtype='int32')), non_blocking=True)
    if env_ids is None:
      capi.fill_goals(self._rects_d, self._goal_z_d, self.goals)
    else:
      rects = torch.from_numpy(np.ascontiguousarray(
        self._goal_lims[ids].reshape(len(ids), 4), dtype='int32')).to(self.dev, non_blocking=True)
      ids_d = torch.from_numpy(ids.astype('int32')).to(self.dev, non_blocking=True)
      capi.fill_goals(rects, self._goal_z_d, self.goals, env_ids=ids_d)

  # -- episode control ----------------------------------------------------------------- #
  def _deliver(self, out, block):
    """ParallelEnv's ``block`` convention (utils.py:393-428): the time step itself, or a
    callable that waits for the kernels queued so far and returns it.  (A blocking
    call does not synchronise either: the tensors are ordered on the current stream.)"""
    if self._block if block is None else block:
      return out
    done = torch.cuda.Event()
    done.record(torch.cuda.current_stream(self.dev))
    def ready():
      done.synchronize()
      return out
    return ready

  def __call__(self, *args, **kwargs):
    """Calls step (utils.py:263-265)."""
    return self.step(*args, **kwargs)

  def close(self):
    """Nothing to release: no worker processes, no physics client (env.py:333-338)."""

  def reset(self, env_ids=None, rock_orders=None, goal_lims=None, block=None):
    """Start new episodes (env.py:266-293).  ``rock_orders`` / ``goal_lims``
    override the random draws (used to replay recorded episodes)."""
    ids = np.arange(self.E) if env_ids is None else np.asarray(list(env_ids), dtype='int64')
    n = len(ids)
    self._mark_full(None if env_ids is None else ids)
    self._cursor[ids] = 1
    self._done[ids] = False
    if self._sampler.vector and rock_orders is None and goal_lims is None:
      # one counter-based stream for the whole batch, drawn on the device: no host
      # loop, no upload (same distributions, not the reference's draw sequence)
      g = self.obs.geo
      ids_d = None if env_ids is None else torch.from_numpy(ids.astype('int32')).to(self.dev)
      capi.env_draw(self.obs.state, self._rects_d, len(self.bank),
                    (g.overhead_h, g.overhead_w), (g.object_h, g.object_w),
                    self._goal_size_ratio, self._seed, self._episode, ids_d)
      self._episode += 1
      self._host_stale = True
      capi.fill_goals(self._rects_d, self._goal_z_d, self.goals)     # idempotent for the rest
      self.obs.begin(None, None if env_ids is None else ids)
    else:
      # draw order of the reference: episode list first (env.py:268-272), then the
      # goal (rewarder.reset, env.py:283); an override skips that draw
      if rock_orders is not None:
        orders = np.asarray(rock_orders, dtype='int64').reshape(n, self._length)
      else:
        orders = self._sampler.orders(ids)
      if goal_lims is not None:
        lims = np.asarray(goal_lims, dtype='int64').reshape(n, 2, 2)
      else:
        lims = self._sampler.goals(ids)
      self._sync_host()
      self._order_h[ids] = orders[:, ::-1]     # the reference pops from the end (env.py:245)
      self.set_goals(lims, None if env_ids is None else ids)
      self.obs.begin(self._order_h, None if env_ids is None else ids)
    self._observe()
    return self._deliver(
      (self.observation, self._zero_reward, self.obs.state.done.view(torch.bool)), block)

  @property
  def _current(self):
    """Mesh id of the spawned rock of every environment (host mirror)."""
    rows = np.arange(self.E)
    return self._order[rows, np.maximum(self._cursor, 1) - 1].astype('int64')

  @property
  def _n_placed(self):
    return self.obs.state.n_placed.cpu().numpy().astype('int64')

  @property
  def _rest(self):
    """Rest positions [E, length, 3] of the placed rocks (device history)."""
    return self.obs.state.hist_rest.cpu().numpy()[..., :3]

  @property
  def _quats(self):
    return self.obs.state.hist_rest.cpu().numpy()[..., 3:]

  def _observe(self, appended=False):
    self.obs.observe_walls(appended)
    self.obs.observe_rocks()

  def _pack(self, out=None):
    wall_goal, rock = capi.pack_obs(self.obs.walls, self.goals, self.obs.rocks,
                                    dtype=self._dtype, scale=self._scale,
                                    repeat_wall=self.R > 1, out=out)
    return wall_goal, (rock if self.R > 1 else rock[:, 0])

  @property
  def observation(self):
    """Packed observation in the env dtype (env.py:226-231; the TestStackEnv
    layout of env.py:472-480 when orientation_freedom > 0)."""
    return self._pack()

  def planes(self):
    """Planar float32 maps for device-side scoring: (walls, goals, rocks)."""
    return self.obs.walls, self.goals, self.obs.rocks

  def planes_u8(self):
    """The same maps cast like the uint8 observation (env.py:171-178), planar."""
    self._planes8 = capi.quantise_planes(self.obs.walls, self.goals, self.obs.rocks,
                                         self._scale, out=getattr(self, '_planes8', None))
    return self._planes8

  # -- one step ------------------------------------------------------------------------- #
  def _as_action(self, action):
    if self.R > 1:
      views, flat = action
    else:
      views, flat = None, action
    def dev64(a):
      if a is None:
        return None
      if isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.int64:
        return a
      a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
      return torch.from_numpy(np.ascontiguousarray(a, dtype='int64')).to(self.dev)
    return dev64(views), dev64(flat)

  def _step_device(self, views, flat):
    """The kernel chain of one step (no host synchronisation when settle is None):
    -> (observation, reward)."""
    obs = self.obs
    obs.poses_device(views, flat)
    appended = True
    if self._settle is None:
      obs.advance()
    else:
      placed = obs.pose_buf.cpu().numpy()
      if bool(obs.status.any()):
        raise AssertionError('Invalid action.')
      res = self._settle(self._current.copy(), placed[:, :3].copy(), placed[:, 3:].copy())
      rest = np.concatenate([np.asarray(res[0], dtype='float64').reshape(self.E, 3),
                             np.asarray(res[1], dtype='float64').reshape(self.E, 4)], axis=1)
      obs.advance(torch.from_numpy(np.ascontiguousarray(rest)).to(self.dev), obs.pose_buf)
      if len(res) > 2 and res[2] is not None:
        obs.set_poses(res[2])            # earlier rocks moved: the whole scene is redrawn
        appended = False
    self._observe(appended)
    return self._reward_and_pack()

  def _make_persistent(self):
    """Fixed observation buffers + the per-environment "rewrite everything" flags."""
    g = self.obs.geo
    tdt = {'float32': torch.float32, 'uint8': torch.uint8}[self._dtype]
    lead = (self.E, self.R) if self.R > 1 else (self.E,)
    self._obs_buf = (
      torch.empty(lead + (g.overhead_h, g.overhead_w, 2), dtype=tdt, device=self.dev),
      torch.empty((self.E, self.R, g.object_h, g.object_w, 1), dtype=tdt, device=self.dev))
    self._obs_full = torch.ones((self.E,), dtype=torch.uint8, device=self.dev)

  def _mark_full(self, ids=None):
    """New goal / new episode: every row of the persistent observation is stale."""
    if self._obs_full is None:
      return
    if ids is None:
      self._obs_full.fill_(1)
    else:
      self._obs_full[torch.as_tensor(np.asarray(ids), device=self.dev, dtype=torch.long)] = 1

  def _reward_and_pack(self):
    """Packed observation + reward of the step in one launch (a14 + a11/a12)."""
    g = self.obs.geo
    persistent = {} if self._obs_buf is None else dict(
      out=self._obs_buf, rows=self.obs.wall_rows, full=self._obs_full)
    # self.goals is always fill_goals(self._rects_d, self._goal_z_d): the kernel takes the
    # rectangle instead of reading the map
    wall_goal, rock, r = capi.pack_rewards(
      self.obs.state, self.obs.walls, None, self.obs.rocks, self._goal_z_d, self._rects_d,
      self.metric, self.scale, (g.pixel_h, g.pixel_w), self._pmax, self._pexp, self._oexp,
      dtype=self._dtype, obs_scale=self._scale, repeat_wall=self.R > 1, **persistent)
    if self.metric == 'all':
      # the reference returns the four metrics as the info dict (env.py:258-262)
      r = {name: r[:, k] for k, name in enumerate(METRIC_NAMES)}
    return (wall_goal, rock if self.R > 1 else rock[:, 0]), r

  def step(self, action, block=None):
    """action: [E] flat indices, or (views [E], flat indices [E]) when
    orientation_freedom > 0 (env.py:233-264, 482-520).  Device int64 tensors are
    used in place; anything else is uploaded.  ``block``: see ``_deliver``."""
    if self._done.any():
      raise RuntimeError('reset() the finished environments before stepping them')
    views, flat = self._as_action(action)
    if self._graph is not None and self._graph_policy is None:
      if views is not None:
        self._g_views.copy_(views, non_blocking=True)
      self._g_flat.copy_(flat, non_blocking=True)
      self._graph.replay()
      reward, observation = self._g_reward, self._g_obs
    else:
      observation, reward = self._step_device(views, flat)
    self._advance_host()
    return self._deliver((observation, reward, self.obs.state.done.view(torch.bool)), block)

  def _advance_host(self):
    more = self._cursor < self._length
    self._cursor += more
    self._done = ~more

  def contact_precheck(self, action, eps=2. ** -13):
    """Heightmap contact pre-check of ``action`` BEFORE stepping (the
    map analogue of Simulator._drop's contact count, simulator.py:337-341):
    -> (contact cells [E] int32, octant mask [E] int32, supported [E] bool)."""
    views, flat = self._as_action(action)
    return capi.contact_precheck(self.obs.walls, self.obs.rocks, views, flat, eps=eps)

  def check_actions(self):
    """Synchronises and raises like env.py:237 if any action of the steps so far
    was outside the action space (such actions place nothing valid: NaN pose)."""
    if bool(self.obs.status.any()):
      raise AssertionError('Invalid action.')

  # -- CUDA graph ------------------------------------------------------------------------ #
  def capture(self, policy=None):
    """Record one step -- ``policy(self)`` first when given -- as ONE CUDA graph.
    Afterwards ``step(action)`` (or ``step_policy()``) replays it: one launch per
    step instead of 7-9.  Needs settle=None (a host hook cannot be captured) and
    at least one eager step before (lazy kernel attributes).  Returns self."""
    if self._settle is not None:
      raise RuntimeError('a settle hook runs on the host and cannot be captured')
    if self._obs_buf is None:
      self._make_persistent()
    self._mark_full()
    dev = self.dev
    self._g_views = torch.zeros(self.E, dtype=torch.int64, device=dev) if self.R > 1 else None
    self._g_flat = torch.zeros(self.E, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    stream.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
      with torch.cuda.graph(graph, stream=stream):
        if policy is not None:
          action = policy(self)
          views, flat = self._as_action(action)
          self._g_action = action
        else:
          views, flat = self._g_views, self._g_flat
        self._g_obs, self._g_reward = self._step_device(views, flat)
    torch.cuda.current_stream(dev).wait_stream(stream)
    self._graph, self._graph_policy = graph, policy
    return self

  def step_policy(self):
    """Replay the graph captured with a policy: action selection + step."""
    if self._graph is None or self._graph_policy is None:
      raise RuntimeError('capture(policy) first')
    if self._done.any():
      raise RuntimeError('reset() the finished environments before stepping them')
    self._graph.replay()
    self._advance_host()
    return self._g_obs, self._g_reward, self.obs.state.done.view(torch.bool)

  # -- rewards (rewarder.py:144-179, 261-307) ------------------------------------------ #
  def reward_terms(self):
    """(intersection, union, goal volume) per environment, device tensors."""
    return capi.reward_sums(self.obs.walls, self.goals, self._goal_z_d)

  def _reward(self):
    g = self.obs.geo
    r = capi.rewards(self.obs.state, self.obs.walls, self.goals, self._goal_z_d, self._rects_d,
                     self.metric, self.scale, (g.pixel_h, g.pixel_w), self._pmax, self._pexp,
                     self._oexp)
    if self.metric == 'all':
      # the reference returns the four metrics as the info dict (env.py:258-262)
      return {name: r[:, k] for k, name in enumerate(METRIC_NAMES)}
    return r


class HeightPolicy(object):
  """Device-side ``Baseline('height', batched=True, batchwise=True)`` for a
  BatchedStackEnv: returns the action tensor(s) ``step`` expects (column views of
  the selection kernel's ``best`` table, read in place by the pose kernel)."""

  def __init__(self, goal=True, minorder=1, threshold=0.75):
    # float32 maps come straight from the rasteriser: multiples of 2^-14 m
    self._scorer = PlacementScorer('height', goal, minorder, threshold,
                                   quantum_log2=HEIGHT_QUANTUM_LOG2)

  def __call__(self, env):
    if env._dtype == 'uint8':
      # The reference policy scores the observation it is given: for the
      # registered uint8 envs that is the quantised maps (float64 arithmetic).
      walls, goals, rocks = env.planes_u8()
      level = env._level8_d
    else:
      walls, goals, rocks = env.planes()
      level = env._goal_z_d          # a goal rectangle is never empty: goal.max() == goal_z
    best = self._scorer(walls, goals, rocks, level=level)['best']
    return (best[:, 0], best[:, 1]) if env.R > 1 else best[:, 1]
