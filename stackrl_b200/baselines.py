"""Host-side mirror of ``stackrl.baselines`` backed by the sm_100a kernels.

Same names, arguments and return conventions as the reference module
(/root/reference/stackrl/baselines.py) so that callers written against it --
``Baseline(method=...)``, ``stackrl.test.run``'s ``policy(o) -> (a, v)``, the
heat-map sweep -- keep working; the arithmetic runs on the GPU through the C
ABI (stackrl_b200.capi).  Two surfaces:

* drop-in functions on ONE reference-layout observation
  ``(wall_goal [H,W,2], rock [h,w,1])`` given as numpy arrays (host) -- they
  upload, run the kernel and return numpy arrays of the reference's dtype;
* ``PlacementScorer`` on device-resident batches (planar ``walls [E,H,W]``,
  ``goals [E,H,W]``, ``rocks [E,R,h,h]`` torch CUDA tensors), the form the
  batched environment and bench.py use.

Nothing here computes on the CPU; without a CUDA device every call raises.
"""
import numpy as np
import torch

from stackrl_b200 import capi


def _device():
  if not torch.cuda.is_available():
    raise RuntimeError('stackrl_b200 needs a CUDA device (no CPU fallback)')
  return torch.device('cuda', torch.cuda.current_device())


def _split(inputs):
  """Reference-layout observation -> (wall, goal, rock) numpy planes.
  Accepts one observation or the batched TestStackEnv layout."""
  wall_goal = np.asarray(inputs[0])
  rock = np.asarray(inputs[1])
  if wall_goal.shape[-1] != 2 or rock.shape[-1] != 1:
    raise ValueError('expected ([.., H, W, 2], [.., h, w, 1]) observations')
  if rock.shape[-2] != rock.shape[-3]:
    # baselines.py:39 slices both axes with rock.shape[0].
    raise ValueError('the reference only defines square rock maps')
  return wall_goal[..., 0], wall_goal[..., 1], rock[..., 0]


def _upload(x, dtype=None):
  t = torch.from_numpy(np.ascontiguousarray(x)).to(_device())
  return t if dtype is None else t.to(dtype)


def _planes(inputs):
  """One observation -> device planes walls [1,H,W], goals [1,H,W],
  rocks [1,1,h,h] in the observation's own dtype."""
  wall, goal, rock = _split(inputs)
  if wall.ndim != 2:
    raise ValueError('this function takes ONE observation; use PlacementScorer '
                     'or Baseline(batched=True) for batches')
  return _upload(wall[None]), _upload(goal[None]), _upload(rock[None, None])


def _height_device(walls, goals, rocks, quantum_log2=None, level=None):
  """[E,R,Ph,Pw] drop map in the reference's arithmetic for the obs dtype.
  ``quantum_log2``: see capi.maxplus_f32 (a hint, never changes the result).
  ``level``: goal.max() per environment when the caller already has it."""
  if walls.dtype == torch.float32:
    if level is None:
      level = capi.goal_level(goals)        # get_inputs: goal.max() (baselines.py:23)
    return capi.maxplus_f32(walls, rocks, level, quantum_log2=quantum_log2)
  if walls.dtype == torch.uint8:
    # uint8/uint8 is float64 in numpy: IEEE float64 a/g + b/g per cell.
    return capi.maxplus_u8(walls, rocks, capi.goal_level(goals) if level is None else level)
  raise TypeError(
    'observations must be float32 or uint8 (the dtypes the reference registers), '
    'got {}'.format(walls.dtype))


# ---- baselines.py:28-43 ------------------------------------------------------ #
def height(inputs, mask=None, **kwargs):
  """Height based heuristic (max-plus drop map), reference ``height``.

  Returns a float64 [H-h+1, W-w+1] array like the reference's np.zeros
  container.  float32 observations use the float32 kernel (bit-exact with
  numpy's float32 add/max), integer observations the float64 one."""
  walls, goals, rocks = _planes(inputs)
  f = _height_device(walls, goals, rocks)[0, 0].cpu().numpy().astype('float64')
  if mask is not None:
    f = np.where(mask, f, 0.)
  return f


# ---- baselines.py:45-77 ------------------------------------------------------ #
def difference(inputs, mask=None, difference_exponent=2, weights_exponent=2,
               return_height=False, **kwargs):
  """Difference based heuristic: weighted residual between the rock's underside
  and the wall at the drop height (float32 or uint8 observations, the latter in
  float64 throughout like numpy).  float64, bit-exact with the reference for
  ``difference_exponent`` in {1, 2} (numpy's float32 pow for other exponents is
  libm-defined and is refused rather than approximated)."""
  if difference_exponent not in (1, 2):
    raise ValueError('difference_exponent must be 1 or 2 for a bit-reproducible '
                     'result, got {}'.format(difference_exponent))
  walls, goals, rocks = _planes(inputs)
  if walls.dtype not in (torch.float32, torch.uint8):
    raise TypeError('observations must be float32 or uint8, got {}'.format(walls.dtype))
  u8 = walls.dtype == torch.uint8
  level = capi.goal_level(goals)
  if weights_exponent in (0, 2):
    weights = capi.difference_weights(rocks, None if u8 else level, weights_exponent)
  else:
    # Any other exponent goes through numpy's own float64 pow on the host so the
    # weights keep the reference's bits (baselines.py:54-62).
    n = rocks[0, 0].cpu().numpy() if u8 else (rocks[0, 0] / level[0]).cpu().numpy()
    live = n > 0
    di = (np.arange(n.shape[0], dtype='float') - n.shape[0] / 2) ** 2
    dj = (np.arange(n.shape[1], dtype='float') - n.shape[1] / 2) ** 2
    w = np.where(live, (di[:, None] + dj[None, :]) ** (weights_exponent / 2), 0)
    w /= w.sum()
    weights = _upload(w[None, None])
  run = capi.difference_u8 if u8 else capi.difference_f32
  f, top = run(walls, rocks, level, weights, difference_exponent, want_top=return_height)
  f = f[0, 0].cpu().numpy()
  if mask is not None:
    f = np.where(mask, f, 0.)
  if return_height:
    h0 = top[0, 0].cpu().numpy().astype('float64')
    if mask is not None:
      h0 = np.where(mask, h0, 0.)
    return f, h0
  return f


# ---- baselines.py:141-143 ---------------------------------------------------- #
def correlate(inputs, **kwargs):
  """Cross-correlation heuristic: correlate2d(o, n, 'valid') / n.sum(), float32.
  Matches scipy to ~1e-6 relative (library summation order)."""
  walls, goals, rocks = _planes(inputs)
  if walls.dtype != torch.float32:
    raise TypeError('correlate is wired for float32 observations')
  corr, _ = capi.correlate_f32(walls, rocks, capi.goal_level(goals), want_coef=False)
  return corr[0, 0].cpu().numpy()


# ---- baselines.py:79-114 ----------------------------------------------------- #
def corrcoef(inputs, mask=None, localized=False, **kwargs):
  """Correlation-coefficient heuristic.  Default: the reference's OpenCV
  TM_CCOEFF_NORMED path (baselines.py:84-85), float32, matched to ~1e-5
  absolute.  ``localized=True``: the masked variant (baselines.py:87-114),
  float64 container, bit-exact (numpy's pairwise sums in the observation's
  arithmetic type; float32 and uint8 observations)."""
  walls, goals, rocks = _planes(inputs)
  if localized:
    f = capi.corrcoef_localized(walls, rocks, capi.goal_level(goals))[0, 0].cpu().numpy()
    if mask is not None:
      f = np.where(mask, f, 0.)
    return f
  if walls.dtype != torch.float32:
    raise TypeError('corrcoef is wired for float32 observations')
  _, coef = capi.correlate_f32(walls, rocks, capi.goal_level(goals), want_corr=False)
  return coef[0, 0].cpu().numpy()


# ---- baselines.py:145-150 ---------------------------------------------------- #
def random(inputs, seed=None, **kwargs):
  """Random values in the shape of the heuristics (host RNG, like the reference:
  numpy's default_rng stream is the contract, there is nothing to accelerate)."""
  rng = np.random.default_rng(seed)
  return rng.random((np.subtract(np.shape(inputs[0]), np.shape(inputs[1]))[:-1] + 1))


# ---- baselines.py:152-156 ---------------------------------------------------- #
def goal_overlap(inputs, threshold=0.75, **kwargs):
  """Boolean mask of positions whose rock footprint overlaps the unfilled goal
  by at least ``threshold`` of the best overlap."""
  walls, goals, rocks = _planes(inputs)
  counts = capi.goal_overlap(walls, goals, rocks)[0, 0].cpu().numpy().astype('int64')
  return counts >= threshold * counts.max()


methods = {
  'random': random,
  'correlate': correlate,
  'height': height,
  'difference': difference,
  'corrcoef': corrcoef,
}


class PlacementScorer(object):
  """Device-resident batched scoring: the form the reference reaches by
  looping ``Baseline`` over environments and views (policies.py:63-73).

  All tensors stay on the GPU; nothing synchronises.  ``walls``/``goals``
  [E,H,W], ``rocks`` [E,R,h,h] (float32, planar)."""

  def __init__(self, method='height', goal=True, minorder=1, threshold=0.75,
               quantum_log2=None, difference_exponent=2, weights_exponent=2):
    """``method``: 'height' (max-plus drop map) or 'difference' (baselines.py:45-77,
    float64 maps).  ``quantum_log2``: the heightmaps are expected to be multiples of
    2**quantum_log2 (maps straight from the rasteriser: camera.HEIGHT_QUANTUM_LOG2);
    such environments are swept in exact 16-bit fixed point.  Same results."""
    if method not in ('height', 'difference'):
      raise ValueError("PlacementScorer scores with 'height' or 'difference'")
    if method == 'difference' and (difference_exponent not in (1, 2) or
                                   weights_exponent not in (0, 2)):
      raise ValueError('difference on device batches: exponents (1|2, 0|2) only')
    self.method = method
    self.difference_exponent = difference_exponent
    self.weights_exponent = weights_exponent
    self.goal = goal
    self.minorder = minorder
    self.threshold = threshold
    self.quantum_log2 = quantum_log2

  def values(self, walls, goals, rocks, level=None):
    if self.method == 'difference':
      if level is None:
        level = capi.goal_level(goals)
      u8 = walls.dtype == torch.uint8
      weights = capi.difference_weights(rocks, None if u8 else level, self.weights_exponent)
      run = capi.difference_u8 if u8 else capi.difference_f32
      return run(walls, rocks, level, weights, self.difference_exponent)[0]
    return _height_device(walls, goals, rocks, self.quantum_log2, level)

  def __call__(self, walls, goals, rocks, want_shown=False, fused='mask', level=None):
    """-> dict(values [E,R,Ph,Pw], actions [E,R], best [E,2] = (view, flat index),
    shown [E,R,Ph,Pw] float64 if requested, counts [E,R,Ph,Pw] when computed).

    fused='mask' (default): score map kernel + ONE kernel for goal mask, arg-min
    and batch-wise pick (overlap counts stay in shared memory);
    fused='full': everything in one launch (srl_score_f32; float32 batches of
    supported shapes, falls back otherwise); fused=False: the three separate
    kernels (also returns the counts).  ``level`` [E]: goal.max() per environment if
    the caller has it (float32 planes only; saves one reduction kernel)."""
    if fused == 'full' and not want_shown and walls.dtype == torch.float32 and \
        self.method == 'height':
      try:
        values, actions, best = capi.score_f32(
          walls, goals if self.goal else None, rocks,
          None if self.goal else capi.goal_level(goals),
          level_mode=2 if self.goal else 1, minorder=self.minorder or 0,
          overlap_threshold=self.threshold)
        return {'values': values, 'counts': None, 'actions': actions, 'best': best,
                'shown': None}
      except capi.SrlError as err:
        if err.code != capi.SRL_E_UNSUPPORTED:
          raise
    values = self.values(walls, goals, rocks, level)
    counts = None
    if self.goal and fused:
      try:
        actions, shown, best = capi.mask_select(
          values, walls, goals, rocks, minorder=self.minorder or 0,
          overlap_threshold=self.threshold, want_shown=want_shown)
        return {'values': values, 'counts': None, 'actions': actions, 'best': best,
                'shown': shown}
      except capi.SrlError as err:
        if err.code != capi.SRL_E_UNSUPPORTED:
          raise
    if self.goal:
      counts = capi.goal_overlap(walls, goals, rocks)
    actions, shown, best = capi.select(
      values, counts, minorder=self.minorder or 0, overlap_threshold=self.threshold,
      want_shown=want_shown, want_best=True)
    return {'values': values, 'counts': counts, 'actions': actions, 'best': best,
            'shown': shown}


class HostPipeline(object):
  """``PlacementScorer`` for batches that live in HOST memory.

  The batch is cut into chunks of environments; chunk k+1's host->device copy
  (from a pinned staging slab, on a copy stream) overlaps chunk k's kernels, and
  only the actions / batch-wise picks travel back.  This is the end-to-end call
  bench.py times (``e2e``): numpy in, numpy out.

  ``goal_rects=True``: the goal of an environment is handed over as what it is in
  the reference -- a rectangle ``Rewarder._goal_lims`` at height ``goal_z``
  (rewarder.py:252-258) -- i.e. ``goals`` is an int32 [E,4] array (u0, v0, u1, v1)
  plus ``levels`` [E]; the [E,H,W] goal planes the kernels read are filled on the
  device (srl_fill_goals_f32) instead of crossing PCIe (a quarter of the bytes of a
  float32 step)."""

  def __init__(self, scorer, envs, rotations, H, W, h, chunks=4, device=None,
               dtype=torch.float32, goal_rects=False):
    """``dtype``: observation dtype, float32 or uint8 (the dtype of the registered
    Stack-v0/1/2 environments, env.py:171-178: a quarter of the bytes per step)."""
    if dtype not in (torch.float32, torch.uint8):
      raise TypeError('observations must be float32 or uint8, got {}'.format(dtype))
    if goal_rects and dtype != torch.float32:
      raise TypeError('goal rectangles are filled as float32 planes')
    self.scorer = scorer
    self.dev = device if device is not None else _device()
    self.E, self.R = int(envs), int(rotations)
    self.rects = bool(goal_rects)
    self.bounds = [(k * self.E // chunks, (k + 1) * self.E // chunks) for k in range(chunks)]
    self.bounds = [b for b in self.bounds if b[1] > b[0]]
    # One pinned slab and one device slab per chunk, the chunk's arrays back to back
    # (256-byte aligned parts): ONE host->device copy per chunk.
    def parts(n):
      out = [('walls', (n, H, W), dtype), ('rocks', (n, self.R, h, h), dtype)]
      if self.rects:
        out += [('rects', (n, 4), torch.int32), ('levels', (n,), torch.float32)]
      else:
        out += [('goals', (n, H, W), dtype)]
      return out
    self.pin_slabs, self.dev_slabs, self.pin, self.dev_in = [], [], [], []
    self.goal_planes = []
    self.h2d_bytes = 0
    for lo, hi in self.bounds:
      offsets, total = [], 0
      for name, shape, dt in parts(hi - lo):
        nbytes = torch.empty((), dtype=dt).element_size() * int(np.prod(shape))
        offsets.append((name, shape, dt, total, nbytes))
        total += (nbytes + 255) // 256 * 256
        self.h2d_bytes += nbytes
      pin = torch.empty((total,), dtype=torch.uint8).pin_memory()
      dev = torch.empty((total,), dtype=torch.uint8, device=self.dev)
      view = lambda slab: {name: slab[off:off + nbytes].view(dt).view(shape)
                           for name, shape, dt, off, nbytes in offsets}
      self.pin_slabs.append(pin)
      self.dev_slabs.append(dev)
      self.pin.append(view(pin))
      self.dev_in.append(view(dev))
      if self.rects:
        self.goal_planes.append(torch.empty((hi - lo, H, W), dtype=torch.float32,
                                            device=self.dev))
    self.actions = torch.empty((self.E, self.R), dtype=torch.int64).pin_memory()
    self.best = torch.empty((self.E, 2), dtype=torch.int64).pin_memory()
    self.copy_stream = torch.cuda.Stream(device=self.dev)
    self.back_stream = torch.cuda.Stream(device=self.dev)
    self.graph = None
    self.d2h_bytes = (self.actions.numel() + self.best.numel()) * 8

  def stage(self, walls, goals, rocks, levels=None):
    """Copy caller arrays into the pinned staging slabs (host memcpy; bench.py's
    e2e timing starts after this, with the inputs in pinned host memory).
    With ``goal_rects`` ``goals`` is the [E,4] limits array and ``levels`` [E]."""
    for (lo, hi), pin in zip(self.bounds, self.pin):
      pin['walls'].numpy()[...] = walls[lo:hi]
      pin['rocks'].numpy()[...] = rocks[lo:hi]
      if self.rects:
        pin['rects'].numpy()[...] = goals[lo:hi]
        pin['levels'].numpy()[...] = levels[lo:hi]
      else:
        pin['goals'].numpy()[...] = goals[lo:hi]

  def _enqueue(self, main):
    """One step's copies and kernels on ``main`` (+ the copy stream, forked from
    and joined back into ``main``)."""
    self.copy_stream.wait_stream(main)
    ready = []
    with torch.cuda.stream(self.copy_stream):
      for pin, dev in zip(self.pin_slabs, self.dev_slabs):
        dev.copy_(pin, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        ready.append(ev)
    self._keep = []          # chunk outputs stay alive until their device->host copies ran
    for k, ((lo, hi), ev, dev_in) in enumerate(zip(self.bounds, ready, self.dev_in)):
      main.wait_event(ev)
      if self.rects:
        goals = capi.fill_goals(dev_in['rects'], dev_in['levels'], self.goal_planes[k])
        out = self.scorer(dev_in['walls'], goals, dev_in['rocks'], level=dev_in['levels'])
      else:
        out = self.scorer(dev_in['walls'], dev_in['goals'], dev_in['rocks'])
      # results travel back on their own stream: a device->host copy queued on `main`
      # would hold the next chunk's kernels behind the copy engine's round trip
      done = torch.cuda.Event()
      done.record(main)
      self.back_stream.wait_event(done)
      with torch.cuda.stream(self.back_stream):
        self.actions[lo:hi].copy_(out['actions'], non_blocking=True)
        self.best[lo:hi].copy_(out['best'], non_blocking=True)
      self._keep.append(out)
    main.wait_stream(self.back_stream)

  def capture(self):
    """Record one step (every copy and kernel of ``run``) into a CUDA graph; later
    ``run`` calls replay it with ONE launch instead of ~10 per chunk, which is what
    bounds small or uint8 batches (the staging buffers are fixed, so the graph's
    addresses stay valid).  Call after at least one eager ``run`` (kernel attributes
    and lazy module loads must not happen inside a capture).  Returns self."""
    stream = torch.cuda.Stream(device=self.dev)
    stream.wait_stream(torch.cuda.current_stream(self.dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
      with torch.cuda.graph(graph, stream=stream):
        self._enqueue(stream)
    torch.cuda.current_stream(self.dev).wait_stream(stream)
    self.graph = graph
    return self

  def run(self):
    """Score the staged batch: H2D per chunk -> max-plus, goal overlap, select
    -> D2H of actions.  Returns (actions [E,R], best [E,2]) numpy views after a
    full synchronise."""
    main = torch.cuda.current_stream(self.dev)
    if self.graph is not None:
      self.graph.replay()
    else:
      self._enqueue(main)
    main.synchronize()
    return self.actions.numpy(), self.best.numpy()

  def __call__(self, walls, goals, rocks, levels=None):
    self.stage(walls, goals, rocks, levels)
    return self.run()


class Baseline(object):
  """Greedy policy over a heuristic value map: the reference's
  ``stackrl.baselines.Baseline`` (baselines.py:167-217) with ``PyGreedy``'s
  return conventions (agents/policies.py:57-91).

  ``method`` is a name from ``methods`` or any callable
  ``method(inputs, **kwargs) -> [H-h+1, W-w+1]`` (lower is better), exactly as
  in the reference.  For the built-in 'height' method the whole call -- score
  map, goal mask, local-minimum arg-min, batch-wise pick -- runs on the GPU;
  for a user callable the map comes from the callable and the selection still
  runs on the GPU."""

  def __init__(self, method='random', goal=True, minorder=1, value=False,
               unravel=False, batched=False, batchwise=False, **kwargs):
    if isinstance(method, str):
      if method in methods:
        method = methods[method]
      else:
        raise ValueError(
          'Invalid value {} for argument method. Must be in {}'.format(method, methods))
    elif not callable(method):
      raise TypeError('Invalid type {} for argument method.'.format(type(method)))
    self.model = method
    self.goal = goal
    self.kwargs = kwargs
    self.minorder = minorder
    self.value = value
    self.unravel = unravel
    self.batched = batched
    self.batchwise = batchwise

  # -- device pipeline on N views: returns (actions [N], shown [N,Ph,Pw]) ------ #
  def _score(self, views):
    wall, goal, rock = views
    N = wall.shape[0]
    # Every view is its own "environment" with one rock: the reference handles
    # views independently (policies.py:63-64), including their goal maps.
    walls, goals, rocks = _upload(wall), _upload(goal), _upload(rock[:, None])
    if self.model is height:
      values = _height_device(walls, goals, rocks)
    elif self.model is difference and walls.dtype == torch.float32 and \
        self.kwargs.get('difference_exponent', 2) in (1, 2) and \
        self.kwargs.get('weights_exponent', 2) in (0, 2):
      level = capi.goal_level(goals)
      weights = capi.difference_weights(rocks, level,
                                        self.kwargs.get('weights_exponent', 2))
      values, _ = capi.difference_f32(walls, rocks, level, weights,
                                      self.kwargs.get('difference_exponent', 2))
    else:
      maps = [np.asarray(self.model((np.stack([wall[k], goal[k]], -1), rock[k][..., None]),
                                    **self.kwargs), dtype='float64') for k in range(N)]
      values = _upload(np.stack(maps)[:, None])
    threshold = self.kwargs.get('threshold', 0.75)
    if self.goal:
      try:
        actions, shown, _ = capi.mask_select(values, walls, goals, rocks,
                                             minorder=self.minorder or 0,
                                             overlap_threshold=threshold, want_best=False)
        return actions[:, 0].cpu().numpy(), shown[:, 0].cpu().numpy()
      except capi.SrlError as err:
        if err.code != capi.SRL_E_UNSUPPORTED:
          raise
    counts = capi.goal_overlap(walls, goals, rocks) if self.goal else None
    actions, shown, _ = capi.select(values, counts, minorder=self.minorder or 0,
                                    overlap_threshold=threshold, want_best=False)
    return actions[:, 0].cpu().numpy(), shown[:, 0].cpu().numpy()

  def call(self, inputs):
    """One observation -> (argmax index, value map), Baseline.call."""
    wall, goal, rock = _split(inputs)
    a, v = self._score((wall[None], goal[None], rock[None]))
    return int(a[0]), v[0]

  def __call__(self, inputs):
    if self.batched:
      wall, goal, rock = _split(inputs)
      actions, values = self._score((wall, goal, rock))
      if self.unravel:
        outputs = np.array([np.unravel_index(a, values.shape[1:]) for a in actions])
        picked = values[np.arange(len(actions)), outputs[:, 0], outputs[:, 1]]
      else:
        values = values.reshape(len(actions), -1)
        outputs = np.array(actions)
        picked = values[np.arange(len(actions)), actions]
      if self.batchwise:
        k = np.argmax(picked)
        outputs = k, outputs[k]
    else:
      outputs, values = self.call(inputs)
      if self.unravel:
        outputs = np.array(np.unravel_index(outputs, values.shape))
      elif self.value:
        values = values.ravel()
    if self.value:
      outputs = outputs, values
    return outputs
