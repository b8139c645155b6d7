#!/usr/bin/env python
"""bench.py -- placement evaluations / s of the placement-scoring hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): 4096 environments per GPU,
32x32 wall heightmap, 16x16 rock underside map, 8 rotations -> 4096*8*17*17 =
9 469 952 placement evaluations per step per GPU.  One step is one pass of the
scoring hot path over one batch of synthetic observations, i.e. what the
reference's ``Baseline('height', batched, batchwise)`` does per observation:
max-plus drop map (baselines.py:28-43), goal-overlap mask (:152-156), masked
local-minimum arg-min and batch-wise pick (:201-217, policies.py:57-91).
Environments are independent, so N GPUs run N shards with no data-path
collective ("weak" scaling); the ranks exchange timings and checksums with one
all-gather at the end.

Prints ONE JSON line (contract: task statement / DESIGN.md "Measurement").
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'placement_evals_per_s'
UNIT = 'placement evals/s'
CFG = dict(envs=4096, rotations=8, H=32, W=32, h=16)
NSETS = 8          # distinct input/output sets cycled through (beats the 126 MB L2)


def workload_name():
  return ('C2: batched max-plus placement search, {envs} envs/GPU, {H}x{W} wall, '
          '{h}x{h} rock, {rotations} rotations (score map + goal mask + arg-min)'
          ).format(**CFG)


def evals_per_step():
  P = (CFG['H'] - CFG['h'] + 1) * (CFG['W'] - CFG['h'] + 1)
  return CFG['envs'] * CFG['rotations'] * P


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return json.load(f), 'measured (MEASURED_PEAKS.json)'
  return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


def ncu_traffic(kernel):
  """dram bytes per launch of `kernel` from the committed ncu summary, if any."""
  path = os.path.join(ROOT, 'profiles', 'ncu_summary.json')
  if os.path.exists(path):
    with open(path) as f:
      return json.load(f).get(kernel, {}).get('dram_bytes_per_launch')
  return None


# --------------------------------------------------------------------------- #
# clocks: sampled with NVML during the timed region
# --------------------------------------------------------------------------- #
class ClockSampler(object):
  REASONS = {
    0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap',
    0x8: 'hw_slowdown', 0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown',
    0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown',
    0x100: 'display_clock_setting',
  }

  def __init__(self, index, period=0.004):
    self.samples, self.reasons = [], set()
    self.max_mhz = None
    self._stop = threading.Event()
    self._thread = None
    self._period = period
    try:
      import pynvml
      pynvml.nvmlInit()
      self._nv = pynvml
      self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
    except Exception:   # pragma: no cover - NVML missing
      self._nv = None

  def _sample(self):
    nv = self._nv
    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
    try:
      mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
    except Exception:
      mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
    for bit, name in self.REASONS.items():
      if mask & bit and name != 'gpu_idle':
        self.reasons.add(name)

  def _run(self):
    while not self._stop.is_set():
      self._sample()
      time.sleep(self._period)

  def start(self):
    if self._nv is None:
      return
    self._stop.clear()
    self._thread = threading.Thread(target=self._run, daemon=True)
    self._thread.start()

  def stop(self):
    if self._thread is not None:
      self._stop.set()
      self._thread.join()
      self._thread = None

  def summary(self, how):
    if not self.samples:
      return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [],
              'samples': 0, 'how': 'nvml unavailable'}
    return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
            'reasons': sorted(self.reasons), 'samples': len(self.samples),
            'how': how}


# --------------------------------------------------------------------------- #
# CPU baseline: the oracle's loop-form restatement of Baseline('height').call
# --------------------------------------------------------------------------- #
def _cpu_maps(args):
  """Worker: score `count` (env, rotation) views with the reference-shaped
  Python-loop + numpy path (oracle.scoring_np.baseline_call_loop)."""
  seed, count = args
  from oracle import scoring_np
  from stackrl_b200 import synth
  walls, rocks, _ = synth.placement_batch(seed, count, 1, CFG['H'], CFG['W'], CFG['h'])
  goals = synth.goals(seed + 7, count, CFG['H'], CFG['W'])
  obs = [(np.stack([walls[e], goals[e]], -1), rocks[e, 0][..., None]) for e in range(count)]
  t0 = time.perf_counter()
  for o in obs:
    scoring_np.baseline_call_loop(o)
  return time.perf_counter() - t0


CPU_SAMPLE = ('oracle.scoring_np.baseline_call_loop: numpy port of Baseline("height").call '
              '(baselines.py:28-43 Python double loop, :152-156, :201-217)')


def cpu_baseline_one_core(maps):
  P = evals_per_step() // (CFG['envs'] * CFG['rotations'])
  _cpu_maps((1, 4))
  dt = _cpu_maps((123, maps))
  return maps * P / dt, dt


def run_reference(args, rank, world):
  """--impl reference: the reference's CPU algorithm (oracle port; the Python
  reference itself cannot travel to the GPU box) on all host cores."""
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  maps_per_core = 40          # ~0.15 s of work per core per step at ~3.5 ms/view
  P = evals_per_step() // (CFG['envs'] * CFG['rotations'])
  import multiprocessing as mp
  ctx = mp.get_context('spawn')
  with ctx.Pool(cores) as pool:
    for _ in range(max(1, args.warmup)):
      pool.map(_cpu_maps, [(1, 2)] * cores)
    t0 = time.perf_counter()
    for s in range(args.steps):
      pool.map(_cpu_maps, [(1000 * s + k, maps_per_core) for k in range(cores)])
    wall = time.perf_counter() - t0
  value = args.steps * cores * maps_per_core * P / wall
  sample = ('{} views ({}x{} wall, {}x{} rock) per step over {} processes; {}'
            ).format(cores * maps_per_core, CFG['H'], CFG['W'], CFG['h'], CFG['h'], cores,
                     CPU_SAMPLE)
  line = {
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
    'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
    'ms_per_step': 1e3 * wall / args.steps, 'higher_is_better': True,
    'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': {'workload': workload_name(), 'sample': sample},
    'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                     'sample': sample},
    'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
            'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  print(json.dumps(line))


# --------------------------------------------------------------------------- #
# GPU arm
# --------------------------------------------------------------------------- #
def _time_loop(torch, fn, steps):
  ev0 = torch.cuda.Event(enable_timing=True)
  ev1 = torch.cuda.Event(enable_timing=True)
  for k in range(3):                  # warm-up: lazy module loads, allocator
    fn(k)
  torch.cuda.synchronize()
  ev0.record()
  for k in range(steps):
    fn(k)
  ev1.record()
  torch.cuda.synchronize()
  return ev0.elapsed_time(ev1) / steps


def extra_metrics(torch, dev, cpu_baseline=True):
  """Short measurements of the other two BASELINE metrics on this GPU: mesh
  rasterisation (config 3 geometry) and full env observations (config 4
  geometry, static settle).  Reported under `extra`; not the headline."""
  from stackrl_b200 import capi, envs, meshes
  from stackrl_b200.camera import ObserverGeometry
  out = {}
  # -- config 3: 4096 synthetic rocks, 32x32 px at 0.005 m/px ------------------- #
  n = 4096
  verts, tris = meshes.synthetic_rocks(4, n, max_dimension=0.16, frequency=10)   # ~2k tris
  bank = meshes.MeshBank()
  for k in range(n):
    bank.add(verts[k], tris)
  from stackrl_b200.observer import BatchedObserver
  obs = BatchedObserver(bank, n, 1, overhead_resolution=128, object_resolution=32,
                        pixel_size=0.005, max_z=0.375, device=dev)
  ids = np.arange(n)
  obs.observe_rocks(ids)
  torch.cuda.synchronize()
  g = obs.geo
  def rocks_only(_):
    capi.raster(obs._verts, obs._tris, obs._rock_inst, obs._rock_jobs, g.object_h,
                g.object_w, capi.RASTER_ROCK,
                out=obs.rocks.view(n, g.object_h, g.object_w))
  ms = _time_loop(torch, rocks_only, 20)
  ntri, nvert = len(tris), verts.shape[1]
  bytes_per_rock = 12 * nvert + 12 * ntri + 4 * 32 * 32
  peaks, _ = measured_peaks()
  out['raster'] = {
    'workload': 'C3: {} synthetic rocks x {} tris ({} verts), 32x32 px at 0.005 m/px'.format(
      n, ntri, nvert),
    'rocks_per_s': n / (ms * 1e-3), 'tris_per_s': n * ntri / (ms * 1e-3), 'ms': ms,
    'roofline': {'bound': 'hbm', 'achieved': n * bytes_per_rock / (ms * 1e-3) / 1e9,
                 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                 'frac': n * bytes_per_rock / (ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                 'bytes_per_rock': bytes_per_rock}}
  if cpu_baseline:
    # CPU baseline of the raster: the oracle's C z-buffer (gcc -O2, one core) on a
    # sample of the same rocks and camera -- the RESTATEMENT, not pybullet's
    # TinyRenderer (absent here).
    from oracle import raster_np
    spawn = ((0., 0., 0.375 + 0.16), (0., 0., 0., 1.))
    view = g.object_view(spawn, 0)
    sample = n                       # every rock of the batch once (~1 s)
    t0 = time.perf_counter()
    for k in range(sample):
      raster_np.render_depth(view, g.object_projection, g.object_h, g.object_w,
                             [(verts[k], tris, np.identity(3), np.array(spawn[0]))])
    dt = time.perf_counter() - t0
    out['raster']['cpu_baseline'] = {
      'value': sample / dt, 'unit': 'rocks/s', 'cores': 1, 'kind': 'port',
      'sample': '{} of the same rocks, depth image only, oracle/csrc/oracle.c z-buffer '
                'through ctypes in {:.2f} s'.format(sample, dt)}
  # -- config 2 on heightmaps as the rasteriser leaves them (multiples of 2^-14 m):
  #    the exact 16-bit fixed-point sweep behind srl_maxplus_f32_q ---------------- #
  from stackrl_b200 import baselines, synth
  E, R, H, W, h = (CFG[k] for k in ('envs', 'rotations', 'H', 'W', 'h'))
  walls_h, rocks_h, _ = synth.placement_batch(0, E, R, H, W, h)
  q = np.float32(2.0 ** -14)
  wq = torch.from_numpy((np.round(walls_h / q) * q).astype('float32')).to(dev)
  rq = torch.from_numpy((np.round(rocks_h / q) * q).astype('float32')).to(dev)
  gq = torch.from_numpy(synth.goals(7, E, H, W)).to(dev)
  lq = gq.amax(dim=(1, 2))
  vq = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32, device=dev)
  same = torch.equal(capi.maxplus_f32(wq, rq, lq), capi.maxplus_f32(wq, rq, lq, quantum_log2=-14))
  ms_q = _time_loop(torch, lambda _: capi.maxplus_f32(wq, rq, lq, out=vq, quantum_log2=-14), 50)
  ms_f = _time_loop(torch, lambda _: capi.maxplus_f32(wq, rq, lq, out=vq), 50)
  scorer_q = baselines.PlacementScorer('height', quantum_log2=-14)
  ms_s = _time_loop(torch, lambda _: scorer_q(wq, gq, rq), 50)
  evals = evals_per_step()
  out['quantised_heightmaps'] = {
    'workload': 'C2 shapes, walls/rocks rounded to multiples of 2^-14 m (what the '
                'float32 depth->elevation formulas of observer.py:259-260 produce), '
                'level 0.25; L2-warm single set',
    'maxplus_fixed_point_ms': ms_q, 'maxplus_fixed_point_evals_per_s': evals / (ms_q * 1e-3),
    'maxplus_float_ms': ms_f, 'bit_identical_to_float_sweep': bool(same),
    'scorer_ms': ms_s, 'scorer_evals_per_s': evals / (ms_s * 1e-3)}
  # -- config 2 shapes as uint8 observations (the dtype of the registered Stack-v0/1/2
  #    environments, env.py:171-178): device-resident and end to end from host memory -- #
  w8 = torch.from_numpy(synth.to_dtype(walls_h, 'uint8')).to(dev)
  r8 = torch.from_numpy(synth.to_dtype(rocks_h, 'uint8')).to(dev)
  g8 = torch.from_numpy(synth.to_dtype(synth.goals(7, E, H, W), 'uint8')).to(dev)
  scorer8 = baselines.PlacementScorer('height')
  ms_8 = _time_loop(torch, lambda _: scorer8(w8, g8, r8), 30)
  pipe8 = baselines.HostPipeline(scorer8, E, R, H, W, h, chunks=4, device=dev, dtype=torch.uint8)
  pipe8.stage(w8.cpu().numpy(), g8.cpu().numpy(), r8.cpu().numpy())
  ms_8e = _time_loop(torch, lambda _: pipe8.run(), 20)
  same8 = bool(np.array_equal(pipe8.run()[0], scorer8(w8, g8, r8)['actions'].cpu().numpy()))
  out['uint8_observations'] = {
    'workload': 'C2 shapes cast like StackEnv._return (uint8, goal level 170): float64 '
                'max-plus values through the integer-key sweep + goal mask + arg-min',
    'scorer_ms': ms_8, 'scorer_evals_per_s': evals / (ms_8 * 1e-3),
    'e2e_ms': ms_8e, 'e2e_evals_per_s': evals / (ms_8e * 1e-3),
    'h2d_bytes_per_step': pipe8.h2d_bytes, 'd2h_bytes_per_step': pipe8.d2h_bytes,
    'host_pipeline_matches_device': same8}
  del w8, r8, g8, pipe8
  # -- the DQN's Siamese correlation layer (nets/layers.py:21-38) -- #
  Bs, Cs, Hs, hs = 148, 16, 128, 32        # config.gin:55 geometry; 148 samples = whole waves
  gen = torch.Generator(device=dev).manual_seed(0)
  xs = torch.randn((Bs, Hs, Hs, Cs), device=dev, generator=gen)
  fs = torch.randn((Bs, hs, hs, Cs), device=dev, generator=gen)
  os_ = torch.empty((Bs, Hs - hs + 1, Hs - hs + 1, 1), device=dev)
  ms_c = _time_loop(torch, lambda _: capi.siam_correlation_f32(xs, fs, out=os_), 10)
  flops = 2.0 * Bs * (Hs - hs + 1) ** 2 * hs * hs * Cs
  fma_peak = 2 * max(capi.microbench_fma(v, 400) for v in (0, 1, 2)) / 1e12
  out['siam_correlation'] = {
    'workload': '{} samples, {}x{}x{} wall features * {}x{}x{} rock features, float32 '
                '(stackrl.nets.correlation, config.gin geometry)'.format(Bs, Hs, Hs, Cs, hs, hs, Cs),
    'ms': ms_c, 'samples_per_s': Bs / (ms_c * 1e-3),
    'roofline': {'bound': 'fp32-fma', 'achieved': flops / (ms_c * 1e-3) / 1e12,
                 'peak': fma_peak, 'unit': 'TFLOP/s',
                 'frac': flops / (ms_c * 1e-3) / 1e12 / fma_peak,
                 'peak_source': 'srl_microbench_fma, best of FFMA / FFMA2 measured in this run '
                                '(nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4)'}}
  del xs, fs, os_
  # -- config 4 slice: env observations, 64x64 wall, 16x16 rock ----------------- #
  E = 4096
  bank2 = meshes.MeshBank()
  v2, t2 = meshes.synthetic_rocks(5, 64, 1, max_dimension=0.12)
  for k in range(64):
    bank2.add(v2[k], t2)
  env = envs.BatchedStackEnv(bank2, E, episode_length=12, observable_size_ratio=4,
                             resolution_factor=4, dtype='float32', rewarder='iou', seed=5,
                             device=dev)
  policy = envs.HeightPolicy()
  env.reset()
  for _ in range(4):
    env.step(policy(env))
  torch.cuda.synchronize()
  def observe(_):
    env.obs.observe_walls()
    env.obs.observe_rocks(env._current)
    env.reward_terms()
    env.observation
  ms = _time_loop(torch, observe, 10)
  t0 = time.perf_counter()
  for _ in range(4):
    env.step(policy(env))
  torch.cuda.synchronize()
  step_ms = (time.perf_counter() - t0) / 4 * 1e3
  out['env_obs'] = {
    'workload': 'C4 slice: {} envs/GPU, 64x64 wall, 16x16 rock, ~5 placed rocks of {} tris, '
                'float32 obs + IoU terms'.format(E, len(t2)),
    'obs_per_s': E / (ms * 1e-3), 'ms': ms,
    'full_step_ms_with_host_glue': step_ms, 'full_steps_per_s': E / (step_ms * 1e-3)}
  # -- last (a failed capture must not disturb anything above): the host pipeline
  #    replayed as ONE CUDA graph per step, float32 and uint8 ------------------------ #
  try:
    E, R, H, W, h = (CFG[k] for k in ('envs', 'rotations', 'H', 'W', 'h'))
    walls_h, rocks_h, _ = synth.placement_batch(0, E, R, H, W, h)
    goals_h = synth.goals(7, E, H, W)
    res = {}
    for name, dt in (('float32', torch.float32), ('uint8', torch.uint8)):
      obs = [synth.to_dtype(x, name) for x in (walls_h, goals_h, rocks_h)]
      pipe = baselines.HostPipeline(baselines.PlacementScorer('height'), E, R, H, W, h,
                                    chunks=4, device=dev, dtype=dt)
      pipe.stage(*obs)
      eager = pipe.run()[0].copy()
      ms_e = _time_loop(torch, lambda _: pipe.run(), 20)
      pipe.capture()
      ms_g = _time_loop(torch, lambda _: pipe.run(), 20)
      res[name] = {'eager_ms': ms_e, 'graph_ms': ms_g,
                   'graph_evals_per_s': evals_per_step() / (ms_g * 1e-3),
                   'same_actions': bool(np.array_equal(eager, pipe.run()[0]))}
      del pipe
    out['host_pipeline_cuda_graph'] = res
  except Exception as exc:
    out['host_pipeline_cuda_graph'] = {'error': repr(exc)}
  return out


def run_graft(args, rank, local_rank, world):
  import torch
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device: stackrl_b200 has no CPU path')
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)

  from stackrl_b200 import baselines, capi, sharding, synth
  dist = sharding.init('nccl', dev) if world > 1 else None

  E, R, H, W, h = (CFG[k] for k in ('envs', 'rotations', 'H', 'W', 'h'))
  P = (H - h + 1) * (W - h + 1)
  # Synthetic shard: rank r owns environments [r*E, (r+1)*E) of the
  # global batch; the NSETS sets differ by a roll of the environment axis.
  walls_h, rocks_h, _ = synth.placement_batch(1000 * rank, E, R, H, W, h)
  goals_h = synth.goals(1000 * rank + 7, E, H, W)
  sets = []
  for s in range(NSETS):
    w = torch.roll(torch.from_numpy(walls_h).to(dev), shifts=s, dims=0).contiguous()
    g = torch.from_numpy(goals_h).to(dev).clone()
    r = torch.roll(torch.from_numpy(rocks_h).to(dev), shifts=-s, dims=0).contiguous()
    sets.append(dict(
      walls=w, goals=g, rocks=r, level=g.amax(dim=(1, 2)),
      values=torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32, device=dev),
      counts=torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.int32, device=dev),
      actions=torch.empty((E, R), dtype=torch.int64, device=dev),
      best=torch.empty((E, 2), dtype=torch.int64, device=dev)))
  set_bytes = sum(t.numel() * t.element_size() for t in sets[0].values())
  lib, P_ = capi.lib, capi._P
  main_stream = torch.cuda.current_stream()
  stream = P_(main_stream.cuda_stream)
  side_stream = torch.cuda.Stream(device=dev) if args.overlap else None
  side = P_(side_stream.cuda_stream) if args.overlap else None
  done_events = [torch.cuda.Event() for _ in range(NSETS)]
  mp_events = []

  fused = args.fused
  masked = not args.separate and not fused

  def step(k, timed=False):
    s = sets[k % NSETS]
    if timed:
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
    if fused:
      # ONE launch: score maps (written in full), goal mask, arg-min, batch-wise pick.
      capi._check(lib.srl_score_f32(
        P_(s['walls'].data_ptr()), P_(s['goals'].data_ptr()), P_(s['rocks'].data_ptr()),
        P_(None), P_(s['values'].data_ptr()), P_(s['actions'].data_ptr()),
        P_(s['best'].data_ptr()), E, R, H, W, h, 2, 1, 0.75, stream))
      if timed:
        b.record()
        mp_events.append((a, b))
      return
    if args.overlap and k >= NSETS:
      main_stream.wait_event(done_events[k % NSETS])   # set k's values are free again
    capi._check(lib.srl_maxplus_f32(
      P_(s['walls'].data_ptr()), P_(s['rocks'].data_ptr()), P_(s['level'].data_ptr()),
      P_(s['values'].data_ptr()), E, R, H, W, h, 0.0, stream))
    if timed:
      b.record()
      mp_events.append((a, b))
    if masked and args.overlap:
      ev = torch.cuda.Event()
      ev.record(main_stream)
      side_stream.wait_event(ev)
      capi._check(lib.srl_mask_select_f32(
        P_(s['values'].data_ptr()), P_(s['walls'].data_ptr()), P_(s['goals'].data_ptr()),
        P_(s['rocks'].data_ptr()), P_(s['actions'].data_ptr()), P_(None),
        P_(s['best'].data_ptr()), E, R, H, W, h, 1, 0.75, side))
      done_events[k % NSETS].record(side_stream)
      return
    if masked:
      capi._check(lib.srl_mask_select_f32(
        P_(s['values'].data_ptr()), P_(s['walls'].data_ptr()), P_(s['goals'].data_ptr()),
        P_(s['rocks'].data_ptr()), P_(s['actions'].data_ptr()), P_(None),
        P_(s['best'].data_ptr()), E, R, H, W, h, 1, 0.75, stream))
      return
    capi._check(lib.srl_goal_overlap_f32(
      P_(s['walls'].data_ptr()), P_(s['goals'].data_ptr()), P_(s['rocks'].data_ptr()),
      P_(s['counts'].data_ptr()), E, R, H, W, h, stream))
    capi._check(lib.srl_select_f32(
      P_(s['values'].data_ptr()), P_(s['counts'].data_ptr()), P_(s['actions'].data_ptr()),
      P_(None), P_(s['best'].data_ptr()), E, R, H - h + 1, W - h + 1, 1, 0.75, stream))

  def barrier():
    if dist is not None:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- warm-up ------------------------------------------------------------------ #
  warmup = max(args.warmup, 3)
  for k in range(warmup):
    step(k)
  barrier()

  # ---- timed region (device-resident inputs) ------------------------------------ #
  sampler = ClockSampler(local_rank)
  ev0 = torch.cuda.Event(enable_timing=True)
  ev1 = torch.cuda.Event(enable_timing=True)
  barrier()
  sampler.start()
  ev0.record()
  for k in range(args.steps):
    step(k, timed=True)
  if side_stream is not None:
    main_stream.wait_stream(side_stream)
  ev1.record()
  torch.cuda.synchronize()
  sampler.stop()
  elapsed_ms = ev0.elapsed_time(ev1)
  maxplus_ms = sum(a.elapsed_time(b) for a, b in mp_events) / len(mp_events)
  clocks_how = 'nvml during the timed region'
  if len(sampler.samples) < 3:
    # Timed region shorter than the NVML sampling period: sample the same
    # kernel stream for ~0.25 s right after it (not part of any reported time).
    sampler.start()
    t_end = time.perf_counter() + 0.25
    k = 0
    while time.perf_counter() < t_end:
      for _ in range(50):
        step(k)
        k += 1
      torch.cuda.synchronize()
    sampler.stop()
    clocks_how = 'nvml over a 0.25 s repeat of the timed loop (region too short to sample)'
  barrier()

  # ---- end to end: host buffers in, host actions out ------------------------------ #
  scorer = baselines.PlacementScorer('height')
  pipe = baselines.HostPipeline(scorer, E, R, H, W, h, chunks=8, device=dev)
  pipe.stage(walls_h, goals_h, rocks_h)
  e2e_steps = max(3, min(args.steps, 30))
  for _ in range(3):
    pipe.run()
  barrier()
  ev0.record()
  for _ in range(e2e_steps):
    actions_h, best_h = pipe.run()
  ev1.record()
  torch.cuda.synchronize()
  e2e_ms = ev0.elapsed_time(ev1)
  # The pipelined host path must agree with the device-resident one.
  same = bool(np.array_equal(actions_h, sets[0]['actions'].cpu().numpy()))

  # ---- gather: max over ranks, checksums ------------------------------------------ #
  stats = sharding.gather_stats(
    [elapsed_ms, e2e_ms, maxplus_ms, sharding.checksum(sets[0]['actions']) % 2 ** 40,
     float(same)], dev)
  if rank != 0:
    if dist is not None:
      dist.destroy_process_group()
    return
  elapsed_ms = float(stats[:, 0].max())
  e2e_ms = float(stats[:, 1].max())
  maxplus_ms = float(stats[:, 2].max())
  ms_per_step = elapsed_ms / args.steps
  value = world * evals_per_step() / (ms_per_step * 1e-3)
  e2e_value = world * evals_per_step() / (e2e_ms / e2e_steps * 1e-3)

  # ---- roofline of the dominant kernel (max-plus) ---------------------------------- #
  peaks, peak_src = measured_peaks()
  cells = evals_per_step() * h * h                  # (add, max) cells per launch
  kernel_s = maxplus_ms * 1e-3
  micro = {v: capi.microbench_addmax(v, 400) for v in (0, 2, 7)}
  peak_cells = max(micro.values())
  alg_bytes = 4 * (E * H * W + E * R * h * h + E * R * P)
  roofline = {
    'bound': 'fp32-alu',
    'kernel': 'score_fused_kernel<17,16>' if fused else 'maxplus_stream_kernel<17,16>',
    'kernel_ms': maxplus_ms,
    'share_of_step': maxplus_ms / ms_per_step,
    'achieved': 2 * cells / kernel_s / 1e12,
    'peak': 2 * peak_cells / 1e12,
    'unit': 'Tops/s',
    'frac': (cells / kernel_s) / peak_cells,
    'traffic': ncu_traffic('score_fused_kernel' if fused else 'maxplus_stream_kernel'),
    'peak_source': 'srl_microbench_addmax, best (add,max) issue rate measured in this run '
                   '(FADD+FMNMX {:.3g}, FADD2+FMNMX3 {:.3g}, FADD2+VIMNMX3 {:.3g} cells/s); '
                   'MEASURED_PEAKS.json has no non-tensor FP32 figure'.format(
                     micro[0], micro[2], micro[7]),
    'ops_per_eval': 2 * h * h,
    'hbm': {'achieved': alg_bytes / kernel_s / 1e9, 'peak': peaks['hbm_gbs'],
            'unit': 'GB/s', 'frac': alg_bytes / kernel_s / 1e9 / peaks['hbm_gbs'],
            'bytes_per_eval': alg_bytes / evals_per_step(), 'peak_source': peak_src},
  }

  line = {
    'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
    'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms_per_step,
    'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
    'dtype': 'f32', 'data': 'synthetic',
    'config': {'workload': workload_name(), 'evals_per_step_per_gpu': evals_per_step(),
               'l2': '{} distinct input/output sets cycled, {:.0f} MB each ({:.0f} MB total '
                     '> 126 MB L2)'.format(NSETS, set_bytes / 1e6, NSETS * set_bytes / 1e6),
               'parallelism': 'env-sharded x{}, no data-path collective'.format(world),
               'shard_checksums': [int(c) for c in stats[:, 3].tolist()],
               'host_pipeline_matches_device': bool(stats[:, 4].min() == 1.0)},
    'clocks': sampler.summary(clocks_how),
    'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': pipe.h2d_bytes,
            'd2h_bytes_per_step': pipe.d2h_bytes, 'steps': e2e_steps,
            'api': 'stackrl_b200.baselines.HostPipeline(PlacementScorer): pinned host '
                   'observations -> actions'},
    'gpu_launches': (1 if fused else 2 if masked else 3) * args.steps,
    'kernels_per_step': ['score_fused_kernel'] if fused else
    ['maxplus_stream_kernel', 'mask_select_packed_kernel'] if masked else
    ['maxplus_stream_kernel', 'goal_overlap_kernel', 'select_kernel'],
    'roofline': roofline,
  }
  if world == 1 and not args.no_extra:
    try:
      line['extra'] = extra_metrics(torch, dev, cpu_baseline=not args.no_cpu_baseline)
    except Exception as exc:   # the headline must survive a failure of the extras
      line['extra'] = {'error': repr(exc)}
  if world == 1 and not args.no_cpu_baseline:
    maps = 8192
    v, wall = cpu_baseline_one_core(maps)
    line['cpu_baseline'] = {
      'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
      'sample': '{} views of the same shapes in {:.1f} s on 1 core; {}; host has {} '
                'cores'.format(maps, wall, CPU_SAMPLE, os.cpu_count())}
  print(json.dumps(line))
  if dist is not None:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=10)
  ap.add_argument('--impl', default='graft', choices=['graft', 'reference'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-extra', action='store_true')
  ap.add_argument('--fused', action='store_true',
                  help='run the single fully fused kernel (srl_score_f32)')
  ap.add_argument('--overlap', action='store_true',
                  help='goal mask / arg-min kernel of step k on a second stream, '
                       'concurrent with the max-plus kernel of step k+1')
  ap.add_argument('--separate', action='store_true',
                  help='run the three separate kernels instead of the fused one')
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if args.impl == 'reference':
    run_reference(args, rank, world)
  else:
    run_graft(args, rank, local_rank, world)


if __name__ == '__main__':
  main()
