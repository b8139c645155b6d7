#!/usr/bin/env python
"""bench.py -- the observation + placement-scoring hot path of stackrl on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]
                    [--workload c2|c4|c5]

Workloads (BASELINE.json configs):
  c2 (default, the headline): batched max-plus placement search, 4096 environments
     per GPU, 32x32 wall, 16x16 rock, 8 rotations -> 9 469 952 placement evaluations
     per step per GPU.  One step = what the reference's Baseline('height', batched,
     batchwise) does per observation: max-plus drop map (baselines.py:28-43),
     goal-overlap mask (:152-156), masked local-minimum arg-min and batch-wise pick
     (:201-217, policies.py:57-91).  Weak scaling: every GPU owns 4096 environments.
  c4: full obs + reward pipeline for a DQN rollout: 65 536 environments IN TOTAL
     (sharded E/N), 64x64 wall, 16x16 rock; one step = HeightPolicy + BatchedStackEnv
     .step (pose, instance append, wall + rock rasterisation, IoU reward, packing),
     episodes of 30 rocks with the resets inside the timed region -> env obs / s.
  c5: heat-map sweep: 7 282 walls x 36 rotations = 262 152 candidate maps IN TOTAL
     (sharded by walls), 128x128 wall, 32x32 rock, every float32 score map AND the
     float64 value map the heat-map consumes (test.py:221-224, 269-280) written.

Environments are independent, so N GPUs run N shards with no data-path collective;
the ranks exchange timings and checksums with one all-gather at the end.

``--impl reference`` times the reference's OWN code (oracle/refload.py executing the
unmodified files staged by oracle/make_ref.py) on the host cores for the same
workload.  Prints ONE JSON line (contract: task statement / DESIGN.md "Measurement").
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

NSETS = 8          # distinct input/output sets cycled through in c2 (beats the 126 MB L2)

WORKLOADS = {
  'c2': dict(metric='placement_evals_per_s', unit='placement evals/s', scaling='weak',
             envs=4096, rotations=8, H=32, W=32, h=16),
  'c4': dict(metric='env_obs_per_s', unit='env obs/s', scaling='strong',
             envs=65536, rotations=1, H=64, W=64, h=16, episode_length=30, bank=64),
  'c5': dict(metric='placement_evals_per_s', unit='placement evals/s', scaling='strong',
             envs=7282, rotations=36, H=128, W=128, h=32),
}
# CPU sample of c5: 4 of the 36 views of a wall (one view is ~0.1 s of reference code)
_CPU_SAMPLE = {'c2': dict(WORKLOADS['c2']), 'c5': dict(WORKLOADS['c5'], rotations=4)}


def config_of(name):
  """The `config` object of the JSON line: identical in both arms."""
  w = WORKLOADS[name]
  text = {
    'c2': 'C2: batched max-plus placement search, {envs} envs/GPU, {H}x{W} wall, {h}x{h} '
          'rock, {rotations} rotations (score map + goal mask + arg-min)',
    'c4': 'C4: full obs+reward pipeline, {envs} envs in total, {H}x{W} wall, {h}x{h} rock, '
          'episodes of {episode_length} rocks (height policy + env step: pose, wall/rock '
          'raster, IoU reward, packed float32 observation)',
    'c5': 'C5: heat-map sweep, {envs} walls x {rotations} rotations = 262152 candidate maps '
          'in total, {H}x{W} wall, {h}x{h} rock (score maps + float64 value maps + arg-min)',
  }[name].format(**w)
  cfg = {'workload': text}
  cfg.update({k: v for k, v in w.items() if k not in ('metric', 'unit', 'scaling')})
  return cfg


def positions(w):
  return (w['H'] - w['h'] + 1) * (w['W'] - w['h'] + 1)


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
# clocks: sampled with NVML while the timed kernels run
# --------------------------------------------------------------------------- #
class ClockSampler(object):
  REASONS = {
    0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap',
    0x8: 'hw_slowdown', 0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown',
    0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown',
    0x100: 'display_clock_setting',
  }

  def __init__(self, index, period=0.002):
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


def bind_cores(local_rank, world):
  """Give every rank its own slice of the host cores (the pinned staging slabs are
  first touched, hence placed, by the rank that fills them)."""
  try:
    cores = sorted(os.sched_getaffinity(0))
    if world > 1 and len(cores) >= world:
      per = len(cores) // world
      mine = cores[local_rank * per:(local_rank + 1) * per]
      os.sched_setaffinity(0, mine)
      return mine
  except (AttributeError, OSError):
    pass
  return None


# --------------------------------------------------------------------------- #
# CPU arms: the reference's own code through oracle/refload.py
# --------------------------------------------------------------------------- #
def _reference():
  """(namespace of reference modules, kind).  kind 'reference' when the unmodified
  files could be loaded (build container: /root/reference; GPU box: oracle/_ref),
  else the numpy port of the oracle ('port')."""
  from oracle import refload
  if refload.available():
    from oracle import fake_pybullet
    return refload.load(pybullet=fake_pybullet.FakeBullet()), 'reference'
  return None, 'port'


def _cpu_c2(args):
  """Worker: score `count` environments of `rot` views with the reference's
  Baseline('height', batched, batchwise) (PyGreedy loop over the views)."""
  name, seed, count = args
  w = _CPU_SAMPLE[name]
  from stackrl_b200 import synth
  rot = w['rotations']
  walls, rocks, _ = synth.placement_batch(seed, count, rot, w['H'], w['W'], w['h'])
  goals = synth.goals(seed + 7, count, w['H'], w['W'])
  ns, kind = _reference()
  if ns is not None:
    policy = ns.baselines.Baseline(method='height', batched=True, batchwise=True, value=True)
  else:
    from oracle import scoring_np
    policy = lambda o: scoring_np.greedy(
      o, lambda x: scoring_np.baseline_call_loop(x), batched=True, batchwise=True, value=True)
  obs = []
  for e in range(count):
    wg = np.stack([walls[e], goals[e]], -1)
    obs.append((np.stack([wg] * rot), rocks[e][..., None]))
  t0 = time.perf_counter()
  for o in obs:
    policy(o)
  return time.perf_counter() - t0, kind


def _cpu_c4(args):
  """Worker: `steps` steps of ONE reference StackEnv (env.py) + Baseline('height') on
  the static fake pybullet (its camera is the oracle's C z-buffer: the RESTATEMENT of
  TinyRenderer), 64x64 wall / 16x16 rock.  Returns (seconds, steps, kind)."""
  name, seed, steps = args
  w = WORKLOADS[name]
  ns, kind = _reference()
  if ns is None:
    return 0., 0, 'unavailable'
  import glob
  urdfs = sorted(glob.glob(os.path.join(ns.root, 'stackrl/envs/data/generated', '*.urdf')))
  env = ns.env.StackEnv(urdfs=urdfs, episode_length=w['episode_length'], resolution_factor=4,
                        observable_size_ratio=4, dtype='float32', rewarder='iou', seed=seed)
  policy = ns.baselines.Baseline(method='height')
  obs = env.reset()
  done_steps = 0
  t0 = time.perf_counter()
  while done_steps < steps:
    obs, reward, done, _ = env.step(policy(obs))
    done_steps += 1
    if done:
      obs = env.reset()
  dt = time.perf_counter() - t0
  env.close()
  return dt, done_steps, kind


def _pool_run(fn, jobs, cores):
  import multiprocessing as mp
  ctx = mp.get_context('spawn')
  with ctx.Pool(cores) as pool:
    return pool.map(fn, jobs)


def cpu_arm(name, steps, warmup, cores):
  """Times the reference on `cores` processes; returns (value, wall s/step, sample, kind)."""
  w = WORKLOADS[name]
  import multiprocessing as mp
  ctx = mp.get_context('spawn')
  P = positions(w)
  with ctx.Pool(cores) as pool:
    if name in ('c2', 'c5'):
      # environments (all their views) per core and step: ~0.6 s of reference code per
      # step and core, so the default two steps are ~20 core-seconds of CPU work
      per_core = 48 if name == 'c2' else 2
      views = _CPU_SAMPLE[name]['rotations']
      fn = _cpu_c2
      job = lambda s, k: (name, 1000 * s + k, per_core)
      for _ in range(max(1, warmup)):
        pool.map(fn, [job(0, k) for k in range(cores)])
      t0 = time.perf_counter()
      kinds = []
      for s in range(steps):
        kinds += [k for _, k in pool.map(fn, [job(s + 1, k) for k in range(cores)])]
      wall = time.perf_counter() - t0
      units = steps * cores * per_core * views * P
      sample = ('{} environments x {} views ({}x{} wall, {}x{} rock) per step over {} '
                'processes: stackrl.baselines.Baseline("height", batched, batchwise) of the '
                'unmodified reference'.format(cores * per_core, views, w['H'], w['W'], w['h'],
                                              w['h'], cores))
    else:
      per = 36                     # env steps per core and step (~0.6 s)
      for _ in range(max(1, warmup)):
        pool.map(_cpu_c4, [(name, k, 2) for k in range(cores)])
      t0 = time.perf_counter()
      kinds, units = [], 0
      for s in range(steps):
        res = pool.map(_cpu_c4, [(name, 100 * s + k, per) for k in range(cores)])
        units += sum(r[1] for r in res)
        kinds += [r[2] for r in res]
      wall = time.perf_counter() - t0
      sample = ('{} processes x {} env steps per step: the unmodified reference StackEnv + '
                'Observer + Rewarder + Baseline("height") on the static fake pybullet (depth '
                'images from the oracle z-buffer, the restatement of TinyRenderer; env '
                'construction and reset outside the count)'.format(cores, per))
  kind = 'reference' if kinds and all(k == 'reference' for k in kinds) else 'port'
  return units / wall, wall / steps, sample, kind


def run_reference(args, rank, world):
  """--impl reference: rank 0 times the reference's CPU implementation."""
  if rank != 0:
    return
  w = WORKLOADS[args.workload]
  cores = os.cpu_count() or 1
  value, s_per_step, sample, kind = cpu_arm(args.workload, args.steps, args.warmup, cores)
  line = {
    'impl': 'reference', 'metric': w['metric'], 'value': value, 'unit': w['unit'],
    'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
    'ms_per_step': 1e3 * s_per_step, 'higher_is_better': True,
    'scaling': w['scaling'], 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': config_of(args.workload),
    'cpu_baseline': {'value': value, 'unit': w['unit'], 'cores': cores, 'kind': kind,
                     'sample': sample},
    'e2e': {'value': value, 'unit': w['unit'], 'h2d_bytes_per_step': 0,
            'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  print(json.dumps(line))


# --------------------------------------------------------------------------- #
# GPU arm helpers
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


class Setup(object):
  """torch / device / process group of one rank."""

  def __init__(self, args, rank, local_rank, world):
    import torch
    if not torch.cuda.is_available():
      raise SystemExit('bench.py needs a CUDA device: stackrl_b200 has no CPU path')
    torch.cuda.set_device(local_rank)
    self.torch = torch
    self.dev = torch.device('cuda', local_rank)
    self.rank, self.local_rank, self.world = rank, local_rank, world
    self.cores = bind_cores(local_rank, world) if not args.no_bind else None
    from stackrl_b200 import sharding
    self.sharding = sharding
    self.dist = sharding.init('nccl', self.dev) if world > 1 else None

  def barrier(self):
    if self.dist is not None:
      self.dist.barrier()
    self.torch.cuda.synchronize()

  def finish(self):
    if self.dist is not None:
      self.dist.destroy_process_group()


def timed_region(su, step, steps, warmup, preroll_s=0.3):
  """W warm-up steps, an untimed pre-roll of the same steps (so that NVML samples the
  clocks of exactly these kernels even when K steps last a few milliseconds), then
  EXACTLY K steps between barrier + synchronize on both sides.  Returns (elapsed ms,
  clock sampler, how)."""
  torch = su.torch
  for k in range(warmup):
    step(k, False)
  su.barrier()
  sampler = ClockSampler(su.local_rank)
  sampler.start()
  t_end = time.perf_counter() + preroll_s
  k = warmup
  while time.perf_counter() < t_end:
    for _ in range(8):
      step(k, False)
      k += 1
    torch.cuda.synchronize()
  ev0 = torch.cuda.Event(enable_timing=True)
  ev1 = torch.cuda.Event(enable_timing=True)
  su.barrier()
  ev0.record()
  for k in range(steps):
    step(k, True)
  ev1.record()
  su.barrier()
  sampler.stop()
  how = ('nvml every ~2 ms over one continuous run of identical steps: {:.1f} s untimed '
         'pre-roll + the {} timed steps ({:.1f} ms), no idle gap'.format(
           preroll_s, steps, ev0.elapsed_time(ev1)))
  return ev0.elapsed_time(ev1), sampler, how


# --------------------------------------------------------------------------- #
# c2: batched max-plus placement search (the headline)
# --------------------------------------------------------------------------- #
def run_c2(args, su):
  torch, dev, rank, world = su.torch, su.dev, su.rank, su.world
  from stackrl_b200 import baselines, capi, synth
  w = WORKLOADS['c2']
  E, R, H, W, h = (w[k] for k in ('envs', 'rotations', 'H', 'W', 'h'))
  P = positions(w)
  evals = E * R * P
  # Synthetic shard: rank r owns environments [r*E, (r+1)*E) of the
  # global batch; the NSETS sets differ by a roll of the environment axis.
  walls_h, rocks_h, levels_h = synth.placement_batch(1000 * rank, E, R, H, W, h)
  rects_h = synth.goal_rects(1000 * rank + 7, E, H, W)
  goals_h = synth.goals(1000 * rank + 7, E, H, W)
  sets = []
  for s in range(NSETS):
    wl = torch.from_numpy(np.roll(walls_h, s, axis=0)).to(dev)
    g = torch.from_numpy(goals_h).to(dev)
    r = torch.from_numpy(np.roll(rocks_h, -s, axis=0)).to(dev)
    sets.append(dict(
      walls=wl, goals=g, rocks=r, level=capi.goal_level(g),
      values=torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32, device=dev),
      counts=torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.int32, device=dev),
      actions=torch.empty((E, R), dtype=torch.int64, device=dev),
      best=torch.empty((E, 2), dtype=torch.int64, device=dev)))
  set_bytes = sum(t.numel() * t.element_size() for t in sets[0].values())
  lib, P_ = capi.lib, capi._P
  main_stream = torch.cuda.current_stream()
  stream = P_(main_stream.cuda_stream)
  mp_events = []
  fused = args.fused
  masked = not args.separate and not fused

  def step(k, timed):
    s = sets[k % NSETS]
    if timed:
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
    if fused:
      capi._check(lib.srl_score_f32(
        P_(s['walls'].data_ptr()), P_(s['goals'].data_ptr()), P_(s['rocks'].data_ptr()),
        P_(None), P_(s['values'].data_ptr()), P_(s['actions'].data_ptr()),
        P_(s['best'].data_ptr()), E, R, H, W, h, 2, 1, 0.75, stream))
      if timed:
        b.record()
        mp_events.append((a, b))
      return
    capi._check(lib.srl_maxplus_f32(
      P_(s['walls'].data_ptr()), P_(s['rocks'].data_ptr()), P_(s['level'].data_ptr()),
      P_(s['values'].data_ptr()), E, R, H, W, h, 0.0, stream))
    if timed:
      b.record()
      mp_events.append((a, b))
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

  warmup = max(args.warmup, 3)
  elapsed_ms, sampler, clocks_how = timed_region(su, step, args.steps, warmup)
  maxplus_ms = sum(a.elapsed_time(b) for a, b in mp_events) / len(mp_events)

  # ---- end to end: pinned host buffers in, host actions out ------------------------ #
  def e2e(dtype_name, rects):
    dt = {'float32': torch.float32, 'uint8': torch.uint8}[dtype_name]
    scorer = baselines.PlacementScorer('height')
    # chunk counts from tools/exp_e2e.py: the float32 step is the H2D copy (more chunks = a
    # shorter un-overlapped tail), the uint8 step is kernel-bound (fewer, fuller launches)
    pipe = baselines.HostPipeline(scorer, E, R, H, W, h, chunks=6 if rects else 3, device=dev,
                                  dtype=dt, goal_rects=rects)
    if rects:
      pipe.stage(walls_h, rects_h, rocks_h, levels_h)
    else:
      pipe.stage(*[synth.to_dtype(x, dtype_name) for x in (walls_h, goals_h, rocks_h)])
    n = max(3, min(args.steps, 30))
    for _ in range(3):
      pipe.run()
    pipe.capture()
    for _ in range(2):
      pipe.run()
    su.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n):
      actions_h, best_h = pipe.run()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / n
    return dict(ms=ms, steps=n, h2d=pipe.h2d_bytes, d2h=pipe.d2h_bytes,
                actions=actions_h.copy())
  e_f32 = e2e('float32', True)
  e_u8 = e2e('uint8', False)
  same = bool(np.array_equal(e_f32['actions'], sets[0]['actions'].cpu().numpy()))

  stats = su.sharding.gather_stats(
    [elapsed_ms, e_f32['ms'], maxplus_ms, su.sharding.checksum(sets[0]['actions']) % 2 ** 40,
     float(same), e_u8['ms']], dev)
  if rank != 0:
    return None
  elapsed_ms = float(stats[:, 0].max())
  e2e_ms = float(stats[:, 1].max())
  e2e8_ms = float(stats[:, 5].max())
  maxplus_ms = float(stats[:, 2].max())
  ms_per_step = elapsed_ms / args.steps
  value = world * evals / (ms_per_step * 1e-3)

  # ---- roofline of the dominant kernel (max-plus) ---------------------------------- #
  peaks, peak_src = measured_peaks()
  cells = evals * h * h                  # (add, max) cells per launch
  kernel_s = maxplus_ms * 1e-3
  micro = {v: capi.microbench_addmax(v, 400) for v in (0, 2, 7)}
  peak_cells = max(micro.values())
  alg_bytes = 4 * (E * H * W + E * R * h * h + E * R * P)
  nominal_dual = 2 * 148 * 128 * 1.965e9 / 1e12      # FMA-pipe add + ALU-pipe max every clock
  roofline = {
    'bound': 'fp32-alu',
    'kernel': 'score_fused_kernel<17,16>' if fused else 'maxplus_stream_kernel<17,16>',
    'kernel_ms': maxplus_ms,
    'share_of_step': maxplus_ms / ms_per_step,
    'achieved': 2 * cells / kernel_s / 1e12,
    'peak': 2 * peak_cells / 1e12,
    'unit': 'Tops/s',
    'frac': (cells / kernel_s) / peak_cells,
    'frac_of_dual_issue_nominal': 2 * cells / kernel_s / 1e12 / nominal_dual,
    'dual_issue_nominal': nominal_dual,
    'traffic': ncu_traffic('score_fused_kernel' if fused else 'maxplus_stream_kernel'),
    'peak_source': 'srl_microbench_addmax, best (add,max) issue rate measured in this run '
                   '(FADD+FMNMX {:.3g}, FADD2+FMNMX3 {:.3g}, FADD2+VIMNMX3 {:.3g} cells/s); '
                   'MEASURED_PEAKS.json has no non-tensor FP32 figure; dual_issue_nominal = '
                   '148 SMs x 128 lanes x 2 pipes x 1.965 GHz if an add and a max issued '
                   'every clock'.format(micro[0], micro[2], micro[7]),
    'ops_per_eval': 2 * h * h,
    'hbm': {'achieved': alg_bytes / kernel_s / 1e9, 'peak': peaks['hbm_gbs'],
            'unit': 'GB/s', 'frac': alg_bytes / kernel_s / 1e9 / peaks['hbm_gbs'],
            'bytes_per_eval': alg_bytes / evals, 'peak_source': peak_src},
  }
  line = {
    'n_gpus': world, 'ms_per_step': ms_per_step, 'value': value,
    'clocks': sampler.summary(clocks_how),
    'e2e': {'value': world * evals / (e2e_ms * 1e-3), 'unit': w['unit'],
            'h2d_bytes_per_step': e_f32['h2d'], 'd2h_bytes_per_step': e_f32['d2h'],
            'steps': e_f32['steps'], 'ms_per_step': e2e_ms,
            'h2d_gbs_per_rank': e_f32['h2d'] / (e2e_ms * 1e-3) / 1e9,
            'api': 'stackrl_b200.baselines.HostPipeline(PlacementScorer, goal_rects=True): '
                   'pinned host float32 walls + rocks + goal rectangles -> host actions; 6 '
                   'chunks, one H2D copy each, results back on their own stream, replayed as '
                   'one CUDA graph; the host memcpy '
                   'into the pinned slabs (stage()) is before the timed region'},
    'e2e_uint8': {'value': world * evals / (e2e8_ms * 1e-3), 'unit': w['unit'],
                  'h2d_bytes_per_step': e_u8['h2d'], 'd2h_bytes_per_step': e_u8['d2h'],
                  'ms_per_step': e2e8_ms,
                  'h2d_gbs_per_rank': e_u8['h2d'] / (e2e8_ms * 1e-3) / 1e9,
                  'api': 'same pipeline (3 chunks) on uint8 observations, the dtype of the registered '
                         'Stack-v0/1/2 environments (float64 max-plus values, env.py:171-178)'},
    'gpu_launches': (1 if fused else 2 if masked else 3) * args.steps,
    'kernels_per_step': ['score_fused_kernel'] if fused else
    ['maxplus_stream_kernel', 'mask_select_packed_kernel'] if masked else
    ['maxplus_stream_kernel', 'goal_overlap_kernel', 'select_kernel'],
    'roofline': roofline,
    'notes': {
      'evals_per_step_per_gpu': evals,
      'l2': '{} distinct input/output sets cycled, {:.0f} MB each ({:.0f} MB total > 126 MB '
            'L2)'.format(NSETS, set_bytes / 1e6, NSETS * set_bytes / 1e6),
      'parallelism': 'env-sharded x{}, no data-path collective'.format(world),
      'shard_checksums': [int(c) for c in stats[:, 3].tolist()],
      'host_pipeline_matches_device': bool(stats[:, 4].min() == 1.0),
      'host_cores_of_rank0': su.cores},
  }
  return line


# --------------------------------------------------------------------------- #
# c4: env observations per second
# --------------------------------------------------------------------------- #
def run_c4(args, su):
  torch, dev, rank, world = su.torch, su.dev, su.rank, su.world
  from stackrl_b200 import envs, meshes
  w = WORKLOADS['c4']
  lo, hi = su.sharding.shard_range(args.envs or w['envs'], rank, world)
  E = hi - lo
  H, h, L = w['H'], w['h'], w['episode_length']
  bank = meshes.MeshBank()
  v, t = meshes.synthetic_rocks(5, w['bank'], 1, max_dimension=0.12)      # 80 triangles each
  for k in range(w['bank']):
    bank.add(v[k], t)

  def make(dtype):
    env = envs.BatchedStackEnv(bank, E, episode_length=L, observable_size_ratio=H // h,
                               resolution_factor=int(np.log2(h)), dtype=dtype, rewarder='iou',
                               seed=5 + 7919 * rank, device=dev, vector_rng=True)
    return env, envs.HeightPolicy()
  env, policy = make('float32')
  env.reset()
  state = {'left': L}
  # The public call of a rollout: BatchedStackEnv.capture(policy) records policy + step
  # as ONE CUDA graph and step_policy() replays it (--eager: the same chain launch by
  # launch).  Resets (every L steps, inside the timed region) are eager launches.
  for _ in range(2):
    env.step(policy(env))
    state['left'] -= 1
  if not args.eager:
    env.capture(policy)

  def step(k, timed):
    if state['left'] == 0:
      env.reset()
      state['left'] = L
    if args.eager:
      env.step(policy(env))
    else:
      env.step_policy()
    state['left'] -= 1
  warmup = max(args.warmup, 3)
  elapsed_ms, sampler, clocks_how = timed_region(su, step, args.steps, warmup)

  # ---- per-kernel times: one mid-episode step with CUDA events around every call ---- #
  while state['left'] != L // 2:
    step(0, False)
  marks = []
  def mark(name):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    marks.append((name, ev))
  obs = env.obs
  reps = 5
  breakdown = {}
  for rep in range(reps + 1):
    if rep == 1:
      marks = []                  # the first pass is the warm-up of this exact sequence
    mark('start')
    action = policy(env)
    mark('policy (max-plus + goal mask + arg-min)')
    obs.poses_device(None, action)
    obs.advance()
    mark('pose + instance append')
    obs.observe_walls(True)
    mark('wall raster (the appended rock onto the kept depth image)')
    obs.observe_rocks()
    mark('rock images (per-bank table, rasterised once)')
    env._reward_and_pack()
    mark('reward + pack')
    env._advance_host()
    state['left'] -= 1
  torch.cuda.synchronize()
  for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
    if n1 != 'start':
      breakdown[n1] = breakdown.get(n1, 0.) + e0.elapsed_time(e1) / reps
  n_inst = float(obs.counts.float().mean().item())

  # ---- end to end: host actions in, reward + terminal (+ observation) to the host ---- #
  act_pin = torch.empty(E, dtype=torch.int64).pin_memory()
  rew_pin = torch.empty(E, dtype=torch.float32).pin_memory()
  term_pin = torch.empty(E, dtype=torch.uint8).pin_memory()
  e2e_steps = max(3, min(args.steps, L - 2))
  env.reset()
  act_pin.copy_(policy(env))
  torch.cuda.synchronize()
  obs_pin = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in env.observation]
  def e2e_loop(with_obs, steps=None):
    timed = steps is None
    if timed:
      e2e_loop(with_obs, 3)       # untimed warm-up of this exact loop (allocator, pinned copies)
    env.reset()
    su.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(e2e_steps if timed else steps):
      action = act_pin.to(dev, non_blocking=True)           # the agent's action, from the host
      o, r, t_ = env.step(action)
      rew_pin.copy_(r, non_blocking=True)
      term_pin.copy_(t_.view(torch.uint8), non_blocking=True)
      if with_obs:
        obs_pin[0].copy_(o[0], non_blocking=True)
        obs_pin[1].copy_(o[1], non_blocking=True)
      act_pin.copy_(policy(env), non_blocking=True)         # next action back to the host
      torch.cuda.current_stream().synchronize()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / e2e_steps
  e2e_ms = e2e_loop(False)
  e2e_obs_ms = e2e_loop(True)
  obs_bytes = sum(x.numel() * x.element_size() for x in obs_pin)

  # ---- the same step on uint8 observations (the registered environments' dtype) ---- #
  # (N = 1 only: the policy then scores the quantised maps in float64, like the reference's)
  u8_line = None
  if world == 1:
    del obs_pin
    env8, policy8 = make('uint8')
    env8.reset()
    for _ in range(2):
      env8.step(policy8(env8))
    if not args.eager:
      env8.capture(policy8)
    run8 = (lambda: env8.step(policy8(env8))) if args.eager else env8.step_policy
    for _ in range(3):
      run8()
    n8 = min(20, L - 8)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(n8):
      run8()
    ev1.record()
    torch.cuda.synchronize()
    ms8 = ev0.elapsed_time(ev1) / n8
    u8_line = {'value': E / (ms8 * 1e-3), 'unit': w['unit'], 'ms_per_step': ms8, 'steps': n8,
               'what': 'the same step with dtype="uint8": uint8 packed observation, the height '
                       'policy on the quantised maps (float64 max-plus values, env.py:171-178 / '
                       'baselines.py:21-26), steps 6..{} of an episode'.format(5 + n8)}
    del env8, policy8

  stats = su.sharding.gather_stats([elapsed_ms, e2e_ms, e2e_obs_ms, float(E), n_inst], dev)
  if rank != 0:
    return None
  total_E = int(stats[:, 3].sum())
  elapsed_ms = float(stats[:, 0].max())
  ms_per_step = elapsed_ms / args.steps
  value = total_E / (ms_per_step * 1e-3)
  peaks, peak_src = measured_peaks()
  V, F = v.shape[1], len(t)
  # what one observation has to move at least: the packed observation out, the wall map in
  # for the reward sums, the appended rock's mesh, the spawned rock's image in (per-bank
  # table) and out (planar map for the scorer); the pixels the rock changes in the kept
  # wall image are not counted
  alg = {'packed_observation_write': 4 * (2 * H * H + h * h),
         'reward_read': 4 * H * H,
         'rock_image_read_write': 8 * h * h,
         'mesh_read': 12 * V + 12 * F}
  alg_bytes = sum(alg.values())
  step_s = sum(breakdown.values()) * 1e-3
  obs_s = (step_s - breakdown['policy (max-plus + goal mask + arg-min)'] * 1e-3)
  line = {
    'n_gpus': world, 'ms_per_step': ms_per_step, 'value': value,
    'clocks': sampler.summary(clocks_how),
    'e2e': {'value': total_E / (float(stats[:, 1].max()) * 1e-3), 'unit': w['unit'],
            'h2d_bytes_per_step': 8 * E, 'd2h_bytes_per_step': 5 * E + 8 * E,
            'steps': e2e_steps,
            'api': 'BatchedStackEnv.step(action from pinned host memory) -> reward and '
                   'terminal read back to the host, next action of HeightPolicy to the host; '
                   'the observation stays on the GPU for its consumer (the DQN replica)'},
    'e2e_observation_to_host': {
      'value': total_E / (float(stats[:, 2].max()) * 1e-3), 'unit': w['unit'],
      'd2h_bytes_per_step': obs_bytes + 13 * E,
      'api': 'same, plus the packed float32 observation copied to pinned host memory'},
    'uint8_observations': u8_line,
    'gpu_launches': 7 * args.steps,
    'kernels_per_step': ['maxplus_stream_kernel', 'mask_select_packed_kernel',
                         'place_poses_kernel', 'env_advance_kernel',
                         'raster_warp_kernel (walls)',
                         'gather_rows_kernel (rock images)', 'pack_rewards_kernel'],
    'roofline': {
      'bound': 'hbm', 'kernel': 'env observation chain (pose, append, wall raster, rock '
                                'raster, reward, pack) of one mid-episode step',
      'kernel_ms': obs_s * 1e3, 'share_of_step': obs_s / step_s,
      'achieved': E * alg_bytes / obs_s / 1e9, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
      'frac': E * alg_bytes / obs_s / 1e9 / peaks['hbm_gbs'], 'traffic': None,
      'bytes_per_obs': alg, 'peak_source': peak_src,
      'breakdown_ms': breakdown, 'mean_placed_rocks': n_inst},
    'notes': {'envs_per_gpu': E, 'resets_in_timed_region': args.steps // L,
              'launch': 'eager' if args.eager else 'one CUDA graph per step (capture(policy) + '
                                                   'step_policy())',
              'rng': 'vector_rng=True (episode draws on the device, srl_env_draw)',
              'rocks': '{} synthetic rocks of {} triangles / {} vertices'.format(w['bank'], F, V)},
  }
  return line


# --------------------------------------------------------------------------- #
# c5: heat-map sweep
# --------------------------------------------------------------------------- #
def run_c5(args, su):
  torch, dev, rank, world = su.torch, su.dev, su.rank, su.world
  from stackrl_b200 import baselines, capi, synth
  w = WORKLOADS['c5']
  R, H, W, h = (w[k] for k in ('rotations', 'H', 'W', 'h'))
  lo, hi = su.sharding.shard_range(args.envs or w['envs'], rank, world)
  E = hi - lo
  P = positions(w)
  Ph = H - h + 1
  evals = E * R * P
  block = 256                                    # walls generated per host block
  walls = torch.empty((E, H, W), dtype=torch.float32, device=dev)
  goals = torch.empty((E, H, W), dtype=torch.float32, device=dev)
  rocks = torch.empty((E, R, h, h), dtype=torch.float32, device=dev)
  for b0 in range(0, E, block):
    n = min(block, E - b0)
    wl, rk, _ = synth.placement_batch(2 + 31 * (lo + b0), n, R, H, W, h)
    walls[b0:b0 + n] = torch.from_numpy(wl).to(dev)
    rocks[b0:b0 + n] = torch.from_numpy(rk).to(dev)
    goals[b0:b0 + n] = torch.from_numpy(synth.goals(3 + 31 * (lo + b0), n, H, W)).to(dev)
  level = capi.goal_level(goals)
  values = torch.empty((E, R, Ph, Ph), dtype=torch.float32, device=dev)
  shown = torch.empty((E, R, Ph, Ph), dtype=torch.float64, device=dev)
  actions = torch.empty((E, R), dtype=torch.int64, device=dev)
  best = torch.empty((E, 2), dtype=torch.int64, device=dev)
  lib, P_ = capi.lib, capi._P
  stream = P_(torch.cuda.current_stream().cuda_stream)
  mp_events = []

  def step(k, timed):
    if timed:
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
    capi._check(lib.srl_maxplus_f32(
      P_(walls.data_ptr()), P_(rocks.data_ptr()), P_(level.data_ptr()), P_(values.data_ptr()),
      E, R, H, W, h, 0.0, stream))
    if timed:
      b.record()
      mp_events.append((a, b))
    capi._check(lib.srl_mask_select_f32(
      P_(values.data_ptr()), P_(walls.data_ptr()), P_(goals.data_ptr()), P_(rocks.data_ptr()),
      P_(actions.data_ptr()), P_(shown.data_ptr()), P_(best.data_ptr()), E, R, H, W, h, 1, 0.75,
      stream))
  warmup = max(args.warmup, 3)
  elapsed_ms, sampler, clocks_how = timed_region(su, step, args.steps, warmup, preroll_s=0.1)
  maxplus_ms = sum(a.elapsed_time(b) for a, b in mp_events) / len(mp_events)

  # ---- end to end: host observations in, value maps + actions to the host ----------- #
  chunk = 128
  nch = (E + chunk - 1) // chunk
  ring = 3
  pin_in = [dict(walls=torch.empty((chunk, H, W)).pin_memory(),
                 goals=torch.empty((chunk, H, W)).pin_memory(),
                 rocks=torch.empty((chunk, R, h, h)).pin_memory()) for _ in range(ring)]
  pin_out = [torch.empty((chunk, R, Ph, Ph), dtype=torch.float64).pin_memory()
             for _ in range(ring)]
  act_pin = torch.empty((E, R), dtype=torch.int64).pin_memory()
  for k in range(ring):           # the staged host observations (cycled: same bytes per chunk)
    n = min(chunk, E)
    pin_in[k]['walls'][:n].copy_(walls[:n])
    pin_in[k]['goals'][:n].copy_(goals[:n])
    pin_in[k]['rocks'][:n].copy_(rocks[:n])
  torch.cuda.synchronize()
  copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
  scorer = baselines.PlacementScorer('height')
  dev_in = [dict(walls=torch.empty((chunk, H, W), device=dev),
                 goals=torch.empty((chunk, H, W), device=dev),
                 rocks=torch.empty((chunk, R, h, h), device=dev)) for _ in range(ring)]
  h2d = d2h = 0

  def sweep():
    nonlocal h2d, d2h
    main = torch.cuda.current_stream()
    h2d = d2h = 0
    free_in = [None] * ring
    for c in range(nch):
      n = min(chunk, E - c * chunk)
      s = c % ring
      with torch.cuda.stream(copy_in):
        if free_in[s] is not None:
          copy_in.wait_event(free_in[s])
        for name in ('walls', 'goals', 'rocks'):
          dev_in[s][name][:n].copy_(pin_in[s][name][:n], non_blocking=True)
          h2d += pin_in[s][name][:n].numel() * 4
        ready = torch.cuda.Event()
        ready.record(copy_in)
      main.wait_event(ready)
      out = scorer(dev_in[s]['walls'][:n], dev_in[s]['goals'][:n], dev_in[s]['rocks'][:n],
                   want_shown=True)
      done = torch.cuda.Event()
      done.record(main)
      free_in[s] = done
      with torch.cuda.stream(copy_out):
        copy_out.wait_event(done)
        pin_out[s][:n].copy_(out['shown'], non_blocking=True)
        act_pin[c * chunk:c * chunk + n].copy_(out['actions'], non_blocking=True)
        d2h += out['shown'].numel() * 8 + out['actions'].numel() * 8
        out['shown'].record_stream(copy_out)
        out['actions'].record_stream(copy_out)
    main.wait_stream(copy_out)
  sweep()
  su.barrier()
  e2e_steps = max(1, min(args.steps, 3))
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record()
  for _ in range(e2e_steps):
    sweep()
  ev1.record()
  torch.cuda.synchronize()
  e2e_ms = ev0.elapsed_time(ev1) / e2e_steps

  stats = su.sharding.gather_stats(
    [elapsed_ms, e2e_ms, maxplus_ms, float(evals), su.sharding.checksum(actions) % 2 ** 40], dev)
  if rank != 0:
    return None
  total = float(stats[:, 3].sum())
  elapsed_ms = float(stats[:, 0].max())
  ms_per_step = elapsed_ms / args.steps
  maxplus_ms = float(stats[:, 2].max())
  peaks, peak_src = measured_peaks()
  micro = {v_: capi.microbench_addmax(v_, 400) for v_ in (0, 2, 7)}
  peak_cells = max(micro.values())
  cells = evals * h * h
  kernel_s = maxplus_ms * 1e-3
  alg_bytes = 4 * (E * H * W + E * R * h * h + E * R * P)
  line = {
    'n_gpus': world, 'ms_per_step': ms_per_step, 'value': total / (ms_per_step * 1e-3),
    'clocks': sampler.summary(clocks_how),
    'e2e': {'value': total / (float(stats[:, 1].max()) * 1e-3), 'unit': w['unit'],
            'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
            'h2d_gbs_per_rank': h2d / (e2e_ms * 1e-3) / 1e9,
            'd2h_gbs_per_rank': d2h / (e2e_ms * 1e-3) / 1e9,
            'api': 'PlacementScorer(want_shown=True) over chunks of {} walls: pinned host '
                   'walls/goals/rocks in, float64 value maps (what stackrl.test.run stores per '
                   'policy, test.py:221-224) and actions out to pinned host memory, copies on '
                   'their own streams'.format(chunk)},
    'gpu_launches': 2 * args.steps,
    'kernels_per_step': ['maxplus_direct_kernel', 'mask_select_packed_kernel (9 chunks of 4 '
                                                   'views, value maps written)'],
    'roofline': {
      'bound': 'fp32-alu', 'kernel': 'maxplus_direct_kernel', 'kernel_ms': maxplus_ms,
      'share_of_step': maxplus_ms / ms_per_step,
      'achieved': 2 * cells / kernel_s / 1e12, 'peak': 2 * peak_cells / 1e12, 'unit': 'Tops/s',
      'frac': (cells / kernel_s) / peak_cells, 'traffic': ncu_traffic('maxplus_direct_kernel'),
      'peak_source': 'srl_microbench_addmax, best (add,max) issue rate measured in this run',
      'ops_per_eval': 2 * h * h,
      'hbm': {'achieved': alg_bytes / kernel_s / 1e9, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
              'frac': alg_bytes / kernel_s / 1e9 / peaks['hbm_gbs'], 'peak_source': peak_src}},
    'notes': {'walls_per_gpu': E, 'maps_per_gpu': E * R, 'evals_per_step_per_gpu': evals,
              'l2': 'inputs + outputs of one step {:.1f} GB >> 126 MB L2'.format(
                (alg_bytes + 8 * E * R * P) / 1e9),
              'shard_checksums': [int(c) for c in stats[:, 4].tolist()]},
  }
  return line


# --------------------------------------------------------------------------- #
# extras of the default run (N = 1): raster (config 3), secondary measurements
# --------------------------------------------------------------------------- #
def extra_metrics(torch, dev, cpu_baseline=True):
  from stackrl_b200 import baselines, capi, meshes, synth
  from stackrl_b200.observer import BatchedObserver
  out = {}
  # -- config 3: 4096 synthetic rocks, 32x32 px at 0.005 m/px ------------------- #
  n = 4096
  verts, tris = meshes.synthetic_rocks(4, n, max_dimension=0.16, frequency=10)   # 2000 tris
  bank = meshes.MeshBank()
  for k in range(n):
    bank.add(verts[k], tris)
  obs = BatchedObserver(bank, n, 1, overhead_resolution=128, object_resolution=32,
                        pixel_size=0.005, max_z=0.375, device=dev,
                        rock_cache_bytes=0)       # time the rasterisation, not the image table
  obs.observe_rocks(np.arange(n))
  torch.cuda.synchronize()
  ms = _time_loop(torch, lambda _: obs.observe_rocks(), 20)
  ntri, nvert = len(tris), verts.shape[1]
  bytes_per_rock = 12 * nvert + 12 * ntri + 4 * 32 * 32
  peaks, peak_src = measured_peaks()
  out['raster'] = {
    'workload': 'C3: {} synthetic rocks x {} tris ({} verts), 32x32 px at 0.005 m/px'.format(
      n, ntri, nvert),
    'parity': 'bitwise against oracle/csrc/oracle.c, the RESTATEMENT of the renderer; '
              'pybullet TinyRenderer parity unpinned (SURVEY 8c)',
    'rocks_per_s': n / (ms * 1e-3), 'tris_per_s': n * ntri / (ms * 1e-3), 'ms': ms,
    'roofline': {'bound': 'hbm', 'kernel': 'raster_kernel',
                 'achieved': n * bytes_per_rock / (ms * 1e-3) / 1e9,
                 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                 'frac': n * bytes_per_rock / (ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                 'bytes_per_rock': bytes_per_rock, 'peak_source': peak_src,
                 'traffic': ncu_traffic('raster_kernel')}}
  if cpu_baseline:
    from oracle import raster_np
    g = obs.geo
    spawn = ((0., 0., 0.375 + 0.16), (0., 0., 0., 1.))
    view = g.object_view(spawn, 0)
    t0 = time.perf_counter()
    for k in range(n):
      raster_np.render_depth(view, g.object_projection, g.object_h, g.object_w,
                             [(verts[k], tris, np.identity(3), np.array(spawn[0]))])
    dt = time.perf_counter() - t0
    out['raster']['cpu_baseline'] = {
      'value': n / dt, 'unit': 'rocks/s', 'cores': 1, 'kind': 'port',
      'sample': '{} of the same rocks, depth image only, oracle/csrc/oracle.c z-buffer '
                'through ctypes in {:.2f} s (restatement: TinyRenderer is absent)'.format(n, dt)}
  del obs, bank
  # -- config 2 on heightmaps as the rasteriser leaves them (multiples of 2^-14 m) ------ #
  w = WORKLOADS['c2']
  E, R, H, W, h = (w[k] for k in ('envs', 'rotations', 'H', 'W', 'h'))
  evals = E * R * positions(w)
  walls_h, rocks_h, _ = synth.placement_batch(0, E, R, H, W, h)
  q = np.float32(2.0 ** -14)
  wq = torch.from_numpy((np.round(walls_h / q) * q).astype('float32')).to(dev)
  rq = torch.from_numpy((np.round(rocks_h / q) * q).astype('float32')).to(dev)
  gq = torch.from_numpy(synth.goals(7, E, H, W)).to(dev)
  lq = capi.goal_level(gq)
  vq = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32, device=dev)
  same = torch.equal(capi.maxplus_f32(wq, rq, lq), capi.maxplus_f32(wq, rq, lq, quantum_log2=-14))
  ms_q = _time_loop(torch, lambda _: capi.maxplus_f32(wq, rq, lq, out=vq, quantum_log2=-14), 50)
  ms_f = _time_loop(torch, lambda _: capi.maxplus_f32(wq, rq, lq, out=vq), 50)
  out['quantised_heightmaps'] = {
    'workload': 'C2 shapes, walls/rocks rounded to multiples of 2^-14 m (what the float32 '
                'depth->elevation formulas of observer.py:259-260 produce); L2-warm single set',
    'maxplus_fixed_point_ms': ms_q, 'maxplus_fixed_point_evals_per_s': evals / (ms_q * 1e-3),
    'maxplus_float_ms': ms_f, 'bit_identical_to_float_sweep': bool(same)}
  # -- uint8 observations, device resident ------------------------------------------------ #
  w8 = torch.from_numpy(synth.to_dtype(walls_h, 'uint8')).to(dev)
  r8 = torch.from_numpy(synth.to_dtype(rocks_h, 'uint8')).to(dev)
  g8 = torch.from_numpy(synth.to_dtype(synth.goals(7, E, H, W), 'uint8')).to(dev)
  scorer8 = baselines.PlacementScorer('height')
  ms_8 = _time_loop(torch, lambda _: scorer8(w8, g8, r8), 30)
  out['uint8_observations'] = {
    'workload': 'C2 shapes cast like StackEnv._return (uint8, goal level 170): float64 '
                'max-plus values through the integer-key sweep + goal mask + arg-min, '
                'device-resident single set',
    'scorer_ms': ms_8, 'scorer_evals_per_s': evals / (ms_8 * 1e-3)}
  del w8, r8, g8
  # -- the DQN's Siamese correlation layer (nets/layers.py:21-38) -- #
  try:
    from stackrl_b200 import nets
    out['siam_correlation'] = nets.benchmark(
      torch, dev, tensor_peak_tflops=peaks.get('bf16_tflops'),
      tensor_peak_source=peak_src + ': dense bf16 (= fp16) cuBLAS burst figure')
  except Exception as exc:
    out['siam_correlation'] = {'error': repr(exc)}
  return out


def c1_reference_episode():
  """BASELINE config 1: the repo-default env for ONE 30-step episode with the
  lowest-placement baseline ('height', config.gin:118) -- the reference's own code on
  the static fake pybullet (pybullet is absent here: 'reference code on fake
  physics'), timed on one host core."""
  ns, kind = _reference()
  if ns is None:
    return {'unavailable': 'reference files not staged (oracle/make_ref.py)'}
  import glob
  urdfs = sorted(glob.glob(os.path.join(ns.root, 'stackrl/envs/data/generated', '[5-9]?_*.urdf')))
  env = ns.env.StackEnv(urdfs=urdfs, reward_params=2, dtype='uint8', seed=11)   # Stack-v0
  policy = ns.baselines.Baseline(method='height')
  t0 = time.perf_counter()
  obs = env.reset()
  steps, done, total = 0, False, 0.
  while not done:
    obs, reward, done, _ = env.step(policy(obs))
    total += reward
    steps += 1
  dt = time.perf_counter() - t0
  env.close()
  return {'workload': 'C1: Stack-v0 defaults (128x128 wall, 32x32 rock, uint8, 30 rocks), '
                      'Baseline("height"), one episode', 'kind': kind, 'steps': steps,
          'seconds': dt, 'env_steps_per_s': steps / dt, 'episode_return': float(total),
          'cores': 1, 'placement_evals_per_s': steps * 97 * 97 / dt,
          'note': 'fake physics (bodies stay where placed) and the oracle z-buffer as the '
                  'camera: pybullet is not installable here'}


def c1_gpu_episode(torch, dev):
  """The same configuration on this repo's path: BatchedStackEnv at the Stack-v0 defaults
  (128x128 wall, 32x32 rock, uint8 observations, 30 rocks, reward_params=2) + HeightPolicy,
  for config.gin's n_parallel = 2 environments and for a batch of 256 -- one warm-up episode,
  then one timed episode on the host clock (the small case is launch-latency bound: seven
  kernels and no device->host read per step).  Rocks: the reference's generated set where it
  is staged (data files only), else synthetic ones."""
  import glob
  from stackrl_b200 import envs, meshes
  d = os.path.join(ROOT, 'oracle', '_ref', 'stackrl', 'envs', 'data', 'generated')
  urdfs = sorted(glob.glob(os.path.join(d, '[5-9]?_*.urdf')))
  if urdfs:
    bank = meshes.MeshBank.from_urdfs(urdfs)
    rocks = '{} reference rocks ([5-9]?_*.urdf)'.format(len(bank))
  else:
    bank = meshes.MeshBank()
    v, t = meshes.synthetic_rocks(5, 64, 1, max_dimension=0.12)
    for k in range(64):
      bank.add(v[k], t)
    rocks = '64 synthetic rocks of 80 triangles'
  out = {'workload': 'C1: Stack-v0 defaults (128x128 wall, 32x32 rock, uint8, 30 rocks), '
                     'HeightPolicy, one episode', 'rocks': rocks, 'runs': []}
  for E in (2, 256):
    env = envs.BatchedStackEnv(bank, E, episode_length=30, reward_params=2, dtype='uint8',
                               seed=11, device=dev)
    policy = envs.HeightPolicy()
    for timed in (False, True):
      env.seed(11)                 # both passes play the first episode of seed 11
      env.reset()
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      total = torch.zeros(E, dtype=torch.float64, device=dev)
      steps, done = 0, False
      while not done:
        obs, reward, terminal = env.step(policy(env))
        total += reward
        steps += 1
        done = steps >= 30
      ret = total.cpu().numpy()                      # the one device->host read of the episode
      dt = time.perf_counter() - t0
    assert bool(terminal.all())
    out['runs'].append({'envs': E, 'steps': steps, 'seconds': dt,
                        'env_steps_per_s': steps * E / dt, 'ms_per_step': 1e3 * dt / steps,
                        'placement_evals_per_s': steps * E * 97 * 97 / dt,
                        'mean_episode_return': float(ret.mean()),
                        # environment 0 is seeded like the reference run below (seed 11): same
                        # rock order, goal and static physics, so the same episode
                        'episode_return_env0': float(ret[0])})
  return out


# --------------------------------------------------------------------------- #
def run_graft(args, rank, local_rank, world):
  su = Setup(args, rank, local_rank, world)
  line = {'c2': run_c2, 'c4': run_c4, 'c5': run_c5}[args.workload](args, su)
  if line is None:
    su.finish()
    return
  w = WORKLOADS[args.workload]
  head = {
    'metric': w['metric'], 'value': line.pop('value'), 'unit': w['unit'],
    'n_gpus': line.pop('n_gpus'), 'steps': args.steps, 'warmup': max(args.warmup, 3),
    'ms_per_step': line.pop('ms_per_step'), 'higher_is_better': True,
    'scaling': w['scaling'], 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': config_of(args.workload),
  }
  head.update(line)
  if world == 1 and not args.no_extra and args.workload == 'c2':
    try:
      head['extra'] = extra_metrics(su.torch, su.dev, cpu_baseline=not args.no_cpu_baseline)
    except Exception as exc:   # the headline must survive a failure of the extras
      head['extra'] = {'error': repr(exc)}
    try:
      head['extra']['c1_gpu_episode'] = c1_gpu_episode(su.torch, su.dev)
    except Exception as exc:
      head['extra']['c1_gpu_episode'] = {'error': repr(exc)}
    if not args.no_cpu_baseline:
      try:
        head['extra']['c1_reference_episode'] = c1_reference_episode()
      except Exception as exc:
        head['extra']['c1_reference_episode'] = {'error': repr(exc)}
      # environment 0 of the GPU run and the reference run play the same seed-11 episode
      ref_c1, gpu_c1 = head['extra']['c1_reference_episode'], head['extra']['c1_gpu_episode']
      if 'episode_return' in ref_c1 and gpu_c1.get('runs'):
        gpu_c1['reference_episode_return'] = ref_c1['episode_return']
        gpu_c1['abs_difference_env0'] = abs(gpu_c1['runs'][0]['episode_return_env0'] -
                                            ref_c1['episode_return'])
  if world == 1 and not args.no_cpu_baseline:
    cores = os.cpu_count() or 1
    v, s_per_step, sample, kind = cpu_arm(args.workload, 2, 1, cores)
    head['cpu_baseline'] = {
      'value': v, 'unit': w['unit'], 'cores': cores, 'kind': kind,
      'sample': '2 steps of {:.1f} s: {}'.format(s_per_step, sample)}
  print(json.dumps(head))
  su.finish()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=10)
  ap.add_argument('--impl', default='graft', choices=['graft', 'reference'])
  ap.add_argument('--workload', default='c2', choices=['c2', 'c4', 'c5'])
  ap.add_argument('--envs', type=int, default=0,
                  help='override the TOTAL number of environments / walls of c4 / c5')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--eager', action='store_true',
                  help='c4: launch the step kernel by kernel instead of replaying its CUDA graph')
  ap.add_argument('--no-extra', action='store_true')
  ap.add_argument('--no-bind', action='store_true', help='do not pin ranks to core slices')
  ap.add_argument('--fused', action='store_true',
                  help='c2: run the single fully fused kernel (srl_score_f32)')
  ap.add_argument('--separate', action='store_true',
                  help='c2: run the three separate kernels instead of two')
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if args.workload != 'c2' and args.steps == 200:
    args.steps = 60 if args.workload == 'c4' else 5
  if args.impl == 'reference':
    run_reference(args, rank, world)
  else:
    run_graft(args, rank, local_rank, world)


if __name__ == '__main__':
  main()
