#!/usr/bin/env python
"""bench.py -- placement evaluations / s of the max-plus scoring path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): 4096 environments per GPU,
32x32 wall heightmap, 16x16 rock underside map, 8 rotations -> 4096*8*17*17 =
9 469 952 placement evaluations per step per GPU.  A step is one pass of the
scoring hot path over one batch of synthetic observations.  Environments are
independent, so N GPUs run N shards with no data-path collective ("weak"
scaling); rank 0 gathers the timing with one all-reduce (MAX).

Prints ONE JSON line (see the contract in the task statement / DESIGN.md).
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
          '{h}x{h} rock, {rotations} rotations').format(**CFG)


def evals_per_step():
  P = (CFG['H'] - CFG['h'] + 1) * (CFG['W'] - CFG['h'] + 1)
  return CFG['envs'] * CFG['rotations'] * P


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return json.load(f), 'measured (MEASURED_PEAKS.json)'
  return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


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
# CPU baseline: the oracle's loop-form restatement of baselines.height
# --------------------------------------------------------------------------- #
def _cpu_maps(args):
  """Worker: score `count` (env, rotation) maps with the reference-shaped
  Python-loop + numpy max-plus (oracle.scoring_np.height_loop)."""
  seed, count = args
  from oracle import scoring_np
  from stackrl_b200 import synth
  walls, rocks, level = synth.placement_batch(seed, count, 1, CFG['H'], CFG['W'], CFG['h'])
  obs = []
  for e in range(count):
    goal = np.full(walls.shape[1:], level[e], dtype='float32')
    obs.append((np.stack([walls[e], goal], -1), rocks[e, 0][..., None]))
  t0 = time.perf_counter()
  for o in obs:
    scoring_np.height_loop(o)
  return time.perf_counter() - t0


def cpu_baseline(maps_per_core, cores):
  P = evals_per_step() // (CFG['envs'] * CFG['rotations'])
  if cores == 1:
    dt = _cpu_maps((123, maps_per_core))
    wall = dt
  else:
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
      pool.map(_cpu_maps, [(1, 2)] * cores)            # import + warm-up
      t0 = time.perf_counter()
      pool.map(_cpu_maps, [(200 + k, maps_per_core) for k in range(cores)])
      wall = time.perf_counter() - t0
  value = maps_per_core * cores * P / wall
  return value, wall


def run_reference(args, rank, world):
  """--impl reference: the reference's CPU algorithm (oracle port; the Python
  reference itself cannot travel to the GPU box) on all host cores."""
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  maps_per_core = 48          # ~0.15 s of work per core per step at ~3 ms/map
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
  sample = ('{} maps ({}x{} wall, {}x{} rock) per step over {} processes; '
            'oracle.scoring_np.height_loop (numpy port of baselines.py:28-43)'
            ).format(cores * maps_per_core, CFG['H'], CFG['W'], CFG['h'], CFG['h'], cores)
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
def run_graft(args, rank, local_rank, world):
  import torch
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device: stackrl_b200 has no CPU path')
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  dist = None
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)

  from stackrl_b200 import capi, synth

  E, R, H, W, h = (CFG[k] for k in ('envs', 'rotations', 'H', 'W', 'h'))
  P = (H - h + 1) * (W - h + 1)

  # Synthetic batches: each rank owns its shard of environments; the
  # NSETS sets differ by a cheap device-side perturbation of one host batch.
  walls_h, rocks_h, level_h = synth.placement_batch(1000 * rank, E, R, H, W, h)
  sets = []
  for s in range(NSETS):
    w = torch.from_numpy(walls_h).to(dev)
    if s:
      w = torch.roll(w, shifts=s, dims=0).contiguous()
    r = torch.from_numpy(rocks_h).to(dev)
    lvl = torch.from_numpy(level_h).to(dev)
    out = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32, device=dev)
    sets.append((w, r, lvl, out))
  set_bytes = sum(t.numel() * t.element_size() for t in sets[0])

  def step(k):
    w, r, lvl, out = sets[k % NSETS]
    capi.maxplus_f32(w, r, lvl, out=out)

  def barrier():
    if dist is not None:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- warm-up -------------------------------------------------------------- #
  for k in range(max(args.warmup, 3)):
    step(k)
  barrier()

  # ---- timed region (device-resident inputs) -------------------------------- #
  sampler = ClockSampler(local_rank)
  ev0 = torch.cuda.Event(enable_timing=True)
  ev1 = torch.cuda.Event(enable_timing=True)
  barrier()
  sampler.start()
  ev0.record()
  for k in range(args.steps):
    step(k)
  ev1.record()
  torch.cuda.synchronize()
  sampler.stop()
  elapsed_ms = ev0.elapsed_time(ev1)
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
  if dist is not None:
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
  ms_per_step = elapsed_ms / args.steps
  value = world * evals_per_step() / (ms_per_step * 1e-3)

  # ---- end to end: host buffers in, host result out -------------------------- #
  walls_p = torch.from_numpy(walls_h).pin_memory()
  rocks_p = torch.from_numpy(rocks_h).pin_memory()
  level_p = torch.from_numpy(level_h).pin_memory()
  out_p = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32).pin_memory()
  w_d, r_d, l_d, o_d = sets[0]
  h2d = sum(t.numel() * t.element_size() for t in (walls_p, rocks_p, level_p))
  d2h = out_p.numel() * out_p.element_size()

  def e2e_step():
    w_d.copy_(walls_p, non_blocking=True)
    r_d.copy_(rocks_p, non_blocking=True)
    l_d.copy_(level_p, non_blocking=True)
    capi.maxplus_f32(w_d, r_d, l_d, out=o_d)
    out_p.copy_(o_d, non_blocking=True)

  e2e_steps = max(3, min(args.steps, 50))
  for _ in range(3):
    e2e_step()
  barrier()
  ev0.record()
  for _ in range(e2e_steps):
    e2e_step()
  ev1.record()
  torch.cuda.synchronize()
  e2e_ms = ev0.elapsed_time(ev1)
  if dist is not None:
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
  e2e_value = world * evals_per_step() / (e2e_ms / e2e_steps * 1e-3)

  if rank != 0:
    if dist is not None:
      dist.destroy_process_group()
    return

  # ---- roofline of the dominant kernel --------------------------------------- #
  peaks, peak_src = measured_peaks()
  cells = evals_per_step() * h * h                  # (add, max) cells per launch
  kernel_s = ms_per_step * 1e-3                     # the step IS one launch
  peak_cells = max(capi.microbench_addmax(v, 400) for v in (0, 1, 2))
  alg_bytes = 4 * (E * H * W + E * R * h * h + E * R * P)
  roofline = {
    'bound': 'fp32-alu',
    'achieved': 2 * cells / kernel_s / 1e12,
    'peak': 2 * peak_cells / 1e12,
    'unit': 'Tops/s',
    'frac': (cells / kernel_s) / peak_cells,
    'traffic': None,
    'peak_source': 'srl_microbench_addmax (FADD2+FMNMX3 issue rate, measured in this run)',
    'ops_per_eval': 2 * h * h,
    'hbm': {'achieved': alg_bytes / kernel_s / 1e9, 'peak': peaks['hbm_gbs'],
            'unit': 'GB/s', 'frac': alg_bytes / kernel_s / 1e9 / peaks['hbm_gbs'],
            'bytes_per_eval': alg_bytes / evals_per_step(), 'peak_source': peak_src},
  }

  line = {
    'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
    'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
    'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
    'dtype': 'f32', 'data': 'synthetic',
    'config': {'workload': workload_name(), 'evals_per_step_per_gpu': evals_per_step(),
               'l2': '{} distinct input/output sets cycled, {:.0f} MB each ({:.0f} MB total '
                     '> 126 MB L2)'.format(NSETS, set_bytes / 1e6, NSETS * set_bytes / 1e6),
               'parallelism': 'env-sharded x{}'.format(world)},
    'clocks': sampler.summary(clocks_how),
    'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
            'd2h_bytes_per_step': d2h, 'steps': e2e_steps},
    'gpu_launches': args.steps,
    'roofline': roofline,
  }
  if world == 1 and not args.no_cpu_baseline:
    maps = 1536
    v, wall = cpu_baseline(maps, 1)
    line['cpu_baseline'] = {
      'value': v, 'unit': UNIT, 'cores': 1, 'kind': 'port',
      'sample': '{} maps of the same shapes in {:.1f} s on 1 core; '
                'oracle.scoring_np.height_loop (numpy port of baselines.py:28-43); '
                'host has {} cores'.format(maps, wall, os.cpu_count())}
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
