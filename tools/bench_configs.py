"""Max-plus / scoring throughput across the BASELINE geometries (not the headline
bench; a sweep used while tuning).  python tools/bench_configs.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stackrl_b200 import baselines, capi, synth

CONFIGS = [
  ('C2 32x32/16 R8', 4096, 8, 32, 32, 16),
  ('C4 64x64/16 R1', 8192, 1, 64, 64, 16),
  ('C4 64x64/16 R8', 2048, 8, 64, 64, 16),
  ('C5 128x128/32 R36', 592, 36, 128, 128, 32),   # 4736 CTAs = 16 whole waves
  ('C1 128x128/32 R1', 1024, 1, 128, 128, 32),
  ('C1 single obs', 1, 1, 128, 128, 32),
  ('odd 40x56/12 R3', 1024, 3, 40, 56, 12),
]


def timeit(fn, reps):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


def main():
  dev = torch.device('cuda')
  peak = max(capi.microbench_addmax(v, 400) for v in (2, 7))
  print('microbench peak %.3e cells/s' % peak)
  for name, E, R, H, W, h in CONFIGS:
    walls, rocks, level = synth.placement_batch(0, E, R, H, W, h)
    goals = synth.goals(1, E, H, W)
    wd, gd, rd, ld = (torch.from_numpy(x).to(dev) for x in (walls, goals, rocks, level))
    P = (H - h + 1) * (W - h + 1)
    out = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32, device=dev)
    reps = 20 if E * R * P * h * h > 1e9 else 50
    ms = timeit(lambda: capi.maxplus_f32(wd, rd, ld, out=out), reps)
    evals = E * R * P
    scorer = baselines.PlacementScorer()
    ms_full = timeit(lambda: scorer(wd, gd, rd), reps)
    print('%-22s maxplus %8.3f ms  %.3e evals/s  %5.1f%% of peak | scorer %8.3f ms %.3e evals/s' % (
      name, ms, evals / ms * 1e3, 100 * evals * h * h / (ms * 1e-3) / peak, ms_full,
      evals / ms_full * 1e3))


if __name__ == '__main__':
  main()
