"""mask_select alone at config-2 shapes, for several batch sizes (L2-resident vs not)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import capi, synth

R, H, W, h = 8, 32, 32, 16
dev = torch.device('cuda')
for E in (512, 1024, 4096, 16384):
  walls, rocks, level = synth.placement_batch(0, E, R, H, W, h)
  goals = synth.goals(7, E, H, W)
  wd, gd, rd, ld = (torch.from_numpy(x).to(dev) for x in (walls, goals, rocks, level))
  values = capi.maxplus_f32(wd, rd, ld)
  for _ in range(3):
    capi.mask_select(values, wd, gd, rd, minorder=1, overlap_threshold=0.75)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 50
  a.record()
  for _ in range(reps):
    capi.mask_select(values, wd, gd, rd, minorder=1, overlap_threshold=0.75)
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / reps
  print('E=%6d  %.4f ms  %.2f ns/env  %.1f GB/s' % (
    E, ms, ms * 1e6 / E, E * (R * 289 * 4 + 2 * H * W * 4 + R * h * h * 4) / ms / 1e6))
