"""Run the C2 scoring step (max-plus + mask_select) a few times (ncu target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import baselines, capi, synth

E, R, H, W, h = 4096, 8, 32, 32, 16
walls, rocks, level = synth.placement_batch(0, E, R, H, W, h)
goals = synth.goals(7, E, H, W)
dev = torch.device('cuda')
wd, gd, rd = (torch.from_numpy(x).to(dev) for x in (walls, goals, rocks))
scorer = baselines.PlacementScorer('height')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(3):
  scorer(wd, gd, rd)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
  scorer(wd, gd, rd)
b.record()
torch.cuda.synchronize()
print('step %.4f ms' % (a.elapsed_time(b) / reps))
