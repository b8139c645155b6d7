"""Run the C2 scoring step (max-plus + mask_select) a few times (ncu target).
python tools/run_step.py [reps] [sets]: `sets` distinct input/output sets are cycled
like bench.py does (8 sets = 1.1 GB > L2), so that a kernel captured late in the run
sees the cache state of the benchmarked configuration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stackrl_b200 import baselines, capi, synth

E, R, H, W, h = 4096, 8, 32, 32, 16
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
nsets = int(sys.argv[2]) if len(sys.argv) > 2 else 1
walls, rocks, level = synth.placement_batch(0, E, R, H, W, h)
goals = synth.goals(7, E, H, W)
dev = torch.device('cuda')
sets = []
for s in range(nsets):
  g = torch.from_numpy(goals).to(dev)
  sets.append(dict(walls=torch.from_numpy(np.roll(walls, s, axis=0)).to(dev), goals=g,
                   rocks=torch.from_numpy(np.roll(rocks, -s, axis=0)).to(dev),
                   level=capi.goal_level(g),
                   values=torch.empty((E, R, H - h + 1, W - h + 1), device=dev)))


def step(k):
  s = sets[k % nsets]
  capi.maxplus_f32(s['walls'], s['rocks'], s['level'], out=s['values'])
  capi.mask_select(s['values'], s['walls'], s['goals'], s['rocks'], want_shown=False)


for k in range(3 * nsets):
  step(k)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for k in range(reps):
  step(k)
b.record()
torch.cuda.synchronize()
print('step %.4f ms (%d sets cycled)' % (a.elapsed_time(b) / reps, nsets))
