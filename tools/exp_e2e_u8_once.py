"""One eager uint8 HostPipeline step (8 chunks) after warm-up: ncu launch-list target."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import baselines, synth

E, R, H, W, h = 4096, 8, 32, 32, 16
dev = torch.device('cuda')
walls_h, rocks_h, levels_h = synth.placement_batch(0, E, R, H, W, h)
goals_h = synth.goals(7, E, H, W)
scorer = baselines.PlacementScorer('height')
pipe = baselines.HostPipeline(scorer, E, R, H, W, h, chunks=8, device=dev, dtype=torch.uint8)
pipe.stage(*[synth.to_dtype(x, 'uint8') for x in (walls_h, goals_h, rocks_h)])
for _ in range(3):
  pipe.run()
print('ok')
