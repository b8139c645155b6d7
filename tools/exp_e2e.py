"""Where does the end-to-end step of HostPipeline go?  Chunk-count sweep for float32
(goal rectangles) and uint8 observations, next to the device-only scoring time of the
same batch cut the same way (no copies) and the bare H2D copy time.
python tools/exp_e2e.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stackrl_b200 import baselines, synth

E, R, H, W, h = 4096, 8, 32, 32, 16
dev = torch.device('cuda')
walls_h, rocks_h, levels_h = synth.placement_batch(0, E, R, H, W, h)
rects_h = synth.goal_rects(7, E, H, W)
goals_h = synth.goals(7, E, H, W)


def timed(fn, n=30):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / n


for dtype_name, rects in (('float32', True), ('uint8', False)):
  dt = {'float32': torch.float32, 'uint8': torch.uint8}[dtype_name]
  for chunks in (1, 2, 3, 4, 6, 8):
    scorer = baselines.PlacementScorer('height')
    pipe = baselines.HostPipeline(scorer, E, R, H, W, h, chunks=chunks, device=dev, dtype=dt,
                                  goal_rects=rects)
    if rects:
      pipe.stage(walls_h, rects_h, rocks_h, levels_h)
    else:
      pipe.stage(*[synth.to_dtype(x, dtype_name) for x in (walls_h, goals_h, rocks_h)])
    for _ in range(3):
      pipe.run()
    eager = timed(pipe.run)
    pipe.capture()
    graph = timed(pipe.run)
    # the copies alone / the kernels alone on the same chunking
    def copies():
      for pin, d in zip(pipe.pin_slabs, pipe.dev_slabs):
        d.copy_(pin, non_blocking=True)
      torch.cuda.current_stream().synchronize()
    def kernels():
      for k, dev_in in enumerate(pipe.dev_in):
        if rects:
          from stackrl_b200 import capi
          g = capi.fill_goals(dev_in['rects'], dev_in['levels'], pipe.goal_planes[k])
          scorer(dev_in['walls'], g, dev_in['rocks'], level=dev_in['levels'])
        else:
          scorer(dev_in['walls'], dev_in['goals'], dev_in['rocks'])
      torch.cuda.current_stream().synchronize()
    print('%-8s chunks=%2d  eager %.3f ms  graph %.3f ms  copies only %.3f ms  kernels only '
          '%.3f ms  (%.1f MB H2D)' % (dtype_name, chunks, eager, graph, timed(copies),
                                      timed(kernels), pipe.h2d_bytes / 1e6), flush=True)
