"""Rasterisation throughput (BASELINE config 3 geometry). python tools/bench_raster.py [n] [subdiv] [reps]
(subdiv >= 5 is read as a geodesic frequency: 10 = 2000 triangles / 1002 vertices)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stackrl_b200 import capi, meshes
from stackrl_b200.observer import BatchedObserver


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
  sub = int(sys.argv[2]) if len(sys.argv) > 2 else 3
  reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
  dev = torch.device('cuda')
  verts, tris = meshes.synthetic_rocks(4, n, sub, max_dimension=0.16) if sub < 5 else \
    meshes.synthetic_rocks(4, n, max_dimension=0.16, frequency=sub)
  bank = meshes.MeshBank()
  for k in range(n):
    bank.add(verts[k], tris)
  obs = BatchedObserver(bank, n, 1, overhead_resolution=64, object_resolution=32,
                        pixel_size=0.005, max_z=0.375, orientation_freedom=0, device=dev)
  obs.observe_rocks(np.arange(n))
  torch.cuda.synchronize()
  g = obs.geo
  out = obs.rocks.view(n, g.object_h, g.object_w)
  def run():
    capi.raster(obs._verts, obs._tris, obs._rock_inst, obs._rock_jobs, g.object_h, g.object_w,
                capi.RASTER_ROCK, out=out, max_cached_verts=verts.shape[1])
  for _ in range(3):
    run()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    run()
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / reps
  nb = 12 * verts.shape[1] + 12 * len(tris) + 4 * 32 * 32
  print('mode %s: %d rocks x %d tris: %.3f ms  %.3e rocks/s  %.3e tris/s  %.1f GB/s algorithmic' % (
    os.environ.get('SRL_RASTER_MODE', '0'), n, len(tris), ms, n / ms * 1e3, n * len(tris) / ms * 1e3, n * nb / ms / 1e6))


if __name__ == '__main__':
  main()
