"""cProfile of BatchedStackEnv.step host glue (C4 slice geometry)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import envs, meshes

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda')
bank = meshes.MeshBank()
v, t = meshes.synthetic_rocks(5, 64, 1, max_dimension=0.12)
for k in range(64):
  bank.add(v[k], t)
env = envs.BatchedStackEnv(bank, E, episode_length=12, observable_size_ratio=4,
                           resolution_factor=4, dtype='float32', rewarder='iou', seed=5,
                           device=dev)
policy = envs.HeightPolicy()
env.reset()
for _ in range(3):
  env.step(policy(env))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
  a = policy(env)
  torch.cuda.synchronize()
t1 = time.perf_counter()
print('policy %.3f ms' % ((t1 - t0) / 4 * 1e3))
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
for _ in range(6):
  env.step(a)
torch.cuda.synchronize()
t1 = time.perf_counter()
pr.disable()
print('step %.3f ms' % ((t1 - t0) / 6 * 1e3))
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
