"""BatchedStackEnv steps at the registered environments' geometry (128x128 wall, 32x32 rock,
uint8) for a batch, to compare the wall-raster kernels there.
python tools/bench_env_c1_geometry.py [envs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import envs, meshes

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device('cuda')
bank = meshes.MeshBank()
v, t = meshes.synthetic_rocks(5, 64, 1, max_dimension=0.12)      # 80 triangles, ~30 px across
for k in range(64):
  bank.add(v[k], t)
env = envs.BatchedStackEnv(bank, E, episode_length=30, dtype='uint8', seed=3, device=dev,
                           vector_rng=True)
policy = envs.HeightPolicy()
env.reset()
for _ in range(4):
  env.step(policy(env))
obs = env.obs
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.
for _ in range(12):
  action = policy(env)
  obs.poses_device(None, action)
  obs.advance()
  a.record()
  obs.observe_walls(True)
  b.record()
  obs.observe_rocks()
  env._reward_and_pack()
  env._advance_host()
  torch.cuda.synchronize()
  tot += a.elapsed_time(b)
print('E=%d wall raster (appended rock): %.3f ms per step' % (E, tot / 12))
