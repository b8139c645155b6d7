"""A few BatchedStackEnv steps at config-4 geometry (for ncu captures of the step kernels).
python tools/run_env_steps.py [envs] [steps] [persistent observation buffers: 0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import envs, meshes

E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
persistent = len(sys.argv) > 3 and sys.argv[3] == '1'
dev = torch.device('cuda')
bank = meshes.MeshBank()
v, t = meshes.synthetic_rocks(5, 64, 1, max_dimension=0.12)
for k in range(64):
  bank.add(v[k], t)
env = envs.BatchedStackEnv(bank, E, episode_length=30, observable_size_ratio=4,
                           resolution_factor=4, dtype='float32', rewarder='iou', seed=5,
                           device=dev, vector_rng=True, persistent_observation=persistent)
policy = envs.HeightPolicy()
env.reset()
for _ in range(steps):
  env.step(policy(env))
torch.cuda.synchronize()
print('ok', E, steps)
