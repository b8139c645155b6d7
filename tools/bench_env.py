"""Device time of one environment observation (C4 slice geometry): wall raster, rock
raster, reward terms, packed observation.  python tools/bench_env.py [E]"""
import os
import sys

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
                           device=dev, vector_rng=True)
policy = envs.HeightPolicy()
env.reset()
for _ in range(6):
  env.step(policy(env))
torch.cuda.synchronize()


def timeit(fn, reps=20):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


print('mode %s, E=%d' % (os.environ.get('SRL_RASTER_MODE', '0'), E))
print('walls (full redraw) %.3f ms' % timeit(env.obs.observe_walls))
print('walls (appended rock) %.3f ms' % timeit(lambda: env.obs.observe_walls(True)))
print('rocks  %.3f ms' % timeit(env.obs.observe_rocks))
print('reward+pack %.3f ms' % timeit(env._reward_and_pack))
print('policy %.3f ms' % timeit(lambda: policy(env)))
view = torch.zeros(E, dtype=torch.int64, device=dev)
print('poses+advance (state frozen: done envs skip) %.3f ms' % timeit(
  lambda: (env.obs.poses_device(None, view), env.obs.advance())))


def episode(step):
  """Wall-clock of whole episodes: reset + 12 (policy + step), host bookkeeping included."""
  import time
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  n = 0
  for _ in range(3):
    env.reset()
    for _ in range(12):
      step()
      n += 1
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / n * 1e3


ms = episode(lambda: env.step(policy(env)))
print('eager step (policy + step, resets amortised) %.3f ms  %.3e env steps/s' % (ms, E / ms * 1e3))
env.reset()
env.step(policy(env))
env.capture(policy)
ms = episode(env.step_policy)
print('graph step (policy + step, resets amortised) %.3f ms  %.3e env steps/s' % (ms, E / ms * 1e3))
