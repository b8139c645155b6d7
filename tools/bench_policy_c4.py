"""The two kernels of the height policy at config-4 geometry on maps an environment rollout
produced (rasteriser-quantised): max-plus with and without the quantum hint, mask_select.
python tools/bench_policy_c4.py [envs] [rollout steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import capi, envs, meshes

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device('cuda')
bank = meshes.MeshBank()
v, t = meshes.synthetic_rocks(5, 64, 1, max_dimension=0.12)
for k in range(64):
  bank.add(v[k], t)
env = envs.BatchedStackEnv(bank, E, episode_length=30, observable_size_ratio=4,
                           resolution_factor=4, dtype='float32', rewarder='iou', seed=5,
                           device=dev, vector_rng=True)
policy = envs.HeightPolicy()
env.reset()
for _ in range(steps):
  env.step(policy(env))
walls, goals, rocks = env.planes()
level = env._goal_z_d


def timed(fn, reps=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


P = (walls.shape[1] - rocks.shape[2] + 1) ** 2
out = torch.empty((E, rocks.shape[1], walls.shape[1] - rocks.shape[2] + 1,
                   walls.shape[2] - rocks.shape[2] + 1), dtype=torch.float32, device=dev)
for q in (envs.HEIGHT_QUANTUM_LOG2, None):
  ms = timed(lambda: capi.maxplus_f32(walls, rocks, level, out=out, quantum_log2=q))
  print('maxplus quantum_log2=%s: %.3f ms  %.3e evals/s' % (q, ms, E * P / ms * 1e3))
values = capi.maxplus_f32(walls, rocks, level, quantum_log2=envs.HEIGHT_QUANTUM_LOG2)
ms = timed(lambda: capi.mask_select(values, walls, goals, rocks, minorder=1,
                                    overlap_threshold=0.75, want_shown=False))
print('mask_select: %.3f ms  %.2f ns/env' % (ms, ms * 1e6 / E))

# uint8 observations (the registered environments' dtype): quantised planes, float64 values
env8 = envs.BatchedStackEnv(bank, E, episode_length=30, observable_size_ratio=4,
                            resolution_factor=4, dtype='uint8', rewarder='iou', seed=5,
                            device=dev, vector_rng=True)
env8.reset()
for _ in range(steps):
  env8.step(policy(env8))
ms = timed(lambda: env8.planes_u8())
print('planes_u8: %.3f ms' % ms)
w8, g8, r8 = env8.planes_u8()
l8 = env8._level8_d
out8 = torch.empty(out.shape, dtype=torch.float64, device=dev)
ms = timed(lambda: capi.maxplus_u8(w8, r8, l8, out=out8))
print('maxplus_u8: %.3f ms  %.3e evals/s' % (ms, E * P / ms * 1e3))
ms = timed(lambda: capi.mask_select(out8, w8, g8, r8, minorder=1, overlap_threshold=0.75,
                                    want_shown=False))
print('mask_select f64/u8: %.3f ms' % ms)
ms = timed(lambda: policy(env8))
print('HeightPolicy(uint8 env): %.3f ms' % ms)
