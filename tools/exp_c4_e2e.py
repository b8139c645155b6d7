"""Per-step wall-clock of the c4 end-to-end loop (host action in, reward/terminal/next action
out), to find where sporadic slow runs come from.  python tools/exp_c4_e2e.py [envs] [capture]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import envs, meshes

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
capture = len(sys.argv) > 2 and sys.argv[2] == '1'
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
for _ in range(20):
  env.step(policy(env))
if capture:
  env.capture(policy)
  for _ in range(5):
    env.step_policy()
torch.cuda.synchronize()
act_pin = torch.empty(E, dtype=torch.int64).pin_memory()
rew_pin = torch.empty(E, dtype=torch.float32).pin_memory()
term_pin = torch.empty(E, dtype=torch.uint8).pin_memory()
for rep in range(3):
  env.reset()
  act_pin.copy_(policy(env))
  torch.cuda.synchronize()
  times = []
  for _ in range(28):
    t0 = time.perf_counter()
    action = act_pin.to(dev, non_blocking=True)
    t1 = time.perf_counter()
    o, r, t_ = env.step(action)
    t2 = time.perf_counter()
    rew_pin.copy_(r, non_blocking=True)
    term_pin.copy_(t_.view(torch.uint8), non_blocking=True)
    t3 = time.perf_counter()
    act_pin.copy_(policy(env), non_blocking=True)
    t4 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    t5 = time.perf_counter()
    times.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4))
  tot = [sum(x) * 1e3 for x in times]
  print('rep %d capture=%d: mean %.3f ms  min %.3f  max %.3f | h2d %.3f step %.3f d2h %.3f policy %.3f sync %.3f' % (
    rep, capture, sum(tot) / len(tot), min(tot), max(tot),
    *[1e3 * sum(x[k] for x in times) / len(times) for k in range(5)]), flush=True)
