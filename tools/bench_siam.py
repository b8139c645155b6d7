"""Siamese correlation layer throughput (DQN default geometry: 128x128x16 wall
features, 32x32x16 rock features).  python tools/bench_siam.py [B] [C] [H] [h]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stackrl_b200 import capi


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
  C = int(sys.argv[2]) if len(sys.argv) > 2 else 16
  H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
  h = int(sys.argv[4]) if len(sys.argv) > 4 else 32
  g = torch.Generator(device='cuda').manual_seed(0)
  x = torch.randn((B, H, H, C), device='cuda', generator=g)
  w = torch.randn((B, h, h, C), device='cuda', generator=g)
  out = torch.empty((B, H - h + 1, H - h + 1, 1), device='cuda')
  for _ in range(3):
    capi.siam_correlation_f32(x, w, out=out)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 10
  a.record()
  for _ in range(reps):
    capi.siam_correlation_f32(x, w, out=out)
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / reps
  flops = 2.0 * B * (H - h + 1) ** 2 * h * h * C
  print('B=%d %dx%dx%d * %dx%dx%d: %.3f ms  %.2f TFLOP/s fp32 (FFMA peak 148 SMs x 128 x 2 x 1.965 GHz = 74.4)  %.3e samples/s' % (
    B, H, H, C, h, h, C, ms, flops / ms / 1e9, B / ms * 1e3))
  # CPU baseline on a small sample: the oracle's einsum form (float64), one core
  from oracle import nets_np
  n = min(B, 2)
  t0 = time.perf_counter()
  nets_np.correlation(x[:n].cpu().numpy(), w[:n].cpu().numpy())
  dt = time.perf_counter() - t0
  print('oracle (numpy einsum, float64): %.3e samples/s' % (n / dt))


if __name__ == '__main__':
  main()
