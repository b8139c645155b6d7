"""Timing of the secondary kernels (uint8 max-plus, difference, correlate). python tools/bench_misc.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stackrl_b200 import capi, synth


def timeit(fn, reps=10):
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


def main():
  dev = torch.device('cuda')
  for name, E, R, H, W, h in [('C2', 4096, 8, 32, 32, 16), ('C1', 256, 1, 128, 128, 32),
                              ('C1 single', 1, 1, 128, 128, 32)]:
    walls, rocks, level = synth.placement_batch(0, E, R, H, W, h)
    P = (H - h + 1) * (W - h + 1)
    evals = E * R * P
    wd, rd, ld = (torch.from_numpy(x).to(dev) for x in (walls, rocks, level))
    w8 = torch.from_numpy(synth.to_dtype(walls, 'uint8')).to(dev)
    r8 = torch.from_numpy(synth.to_dtype(rocks, 'uint8')).to(dev)
    l8 = torch.full((E,), 170, dtype=torch.uint8, device=dev)
    ms = timeit(lambda: capi.maxplus_u8(w8, r8, l8))
    print('%-10s maxplus_u8   %8.3f ms  %.3e evals/s' % (name, ms, evals / ms * 1e3))
    wts = capi.difference_weights(rd, ld)
    ms = timeit(lambda: capi.difference_f32(wd, rd, ld, wts), 5)
    print('%-10s difference   %8.3f ms  %.3e evals/s' % (name, ms, evals / ms * 1e3))
    ms = timeit(lambda: capi.correlate_f32(wd, rd, ld, want_coef=False), 5)
    print('%-10s correlate    %8.3f ms  %.3e evals/s' % (name, ms, evals / ms * 1e3))
    ms = timeit(lambda: capi.correlate_f32(wd, rd, ld, want_corr=False), 5)
    print('%-10s corrcoef     %8.3f ms  %.3e evals/s' % (name, ms, evals / ms * 1e3))
    ms = timeit(lambda: capi.maxplus_f32(wd, rd, ld))
    print('%-10s maxplus_f32  %8.3f ms  %.3e evals/s' % (name, ms, evals / ms * 1e3))


if __name__ == '__main__':
  main()
