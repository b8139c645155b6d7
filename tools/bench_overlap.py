"""Experiment: overlap mask_select of chunk c with the max-plus sweep of chunk c+1
on a second stream.  python tools/bench_overlap.py [chunks]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stackrl_b200 import capi, synth

E, R, H, W, h = 4096, 8, 32, 32, 16
NSETS = 8


def main():
  chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
  steps = 200
  dev = torch.device('cuda')
  walls_h, rocks_h, _ = synth.placement_batch(0, E, R, H, W, h)
  goals_h = synth.goals(7, E, H, W)
  sets = []
  for s in range(NSETS):
    g = torch.from_numpy(goals_h).to(dev).clone()
    sets.append(dict(
      walls=torch.roll(torch.from_numpy(walls_h).to(dev), s, 0).contiguous(), goals=g,
      rocks=torch.roll(torch.from_numpy(rocks_h).to(dev), -s, 0).contiguous(),
      level=g.amax(dim=(1, 2)),
      values=torch.empty((E, R, 17, 17), dtype=torch.float32, device=dev),
      actions=torch.empty((E, R), dtype=torch.int64, device=dev),
      best=torch.empty((E, 2), dtype=torch.int64, device=dev)))
  lib, P_ = capi.lib, capi._P
  sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
  pa, pb = P_(sa.cuda_stream), P_(sb.cuda_stream)
  bounds = [(c * E // chunks, (c + 1) * E // chunks) for c in range(chunks)]
  last_sel = [None] * NSETS

  def step(k):
    s = sets[k % NSETS]
    if last_sel[k % NSETS] is not None:
      sa.wait_event(last_sel[k % NSETS])
    for lo, hi in bounds:
      n = hi - lo
      capi._check(lib.srl_maxplus_f32(
        P_(s['walls'][lo:hi].data_ptr()), P_(s['rocks'][lo:hi].data_ptr()),
        P_(s['level'][lo:hi].data_ptr()), P_(s['values'][lo:hi].data_ptr()), n, R, H, W, h,
        0.0, pa))
      ev = torch.cuda.Event()
      ev.record(sa)
      sb.wait_event(ev)
      capi._check(lib.srl_mask_select_f32(
        P_(s['values'][lo:hi].data_ptr()), P_(s['walls'][lo:hi].data_ptr()),
        P_(s['goals'][lo:hi].data_ptr()), P_(s['rocks'][lo:hi].data_ptr()),
        P_(s['actions'][lo:hi].data_ptr()), P_(None), P_(s['best'][lo:hi].data_ptr()),
        n, R, H, W, h, 1, 0.75, pb))
    ev = torch.cuda.Event()
    ev.record(sb)
    last_sel[k % NSETS] = ev

  for k in range(10):
    step(k)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(sa)
  for k in range(steps):
    step(k)
  sa.wait_stream(sb)
  e1.record(sa)
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / steps
  print('chunks %d: %.4f ms/step  %.3e evals/s' % (chunks, ms, E * R * 289 / ms * 1e3))
  # correctness vs single-stream path
  ref = sets[0]
  v2 = capi.maxplus_f32(ref['walls'], ref['rocks'], ref['level'])
  a2, _, b2 = capi.mask_select(v2, ref['walls'], ref['goals'], ref['rocks'], want_shown=False)
  print('match', torch.equal(a2, ref['actions']), torch.equal(b2, ref['best']))


if __name__ == '__main__':
  main()
