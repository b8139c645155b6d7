"""Run the C2 max-plus kernel a few times (ncu target).  python tools/run_maxplus.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stackrl_b200 import capi, synth

E, R, H, W, h = 4096, 8, 32, 32, 16
if len(sys.argv) > 2:
  E, R, H, W, h = (int(x) for x in sys.argv[2].split(','))
walls, rocks, level = synth.placement_batch(0, E, R, H, W, h)
QLOG2 = None
if len(sys.argv) > 3:          # quantise like the rasteriser does and pass the hint
  import numpy as np
  QLOG2 = int(sys.argv[3])
  q = np.float32(2.0 ** QLOG2)
  walls = (np.round(walls / q) * q).astype('float32')
  rocks = (np.round(rocks / q) * q).astype('float32')
dev = torch.device('cuda')
wd, rd, ld = (torch.from_numpy(x).to(dev) for x in (walls, rocks, level))
out = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(3):
  capi.maxplus_f32(wd, rd, ld, out=out, quantum_log2=QLOG2)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
  capi.maxplus_f32(wd, rd, ld, out=out, quantum_log2=QLOG2)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print('%d envs: %.4f ms  %.3e evals/s' % (E, ms, out.numel() / ms * 1e3))
