import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys
from stackrl_b200 import capi
names = {0: 'FADD+FMNMX', 1: '2FADD+FMNMX3', 2: 'FADD2+FMNMX3', 3: 'FADD', 4: 'FMNMX(fused)',
         5: 'FMNMX3', 6: 'FADD2', 7: 'FADD2+VIMNMX3', 8: 'VIMNMX3', 9: '2FADD+VIMNMX3', 14: 'warp-specialised FADD2 | VIMNMX3', 15: 'BB-split grp8 FADD2 / VIMNMX3', 17: 'VIADDMNMX.S16x2 (2 cells/inst)', 18: 'VIADDMNMX.S32 (1 cell/inst)'}
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
variants = [int(v) for v in sys.argv[2].split(',') if v] if len(sys.argv) > 2 else sorted(names)
for v in variants:
  c = capi.microbench_addmax(v, iters)
  print('variant %d %-16s %.4e cells/s  %.1f cells/clk/SM @1965MHz' % (v, names[v], c, c / 148 / 1.965e9))
if len(sys.argv) <= 2 or 'fma' in sys.argv:
  for v, name in ((0, 'FFMA'), (1, 'FFMA2 shared multiplicand'), (2, 'FFMA2 distinct operands')):
    f = capi.microbench_fma(v, iters)
    print('fma variant %d %-26s %.4e FMA/s = %.1f TFLOP/s  %.1f FMA/clk/SM @1965MHz' % (
      v, name, f, 2 * f / 1e12, f / 148 / 1.965e9))
