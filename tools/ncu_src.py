"""Summarise the source page of an .ncu-rep: stall totals, opcode mix, hottest SASS lines.
python tools/ncu_src.py report.ncu-rep [kernel-regex]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
cmd = ['ncu', '-i', rep, '--page', 'source', '--csv']
if len(sys.argv) > 2:
  cmd += ['--kernel-name', 'regex:' + sys.argv[2]]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# one section per captured launch (a "Kernel Name" row, the header, the SASS rows): the first
starts = [k for k, r in enumerate(rows) if r and r[0] == 'Kernel Name'] or [0]
rows = rows[starts[0]:(starts[1] if len(starts) > 1 else len(rows))]
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
L, tot, ops, sm = [], collections.Counter(), collections.Counter(), collections.Counter()
ninst = ns = 0
for k, r in enumerate(rows[2:]):
  if len(r) < len(h):
    continue
  n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
  src = r[ix['Source']].strip()
  t = src.split()
  op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
  ops[op] += n; sm[op] += s
  ninst += n; ns += s
  L.append((k, src, n, s, r[ix['Avg. Threads Executed']]))
  for c in h:
    if c.startswith('stall_') and 'Not Issued' not in c:
      tot[c] += int(r[ix[c]] or 0)
print(rows[0][1][:90])
print('warp instructions', ninst, 'samples', ns)
print('stalls:', ', '.join('%s %.1f%%' % (a[6:], 100 * b / ns) for a, b in tot.most_common(9)))
print('opcodes:', ', '.join('%s %.1f%%(%.1f%%s)' % (a, 100 * b / ninst, 100 * sm[a] / ns) for a, b in ops.most_common(14)))
for k, src, n, s, thr in sorted(L, key=lambda x: -x[3])[:14]:
  print('%5d %-64s n=%d samples=%d thr=%s' % (k, src[:64], n, s, thr))
