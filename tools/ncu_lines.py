"""Per-source-line totals of an .ncu-rep source page: the SASS rows of the report are
matched, in order, with `nvdisasm -g` of the same kernel in the built library (the report's
CSV carries no line column).
python tools/ncu_lines.py report.ncu-rep kernel-regex file.cu [top] [launch index]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kre, cu = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name',
                      'regex:' + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# one section per captured launch: a 2-column "Kernel Name" row, the header, the SASS rows
starts = [k for k, r in enumerate(rows) if r and r[0] == 'Kernel Name'] or [0]
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
lo = starts[which]
rows = rows[lo:(starts[which + 1] if which + 1 < len(starts) else len(rows))]
name = rows[0][1]
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
sass = [(r[ix['Source']].strip(), int(r[ix['Instructions Executed']]), int(r[ix['# Samples']]))
        for r in rows[2:] if len(r) >= len(h)]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', os.path.basename(cu).replace('.cu', '') + '.sm_100a.cubin',
                os.path.join(root, 'stackrl_b200', 'libstackrl_b200.so')], cwd=tmp,
               capture_output=True)
cubin = os.path.join(tmp, os.listdir(tmp)[0])
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
# split into functions; pick the one whose demangled name matches best
funcs, cur = {}, None
for line in dis.splitlines():
  m = re.match(r'\s*\.text\.(\S+):', line)
  if m:
    cur = m.group(1)
    funcs[cur] = []
    continue
  if cur is not None:
    funcs[cur].append(line)
dem = {f: subprocess.run(['cu++filt', f], capture_output=True, text=True).stdout.strip()
       for f in funcs}
want = re.sub(r'\s+', '', name)
best = max(funcs, key=lambda f: len(os.path.commonprefix([re.sub(r'\s+', '', dem[f]), want])))
print(name[:100])
print('matched', dem[best][:100])
line_of, cur_line = [], 0
for line in funcs[best]:
  m = re.search(r'//## File "([^"]+)", line (\d+)', line)
  if m:
    cur_line = int(m.group(2)) if m.group(1).endswith(os.path.basename(cu)) else -int(m.group(2))
    continue
  if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
    line_of.append(cur_line)
if len(line_of) != len(sass):
  print('warning: %d SASS rows in the report, %d in the disassembly' % (len(sass), len(line_of)))
tot_n = sum(n for _, n, _ in sass)
tot_s = sum(s for _, _, s in sass)
agg = collections.defaultdict(lambda: [0, 0])
for (src, n, s), ln in zip(sass, line_of):
  agg[ln][0] += n
  agg[ln][1] += s
src_lines = open(os.path.join(root, 'stackrl_b200', 'csrc', os.path.basename(cu))).read().splitlines()
print('warp instructions %d, samples %d' % (tot_n, tot_s))
for ln, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
  text = src_lines[ln - 1].strip()[:70] if 0 < ln <= len(src_lines) else '(other file, line %d)' % -ln
  print('%5d  inst %5.1f%%  samples %5.1f%%  %s' % (ln, 100. * n / tot_n, 100. * s / max(tot_s, 1), text))
