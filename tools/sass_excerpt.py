"""SASS evidence for profiles/: opcode histogram of one kernel of the built library and the
densest window of a given opcode (the inner loop).
python tools/sass_excerpt.py file.cu kernel-regex OPCODE [window]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

cu, kre, opcode = sys.argv[1:4]
window = int(sys.argv[4]) if len(sys.argv) > 4 else 48
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', os.path.basename(cu).replace('.cu', '') + '.sm_100a.cubin',
                os.path.join(root, 'stackrl_b200', 'libstackrl_b200.so')], cwd=tmp,
               capture_output=True)
cubin = os.path.join(tmp, os.listdir(tmp)[0])
dis = subprocess.run(['cuobjdump', '-sass', cubin], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in dis.splitlines():
  m = re.match(r'\s*Function : (\S+)', line)
  if m:
    cur = m.group(1)
    funcs[cur] = []
    continue
  if cur is not None and re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
    funcs[cur].append(re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', line.rstrip()))
dem = {f: subprocess.run(['cu++filt', f], capture_output=True, text=True).stdout.strip()
       for f in funcs}
pick = [f for f in funcs if re.search(kre, dem[f])]
# the instantiation with most of the opcode
def count(f):
  return sum(1 for l in funcs[f] if re.search(r'\b' + re.escape(opcode), l))
best = max(pick, key=count)
lines = funcs[best]
ops = collections.Counter()
for l in lines:
  t = re.sub(r'/\*[0-9a-f]+\*/', '', l).split()
  if t:
    ops[(t[1] if t[0].startswith('@') else t[0]).rstrip(';')] += 1
print('kernel:', dem[best])
print('SASS instructions:', len(lines))
print('opcodes:', ', '.join('%s %d' % kv for kv in ops.most_common(24)))
hit = [1 if re.search(r'\b' + re.escape(opcode), l) else 0 for l in lines]
pre = [0]
for h in hit:
  pre.append(pre[-1] + h)
k = max(range(max(1, len(lines) - window)), key=lambda i: pre[min(len(lines), i + window)] - pre[i])
print('densest %d-instruction window of %s (%d of them):' % (
  window, opcode, pre[min(len(lines), k + window)] - pre[k]))
for l in lines[k:k + window]:
  print(l)
