"""Build the oracle's C restatement into oracle/_build/liboracle.so (gcc).

Test infrastructure: building the checker is not using it.  -ffp-contract=off
keeps every float op a separately rounded IEEE op, the property the parity
tests rely on."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), '_build')
LIB = os.path.join(OUT_DIR, 'liboracle.so')
SRC = os.path.join(HERE, 'oracle.c')


def build(force=False):
  os.makedirs(OUT_DIR, exist_ok=True)
  if not force and os.path.exists(LIB) and \
      os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(__file__)):
    return LIB
  subprocess.check_call([
    os.environ.get('CC', 'gcc'), '-O2', '-std=c99', '-fPIC', '-shared',
    '-ffp-contract=off', '-fno-fast-math', '-o', LIB, SRC, '-lm'])
  return LIB


if __name__ == '__main__':
  print(build(force=True))
