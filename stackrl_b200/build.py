"""In-tree build of libstackrl_b200.so (nvcc, sm_100a only).

    python -m stackrl_b200.build [--force]

The shared library lands next to this file so it travels with the repo
snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libstackrl_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

FLAGS = [
  '-gencode', 'arch=compute_100a,code=sm_100a',
  '-O3', '-lineinfo', '-std=c++17',
  '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
  '--fmad=false',          # no silent FMA contraction anywhere: numpy has none
  '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
]


def sources():
  return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def stale():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + \
    glob.glob(os.path.join(CSRC, '*.cuh')) + \
    glob.glob(os.path.join(ROOT, 'include', '*.h')) + [os.path.abspath(__file__)]
  return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
  if not force and not stale():
    return LIB
  objdir = os.path.join(HERE, 'build')
  os.makedirs(objdir, exist_ok=True)
  objs = []
  procs = []
  for src in sources():
    obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
    objs.append(obj)
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
      ['-c', src, '-o', obj]
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                        stderr=subprocess.STDOUT)))
  failed = False
  for src, p in procs:
    out = p.communicate()[0].decode()
    if p.returncode != 0:
      failed = True
    if out.strip() and (verbose or p.returncode != 0):
      sys.stderr.write(out)
  if failed:
    raise RuntimeError('nvcc failed')
  subprocess.check_call(
    [NVCC, '-shared', '-o', LIB] + objs,   # cudart linked statically (nvcc default)
  )
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
