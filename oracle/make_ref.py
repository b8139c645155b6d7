"""Stage the reference's own hot-path files under oracle/_ref/ (never committed).

Test / baseline infrastructure (see oracle/__init__.py).  ``/root/reference`` only
exists in the build container; the GPU box gets whatever lies in the repo snapshot.
``oracle/_ref/`` is git-ignored (reference sources never enter the history) but
NOT gpurun-ignored, so a byte-for-byte copy of the files below travels with the
snapshot and ``oracle/refload.py`` can execute the UNMODIFIED reference there too:
bench.py's ``--impl reference`` arm and ``cpu_baseline`` then time the reference's
own code (``kind: "reference"``), not a port.

    python -m oracle.make_ref            # run by __graft_entry__.build()

Staged (relative to the reference root):
  stackrl/baselines.py, stackrl/agents/policies.py,
  stackrl/envs/stack/{observer,simulator,rewarder,env}.py, stackrl/envs/data/__init__.py,
  stackrl/envs/data/template.urdf and a 64-rock sample of the registered
  environments' asset glob '[5-9]?_*.urdf' (Stack-v0, envs/stack/__init__.py:3-8)
  with their .obj meshes, for the C1 episode run.
"""
import filecmp
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC = os.environ.get('STACKRL_REFERENCE_SOURCE', '/root/reference')

FILES = [
  'stackrl/baselines.py',
  'stackrl/agents/policies.py',
  'stackrl/envs/stack/observer.py',
  'stackrl/envs/stack/simulator.py',
  'stackrl/envs/stack/rewarder.py',
  'stackrl/envs/stack/env.py',
  'stackrl/envs/stack/__init__.py',
  'stackrl/envs/data/__init__.py',
  'stackrl/envs/data/template.urdf',
  'LICENSE',
]
ROCKS_PER_DECADE = 13      # of each irregularity class 50, 55, ... 95 -> 5 x 13 - 1 = 64


def rock_sample():
  names = []
  for irregularity in (50, 60, 70, 80, 90):
    found = sorted(glob.glob(os.path.join(
      SRC, 'stackrl/envs/data/generated', '{}_*.urdf'.format(irregularity))))
    names += found[:ROCKS_PER_DECADE]
  return names[:64]


def stage(verbose=False):
  """Copy the files if the source tree is here; returns the staging root or None."""
  if not os.path.isfile(os.path.join(SRC, 'stackrl', 'baselines.py')):
    return DST if os.path.isfile(os.path.join(DST, 'stackrl', 'baselines.py')) else None
  pairs = [(os.path.join(SRC, f), os.path.join(DST, f)) for f in FILES
           if os.path.exists(os.path.join(SRC, f))]
  for urdf in rock_sample():
    rel = os.path.relpath(urdf, SRC)
    pairs.append((urdf, os.path.join(DST, rel)))
    pairs.append((urdf[:-5] + '.obj', os.path.join(DST, rel[:-5] + '.obj')))
  for src, dst in pairs:
    if os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False):
      continue
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    if verbose:
      print('staged', os.path.relpath(dst, DST))
  return DST


if __name__ == '__main__':
  print(stage(verbose='-v' in sys.argv))
