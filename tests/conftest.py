import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')
  config.addinivalue_line(
    'markers', 'needs_reference: needs /root/reference (build container only)')


def pytest_collection_modifyitems(config, items):
  from oracle import refload
  if refload.available():
    return
  skip = pytest.mark.skip(reason='/root/reference not present on this machine')
  for item in items:
    if 'needs_reference' in item.keywords:
      item.add_marker(skip)


class Golden(object):
  """Read-only view on a tests/golden/*.npz fixture with '/'-separated keys."""
  def __init__(self, name):
    self._z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
  def __getitem__(self, key):
    return self._z[key]
  def __contains__(self, key):
    return key in self._z.files
  def keys(self, prefix=''):
    return [k for k in self._z.files if k.startswith(prefix)]
  def obs(self, case):
    return (self._z[case + '/wall_goal'], self._z[case + '/rock'])


@pytest.fixture(scope='session')
def scoring_golden():
  return Golden('scoring.npz')


@pytest.fixture(scope='session')
def observe_golden():
  return Golden('observe.npz')


# Reference default geometry, used by the episode fixtures.
GEOM = dict(H=128, W=128, h=32, pixel=0.125 / 32, max_z=0.375,
            object_max_dimension=0.125, object_z=0.125, goal_z=0.25)


def split_depths(golden, key):
  """The depth images recorded since the previous step, in call order."""
  flat = golden[key + '/depths']
  out, at = [], 0
  for rows, cols in golden[key + '/depth_shapes']:
    out.append(flat[at:at + rows * cols].reshape(rows, cols))
    at += rows * cols
  return out
