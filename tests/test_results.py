"""The result writers against the reference's own ``write`` (stackrl/test.py:46-148):
the function's source is taken from the reference file (it only needs os and numpy,
the module around it needs gym / matplotlib / TensorFlow) and run on the same call
sequences; the files must be byte-identical."""
import ast
import os

import numpy as np
import pytest

from oracle import refload
from stackrl_b200 import results


def _reference_write():
  path = os.path.join(refload.REF_ROOT, 'stackrl', 'test.py')
  if not os.path.isfile(path):
    pytest.skip('stackrl/test.py of the reference is not available here')
  tree = ast.parse(open(path).read())
  fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'write'][0]
  ns = {'os': os, 'np': np}
  exec(compile(ast.Module(body=[fn], type_ignores=[]), path, 'exec'), ns)
  return ns['write']


CALLS = [
  dict(keys=['height', 'random'], priority=64, **{'return': [1.5, 0.25], 'return_std': [0.1, 0.2]}),
  dict(keys=['difference'], priority=64, **{'return': [2.0], 'return_std': [0.3]}),
  dict(keys=['height', 'corrcoef'], priority=32, **{'return': [9.0, 0.5], 'return_std': [0.0, 0.1]}),
  dict(keys=['random'], priority=128, **{'return': [0.75], 'return_std': [0.05]}),
]


@pytest.mark.needs_reference
def test_results_csv_equals_the_reference_writer(tmp_path):
  ref_write = _reference_write()
  a, b = str(tmp_path / 'ref' / 'results.csv'), str(tmp_path / 'own' / 'results.csv')
  for call in CALLS:
    ref_write(a, **call)
    results.write(b, **call)
    assert open(a).read() == open(b).read()
  assert open(b).readline() == 'Keys,Priority,Return,ReturnStd\n'
  # a header that does not match: refused, or overwritten with force
  with pytest.raises(ValueError):
    results.write(b, keys=['x'], other=[1])
  with pytest.raises(ValueError):
    ref_write(a, keys=['x'], other=[1])
  results.write(b, force=True, keys=['x'], other=[1])
  ref_write(a, force=True, keys=['x'], other=[1])
  assert open(a).read() == open(b).read() == 'Keys,Other\nx,1\n'


def test_csv_rules_without_the_reference(tmp_path):
  f = str(tmp_path / 'results.csv')
  for call in CALLS:
    results.write(f, **call)
  lines = open(f).read().splitlines()
  assert lines[0] == 'Keys,Priority,Return,ReturnStd'
  rows = {l.split(',')[0]: l.split(',') for l in lines[1:]}
  assert rows['height'][1:] == ['64', '1.5', '0.1']        # priority 32 did not replace it
  assert rows['random'][1:] == ['128', '0.75', '0.05']      # priority 128 did
  assert set(rows) == {'height', 'random', 'difference', 'corrcoef'}


def test_data_npz_layout(tmp_path):
  P, T, shape = 3, 40, (97, 97)
  rng = np.random.default_rng(0)
  flat = rng.integers(0, 97 * 97, (P, T))
  values = rng.standard_normal((P, T, 97 * 97))
  rewards = rng.standard_normal((P, T // P + 1))
  data = results.pack_data(['a', 'b', 'c'], flat, values, rewards, [0, 13, 13, 26], shape)
  path = results.save_data(str(tmp_path / 'run'), **data)
  z = np.load(path)
  assert sorted(z.files) == ['actions', 'episode_bounds', 'keys', 'rewards', 'values']
  assert z['actions'].dtype == np.uint8 and z['actions'].shape == (P, T, 2)
  assert np.array_equal(z['actions'][..., 0].astype(int) * 97 + z['actions'][..., 1], flat)
  assert z['values'].dtype == np.float32 and z['values'].shape == (P, T, 9409)
  assert z['rewards'].dtype == np.float32
  assert z['episode_bounds'].dtype == np.uint16 and list(z['episode_bounds']) == [0, 13, 26, 40]
  big = results.pack_data(['a'], np.zeros((1, 5), int), np.zeros((1, 5, 300 * 300)),
                          np.zeros((1, 5)), [0], (300, 300))
  assert big['actions'].dtype == np.uint16
