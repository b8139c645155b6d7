"""Generate tests/golden/*.npz by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference modules are executed through oracle/refload.py (stub gin / gym /
tensorflow / pybullet).  Inputs are seeded; inputs AND the reference's outputs
are stored so the fixtures are self-contained on the GPU box, where the
reference tree does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refload  # noqa: E402
from stackrl_b200 import synth  # noqa: E402


def scoring_cases():
  """(name, obs) pairs covering the edge cases listed in SURVEY section 4."""
  cases = []
  cases.append(('c2like_f32', synth.observation(0, 32, 32, 16)))
  cases.append(('nonsquare_wall_f32', synth.observation(1, 24, 20, 8)))
  cases.append(('dense_rock_f32', synth.observation(2, 20, 20, 6, zero_fraction=0.)))
  cases.append(('empty_rock_f32', synth.observation(3, 16, 16, 4, zero_fraction=1.)))
  cases.append(('ties_f32', synth.observation(4, 24, 24, 8, quantum=1 / 64)))
  cases.append(('flat_wall_f32', synth.observation(5, 16, 16, 8, flat=True)))
  cases.append(('stackv0_u8', synth.observation(6, 32, 32, 8, dtype='uint8')))
  cases.append(('c4like_f32', synth.observation(7, 64, 64, 16)))
  cases.append(('c5like_f32', synth.observation(8, 128, 128, 32)))
  cases.append(('c5like_u8', synth.observation(9, 128, 128, 32, dtype='uint8')))
  cases.append(('full_rock_window_f32', synth.observation(10, 12, 12, 12)))
  return cases


def make_scoring(ns, path):
  B = ns.baselines
  out = {}
  names = []
  for name, obs in scoring_cases():
    names.append(name)
    out[name + '/wall_goal'] = obs[0]
    out[name + '/rock'] = obs[1]
    small = obs[0].shape[0] <= 64
    out[name + '/height'] = B.height(obs)
    out[name + '/goal_overlap'] = B.goal_overlap(obs)
    out[name + '/goal_overlap_t50'] = B.goal_overlap(obs, threshold=0.5)
    if obs[1].any():
      # (an all-zero rock divides by zero in these; out of contract)
      out[name + '/correlate'] = B.correlate(obs)
      out[name + '/corrcoef'] = B.corrcoef(obs)
      d, dh = B.difference(obs, return_height=True)
      out[name + '/difference'] = d
      out[name + '/difference_height'] = dh
      if small:
        out[name + '/difference_w0'] = B.difference(obs, weights_exponent=0)
        out[name + '/difference_d1'] = B.difference(obs, difference_exponent=1)
        out[name + '/corrcoef_localized'] = B.corrcoef(obs, localized=True)
    # Baseline.call through the real PyGreedy, all selection variants.
    methods = ['height', 'difference'] if obs[1].any() else ['height']
    for method in methods:
      for goal in (True, False):
        for minorder in (0, 1, 2):
          if not goal and minorder != 1:
            continue
          if not small and (method != 'height' or minorder == 2):
            continue
          pol = B.Baseline(method=method, goal=goal, minorder=minorder, value=True)
          a, v = pol(obs)
          key = '{}/select_{}_g{}_m{}'.format(name, method, int(goal), minorder)
          out[key + '/action'] = np.int64(a)
          out[key + '/values'] = v
  out['names'] = np.array(names)

  # Rotation-batched caller (TestStackEnv layout, policies.py:57-91).
  for name, seed, dtype in (('batched_f32', 20, 'float32'), ('batched_u8', 21, 'uint8')):
    obs = synth.batched_observation(seed, 32, 32, 16, 8, dtype=dtype)
    out[name + '/wall_goal'] = obs[0]
    out[name + '/rock'] = obs[1]
    for method in ('height', 'difference'):
      pol = B.Baseline(method=method, value=True, batched=True, batchwise=True)
      (k, idx), v = pol(obs)
      out['{}/{}/k'.format(name, method)] = np.int64(k)
      out['{}/{}/index'.format(name, method)] = np.int64(idx)
      out['{}/{}/values'.format(name, method)] = v
    pol = B.Baseline(method='height', value=True, batched=True, unravel=True)
    a, v = pol(obs)
    out[name + '/height_unravel/actions'] = a
    out[name + '/height_unravel/values'] = v
  np.savez_compressed(path, **out)
  return len(out)


def main():
  ns = refload.load()
  n = make_scoring(ns, os.path.join(HERE, 'scoring.npz'))
  print('scoring.npz: {} arrays'.format(n))
  try:
    from tests.golden import make_golden_observe
  except ImportError:
    return
  make_golden_observe.main(ns)


if __name__ == '__main__':
  main()
