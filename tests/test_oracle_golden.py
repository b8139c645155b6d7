"""The numpy oracle restatement against the reference's own outputs
(tests/golden/scoring.npz, produced by tests/golden/make_golden.py running the
unmodified /root/reference code).  CPU only."""
import numpy as np
import pytest

from oracle import scoring_np as S

CASES = ['c2like_f32', 'nonsquare_wall_f32', 'dense_rock_f32', 'empty_rock_f32',
         'ties_f32', 'flat_wall_f32', 'stackv0_u8', 'c4like_f32', 'c5like_f32',
         'c5like_u8', 'full_rock_window_f32']


def test_case_list_matches_fixture(scoring_golden):
  assert list(scoring_golden['names']) == CASES


@pytest.mark.parametrize('case', CASES)
def test_height_bit_exact(scoring_golden, case):
  obs = scoring_golden.obs(case)
  want = scoring_golden[case + '/height']
  got = S.height(obs)
  assert got.dtype == want.dtype and np.array_equal(got, want)


@pytest.mark.parametrize('case', ['c2like_f32', 'ties_f32', 'stackv0_u8'])
def test_height_loop_form_bit_exact(scoring_golden, case):
  obs = scoring_golden.obs(case)
  assert np.array_equal(S.height_loop(obs), scoring_golden[case + '/height'])


@pytest.mark.parametrize('case', CASES)
def test_goal_overlap_exact(scoring_golden, case):
  obs = scoring_golden.obs(case)
  assert np.array_equal(S.goal_overlap(obs), scoring_golden[case + '/goal_overlap'])
  assert np.array_equal(S.goal_overlap(obs, threshold=0.5),
                        scoring_golden[case + '/goal_overlap_t50'])


@pytest.mark.parametrize('case', [c for c in CASES if c != 'empty_rock_f32'])
def test_difference_bit_exact(scoring_golden, case):
  obs = scoring_golden.obs(case)
  d, h0 = S.difference(obs, return_height=True)
  assert np.array_equal(d, scoring_golden[case + '/difference'])
  assert np.array_equal(h0, scoring_golden[case + '/difference_height'])
  if case + '/difference_w0' in scoring_golden:
    assert np.array_equal(S.difference(obs, weights_exponent=0),
                          scoring_golden[case + '/difference_w0'])
    assert np.array_equal(S.difference(obs, difference_exponent=1),
                          scoring_golden[case + '/difference_d1'])


@pytest.mark.parametrize('case', [c for c in CASES if c != 'empty_rock_f32'])
def test_correlations(scoring_golden, case):
  obs = scoring_golden.obs(case)
  assert np.array_equal(S.correlate(obs), scoring_golden[case + '/correlate'])
  assert np.array_equal(S.corrcoef(obs), scoring_golden[case + '/corrcoef'])
  if case + '/corrcoef_localized' in scoring_golden:
    assert np.array_equal(S.corrcoef(obs, localized=True),
                          scoring_golden[case + '/corrcoef_localized'])


def _select_keys(golden, case):
  return sorted({k.rsplit('/', 1)[0] for k in golden.keys(case + '/select_')})


@pytest.mark.parametrize('case', CASES)
def test_select_matches_baseline_call(scoring_golden, case):
  obs = scoring_golden.obs(case)
  keys = _select_keys(scoring_golden, case)
  assert keys
  for key in keys:
    _, method, g, m = key.split('/')[1].split('_')
    a, v = S.baseline_call(obs, method=method, goal=g == 'g1', minorder=int(m[1:]))
    assert a == int(scoring_golden[key + '/action']), key
    assert np.array_equal(v.ravel(), scoring_golden[key + '/values']), key


@pytest.mark.parametrize('name', ['batched_f32', 'batched_u8'])
def test_batched_batchwise(scoring_golden, name):
  obs = scoring_golden.obs(name)
  for method in ('height', 'difference'):
    (k, idx), v = S.greedy(
      obs, lambda o: S.baseline_call(o, method=method), value=True,
      batched=True, batchwise=True)
    assert k == int(scoring_golden['{}/{}/k'.format(name, method)])
    assert idx == int(scoring_golden['{}/{}/index'.format(name, method)])
    assert np.array_equal(v, scoring_golden['{}/{}/values'.format(name, method)])
  a, v = S.greedy(obs, lambda o: S.baseline_call(o, method='height'),
                  value=True, batched=True, unravel=True)
  assert np.array_equal(a, scoring_golden[name + '/height_unravel/actions'])
  assert np.array_equal(v, scoring_golden[name + '/height_unravel/values'])
