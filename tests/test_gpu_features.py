"""uint8 (float64) max-plus, `difference`, and the remaining drop-in functions
of the baselines mirror against the reference's golden outputs.  Bit-exact."""
import numpy as np
import pytest
import torch

from oracle import scoring_np as S
from stackrl_b200 import synth

pytestmark = pytest.mark.gpu

F32_CASES = ['c2like_f32', 'nonsquare_wall_f32', 'dense_rock_f32', 'ties_f32',
             'flat_wall_f32', 'c4like_f32', 'c5like_f32', 'full_rock_window_f32']


@pytest.fixture(scope='module')
def B():
  from stackrl_b200 import baselines
  return baselines


@pytest.mark.parametrize('case', ['stackv0_u8', 'c5like_u8'])
def test_height_uint8_is_float64_exact(B, scoring_golden, case):
  got = B.height(scoring_golden.obs(case))
  want = scoring_golden[case + '/height']
  assert got.dtype == np.float64 and np.array_equal(got, want)


@pytest.mark.parametrize('case', ['stackv0_u8', 'c5like_u8'])
def test_baseline_uint8_actions(B, scoring_golden, case):
  obs = scoring_golden.obs(case)
  keys = sorted({k.rsplit('/', 1)[0] for k in scoring_golden.keys(case + '/select_height')})
  for key in keys:
    _, _, g, m = key.split('/')[1].split('_')
    pol = B.Baseline(method='height', goal=g == 'g1', minorder=int(m[1:]), value=True)
    a, v = pol(obs)
    assert a == int(scoring_golden[key + '/action']), key
    assert np.array_equal(v, scoring_golden[key + '/values']), key


def test_batched_uint8_batchwise(B, scoring_golden):
  obs = scoring_golden.obs('batched_u8')
  pol = B.Baseline(method='height', value=True, batched=True, batchwise=True)
  (k, idx), v = pol(obs)
  assert k == int(scoring_golden['batched_u8/height/k'])
  assert idx == int(scoring_golden['batched_u8/height/index'])
  assert np.array_equal(v, scoring_golden['batched_u8/height/values'])


def test_uint8_batch_matches_oracle():
  from stackrl_b200 import capi
  E, R, H, W, h = 5, 3, 32, 32, 8
  rng = np.random.default_rng(3)
  walls = rng.integers(0, 200, (E, H, W), dtype=np.uint8)
  rocks = rng.integers(0, 86, (E, R, h, h), dtype=np.uint8)
  rocks[rng.random(rocks.shape) < 0.3] = 0
  level = rng.integers(100, 255, (E,), dtype=np.uint8)
  dev = torch.device('cuda')
  got = capi.maxplus_u8(torch.from_numpy(walls).to(dev), torch.from_numpy(rocks).to(dev),
                        torch.from_numpy(level).to(dev)).cpu().numpy()
  for e in range(E):
    goal = np.full((H, W), level[e], dtype=np.uint8)
    for r in range(R):
      want = S.height((np.stack([walls[e], goal], -1), rocks[e, r][..., None]))
      assert np.array_equal(got[e, r], want)


@pytest.mark.parametrize('case', F32_CASES)
def test_difference_bit_exact(B, scoring_golden, case):
  obs = scoring_golden.obs(case)
  d, h0 = B.difference(obs, return_height=True)
  assert d.dtype == np.float64
  assert np.array_equal(h0, scoring_golden[case + '/difference_height'])
  assert np.array_equal(d, scoring_golden[case + '/difference'])
  if case + '/difference_w0' in scoring_golden:
    assert np.array_equal(B.difference(obs, weights_exponent=0),
                          scoring_golden[case + '/difference_w0'])
    assert np.array_equal(B.difference(obs, difference_exponent=1),
                          scoring_golden[case + '/difference_d1'])


@pytest.mark.parametrize('case', ['stackv0_u8', 'c5like_u8'])
def test_difference_uint8_is_float64_exact(B, scoring_golden, case):
  """uint8 observations: numpy runs baselines.py:64-69 in float64 throughout."""
  obs = scoring_golden.obs(case)
  d, h0 = B.difference(obs, return_height=True)
  assert d.dtype == np.float64 and h0.dtype == np.float64
  assert np.array_equal(h0, scoring_golden[case + '/difference_height'])
  assert np.array_equal(d, scoring_golden[case + '/difference'])
  if case + '/difference_w0' in scoring_golden:
    assert np.array_equal(B.difference(obs, weights_exponent=0),
                          scoring_golden[case + '/difference_w0'])
    assert np.array_equal(B.difference(obs, difference_exponent=1),
                          scoring_golden[case + '/difference_d1'])


@pytest.mark.parametrize('case', ['c2like_f32', 'nonsquare_wall_f32', 'dense_rock_f32',
                                  'ties_f32', 'flat_wall_f32', 'c4like_f32',
                                  'full_rock_window_f32', 'stackv0_u8'])
def test_corrcoef_localized_bit_exact(B, scoring_golden, case):
  """The masked correlation coefficient (baselines.py:87-114): numpy pairwise
  sums in float32 (float32 observations) or float64 (uint8 ones)."""
  want = scoring_golden[case + '/corrcoef_localized']
  got = B.corrcoef(scoring_golden.obs(case), localized=True)
  assert got.dtype == want.dtype and got.shape == want.shape
  assert np.array_equal(got, want)


@pytest.mark.parametrize('side', [3, 5, 6, 10, 12, 20])
def test_difference_pairwise_order_odd_sizes(B, side):
  """Rock sizes whose h*h is not a power of two exercise every branch of
  numpy's pairwise summation (n < 8, remainder loop, uneven tree split)."""
  obs = synth.observation(50 + side, 3 * side + 1, 2 * side + 3, side)
  assert np.array_equal(B.difference(obs), S.difference(obs))
  assert np.array_equal(B.difference(obs, weights_exponent=3),
                        S.difference(obs, weights_exponent=3))


def test_difference_rejects_unreproducible_exponent(B):
  obs = synth.observation(1, 16, 16, 4)
  with pytest.raises(ValueError):
    B.difference(obs, difference_exponent=3)


def test_baseline_difference_actions(B, scoring_golden):
  for case in ('c2like_f32', 'ties_f32', 'c4like_f32'):
    obs = scoring_golden.obs(case)
    for minorder in (0, 1, 2):
      key = '{}/select_difference_g1_m{}'.format(case, minorder)
      a, v = B.Baseline(method='difference', minorder=minorder, value=True)(obs)
      assert a == int(scoring_golden[key + '/action'])
      assert np.array_equal(v, scoring_golden[key + '/values'])
  obs = scoring_golden.obs('batched_f32')
  (k, idx), v = B.Baseline(method='difference', value=True, batched=True,
                           batchwise=True)(obs)
  assert k == int(scoring_golden['batched_f32/difference/k'])
  assert idx == int(scoring_golden['batched_f32/difference/index'])
  assert np.array_equal(v, scoring_golden['batched_f32/difference/values'])


def test_random_method_is_numpy_stream(B):
  obs = synth.observation(2, 16, 16, 4)
  assert np.array_equal(B.random(obs, seed=5), S.random(obs, seed=5))


@pytest.mark.parametrize('case', ['c2like_f32', 'nonsquare_wall_f32', 'dense_rock_f32',
                                  'ties_f32', 'c4like_f32', 'c5like_f32'])
def test_correlate_and_corrcoef_within_tolerance(B, scoring_golden, case):
  """correlate / corrcoef are defined by scipy / OpenCV library summation order
  (not part of the reference tree): tolerance match, stated here -- 1e-5 relative
  for correlate, 2e-5 absolute for the correlation coefficient."""
  obs = scoring_golden.obs(case)
  want = scoring_golden[case + '/correlate']
  got = B.correlate(obs)
  assert got.dtype == want.dtype and got.shape == want.shape
  np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
  want = scoring_golden[case + '/corrcoef']
  got = B.corrcoef(obs)
  assert got.dtype == want.dtype and got.shape == want.shape
  np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
  assert set(B.methods) == {'random', 'correlate', 'height', 'difference', 'corrcoef'}


def test_baseline_difference_uint8_actions(B, scoring_golden):
  obs = scoring_golden.obs('stackv0_u8')
  for goal, minorder in ((1, 0), (1, 1), (1, 2), (0, 1)):
    key = 'stackv0_u8/select_difference_g{}_m{}'.format(goal, minorder)
    a, v = B.Baseline(method='difference', goal=bool(goal), minorder=minorder, value=True)(obs)
    assert a == int(scoring_golden[key + '/action'])
    assert np.array_equal(v, scoring_golden[key + '/values'])


@pytest.mark.parametrize('shape', [(40, 8, 32, 32, 16), (6, 2, 128, 128, 32), (9, 3, 40, 36, 8),
                                   (5, 1, 30, 27, 7)])
def test_uint8_integer_key_kernel_equals_float64_kernel(monkeypatch, shape):
  """maxplus_u8: the VIADDMNMX integer-key sweep (default) against the float64
  DADD kernel, for goal levels whose quotients round in every possible way."""
  from stackrl_b200 import capi
  E, R, H, W, h = shape
  rng = np.random.default_rng(17)
  walls = rng.integers(0, 256, (E, H, W), dtype=np.uint8)
  rocks = rng.integers(0, 256, (E, R, h, h), dtype=np.uint8)
  rocks[rng.random(rocks.shape) < 0.3] = 0
  rocks[0] = 0                                      # an all-masked environment
  rocks[1] = np.maximum(rocks[1], 1)                # no masked cell at all
  levels = np.array([1, 2, 3, 7, 85, 170, 255, 254, 129, 128], dtype=np.uint8)
  level = levels[np.arange(E) % len(levels)]
  dev = torch.device('cuda')
  args = [torch.from_numpy(x).to(dev) for x in (walls, rocks, level)]
  monkeypatch.setenv('SRL_U8_MODE', '0')
  want = capi.maxplus_u8(*args)
  monkeypatch.setenv('SRL_U8_MODE', '1')
  got = capi.maxplus_u8(*args)
  assert torch.equal(got, want)
  # and the float64 kernel itself against the oracle on one environment per level
  for e in range(min(E, len(levels))):
    goal = np.full((H, W), level[e], dtype=np.uint8)
    ref = S.height((np.stack([walls[e], goal], -1), rocks[e, 0][..., None]))
    assert np.array_equal(got[e, 0].cpu().numpy(), ref)
