"""goal_overlap / Baseline.call / PyGreedy conventions on the GPU against the
reference's golden outputs and the oracle.  Indices and masks are exact; the
float64 value maps are bit-exact (they are negations / one float64 add of
bit-exact float32 scores)."""
import numpy as np
import pytest
import torch

from oracle import scoring_np as S
from stackrl_b200 import synth

pytestmark = pytest.mark.gpu

F32_CASES = ['c2like_f32', 'nonsquare_wall_f32', 'dense_rock_f32', 'empty_rock_f32',
             'ties_f32', 'flat_wall_f32', 'c4like_f32', 'c5like_f32',
             'full_rock_window_f32']
ALL_CASES = F32_CASES + ['stackv0_u8', 'c5like_u8']


@pytest.fixture(scope='module')
def B():
  from stackrl_b200 import baselines
  return baselines


@pytest.mark.parametrize('case', ALL_CASES)
def test_goal_overlap_matches_reference(B, scoring_golden, case):
  obs = scoring_golden.obs(case)
  assert np.array_equal(B.goal_overlap(obs), scoring_golden[case + '/goal_overlap'])
  assert np.array_equal(B.goal_overlap(obs, threshold=0.5),
                        scoring_golden[case + '/goal_overlap_t50'])


@pytest.mark.parametrize('case', F32_CASES)
def test_baseline_height_matches_reference(B, scoring_golden, case):
  obs = scoring_golden.obs(case)
  keys = sorted({k.rsplit('/', 1)[0] for k in scoring_golden.keys(case + '/select_height')})
  assert keys
  for key in keys:
    _, _, g, m = key.split('/')[1].split('_')
    pol = B.Baseline(method='height', goal=g == 'g1', minorder=int(m[1:]), value=True)
    a, v = pol(obs)
    assert a == int(scoring_golden[key + '/action']), key
    want = scoring_golden[key + '/values']
    assert v.dtype == want.dtype and np.array_equal(v, want), key


@pytest.mark.parametrize('case', ['c2like_f32', 'ties_f32', 'c4like_f32'])
def test_baseline_with_user_callable(B, scoring_golden, case):
  """method=<callable> (seam b3): the map comes from the callable (here the
  oracle's `difference`), the selection runs on the GPU."""
  obs = scoring_golden.obs(case)
  for minorder in (0, 1, 2):
    key = '{}/select_difference_g1_m{}'.format(case, minorder)
    pol = B.Baseline(method=S.difference, minorder=minorder, value=True)
    a, v = pol(obs)
    assert a == int(scoring_golden[key + '/action'])
    assert np.array_equal(v, scoring_golden[key + '/values'])


def test_batched_batchwise_matches_reference(B, scoring_golden):
  obs = scoring_golden.obs('batched_f32')
  pol = B.Baseline(method='height', value=True, batched=True, batchwise=True)
  (k, idx), v = pol(obs)
  assert k == int(scoring_golden['batched_f32/height/k'])
  assert idx == int(scoring_golden['batched_f32/height/index'])
  assert np.array_equal(v, scoring_golden['batched_f32/height/values'])
  pol = B.Baseline(method='height', value=True, batched=True, unravel=True)
  a, v = pol(obs)
  assert np.array_equal(a, scoring_golden['batched_f32/height_unravel/actions'])
  assert np.array_equal(v, scoring_golden['batched_f32/height_unravel/values'])


def test_unbatched_return_conventions(B, scoring_golden):
  obs = scoring_golden.obs('c2like_f32')
  want = int(scoring_golden['c2like_f32/select_height_g1_m1/action'])
  assert B.Baseline(method='height')(obs) == want
  ij = B.Baseline(method='height', unravel=True)(obs)
  assert tuple(ij) == tuple(np.unravel_index(want, (17, 17)))
  with pytest.raises(ValueError):
    B.Baseline(method='nope')
  with pytest.raises(TypeError):
    B.Baseline(method=3)


@pytest.mark.parametrize('shape', [(6, 8, 32, 32, 16), (3, 4, 64, 64, 16), (2, 3, 40, 56, 8),
                                   (2, 6, 32, 32, 16), (1, 12, 24, 24, 8), (1, 36, 48, 48, 16)])
@pytest.mark.parametrize('goal,minorder', [(True, 1), (True, 0), (True, 2), (False, 1)])
def test_placement_scorer_batched(B, shape, goal, minorder):
  """Device-resident batch: actions / batch-wise picks equal to looping the
  oracle's Baseline.call + PyGreedy over every environment and view."""
  E, R, H, W, h = shape
  walls, rocks, _ = synth.placement_batch(31, E, R, H, W, h)
  goals = synth.goals(32, E, H, W)
  dev = torch.device('cuda')
  scorer = B.PlacementScorer(goal=goal, minorder=minorder)
  out = scorer(torch.from_numpy(walls).to(dev), torch.from_numpy(goals).to(dev),
               torch.from_numpy(rocks).to(dev), want_shown=True)
  actions = out['actions'].cpu().numpy()
  best = out['best'].cpu().numpy()
  shown = out['shown'].cpu().numpy()
  for e in range(E):
    wg = np.stack([walls[e], goals[e]], -1)
    obs = (np.stack([wg] * R), rocks[e][..., None])
    (k, idx), v = S.greedy(
      obs, lambda o: S.baseline_call(o, method='height', goal=goal, minorder=minorder),
      value=True, batched=True, batchwise=True)
    per_view = [S.baseline_call((wg, rocks[e, r][..., None]), goal=goal,
                                minorder=minorder)[0] for r in range(R)]
    assert list(actions[e]) == per_view
    assert (best[e, 0], best[e, 1]) == (k, idx)
    assert np.array_equal(shown[e].reshape(R, -1), v)


@pytest.mark.parametrize('shape', [(37, 64, 64, 16), (5, 32, 48, 16), (9, 44, 36, 16),
                                   (3, 96, 80, 16)])
@pytest.mark.parametrize('minorder', [0, 1])
@pytest.mark.parametrize('dtype', ['float32', 'uint8'])
def test_single_view_maps_off_a_16_byte_boundary(B, shape, minorder, dtype):
  """R == 1 with odd P: the score map of three environments in four starts 4, 8 or 12
  bytes off a 16-byte boundary (staged at the source's phase): actions and value maps
  equal the oracle's for environments of every phase."""
  from stackrl_b200 import capi
  E, H, W, h = shape
  walls, rocks, _ = synth.placement_batch(41, E, 1, H, W, h)
  goals = synth.goals(42, E, H, W)
  dev = torch.device('cuda')
  if dtype == 'uint8':
    walls, goals, rocks = (synth.to_dtype(x, 'uint8') for x in (walls, goals, rocks))
  wd, gd, rd = (torch.from_numpy(x).to(dev) for x in (walls, goals, rocks))
  run = capi.maxplus_u8 if dtype == 'uint8' else capi.maxplus_f32
  values = run(wd, rd, capi.goal_level(gd))
  got = capi.mask_select(values, wd, gd, rd, minorder=minorder)
  for e in list(range(min(E, 4))) + [E - 1]:
    obs = (np.stack([walls[e], goals[e]], -1), rocks[e, 0][..., None])
    a, v = S.baseline_call(obs, goal=True, minorder=minorder)
    assert int(got[0][e, 0]) == a
    assert np.array_equal(got[1][e, 0].cpu().numpy().ravel(), np.asarray(v).ravel())


def test_select_first_index_ties():
  """All-equal scores: np.argmin's first index, on a goal mask whose first
  member is not position 0."""
  from stackrl_b200 import capi
  dev = torch.device('cuda')
  values = torch.full((2, 1, 5, 7), 0.5, dtype=torch.float32, device=dev)
  counts = torch.zeros((2, 1, 5, 7), dtype=torch.int32, device=dev)
  counts[0, 0, 2, 3:] = 4
  counts[0, 0, 3, :] = 4
  counts[1] = 1
  actions, shown, best = capi.select(values, counts, minorder=1)
  # env 0: no interior local minimum inside the mask's first row, so the
  # first masked cell wins; env 1: border cells are never minima (zero
  # padding, quirk Q6), the first interior cell (1, 1) is.
  for e, expect in ((0, None), (1, 1 * 7 + 1)):
    v = values[e, 0].cpu().numpy().astype('float64')
    c = counts[e, 0].cpu().numpy()
    a, s = S.select(v, c >= 0.75 * c.max(), 1)
    assert actions[e, 0].item() == a
    if expect is not None:
      assert a == expect
    assert np.array_equal(shown[e, 0].cpu().numpy(), s)


def test_drop_height_matches_pose_formula():
  """Observer.pose's z (observer.py:401-409) at picked positions."""
  from stackrl_b200 import capi
  E, R, H, W, h = 9, 4, 32, 32, 16
  walls, rocks, _ = synth.placement_batch(41, E, R, H, W, h)
  rng = np.random.default_rng(0)
  picks = np.stack([rng.integers(0, R, E), rng.integers(0, H - h + 1, E),
                    rng.integers(0, W - h + 1, E)], -1).astype('int32')
  dev = torch.device('cuda')
  got = capi.drop_height_f32(torch.from_numpy(walls).to(dev),
                             torch.from_numpy(rocks).to(dev),
                             torch.from_numpy(picks).to(dev)).cpu().numpy()
  for e in range(E):
    r, i, j = picks[e]
    assert got[e] == S.drop_height(walls[e], rocks[e, r], (i, j))


@pytest.mark.parametrize('shape', [(7, 8, 32, 32, 16), (5, 1, 64, 64, 16), (4, 4, 32, 48, 8), (2, 3, 128, 128, 32),
                                   (3, 2, 40, 40, 12), (9, 3, 24, 24, 4)])
@pytest.mark.parametrize('goal,minorder', [(True, 1), (True, 0), (True, 2), (False, 1)])
def test_fused_scoring_equals_separate_kernels(B, shape, goal, minorder):
  """srl_score_f32 (one launch) == srl_maxplus_f32 + srl_goal_overlap_f32 +
  srl_select_f32, and both equal the oracle's Baseline.call per view."""
  from stackrl_b200 import capi
  E, R, H, W, h = shape
  walls, rocks, _ = synth.placement_batch(51, E, R, H, W, h)
  goals = synth.goals(52, E, H, W)
  goals *= np.linspace(0.6, 1.4, E, dtype='float32')[:, None, None]   # per-env goal level
  dev = torch.device('cuda')
  wd, gd, rd = (torch.from_numpy(x).to(dev) for x in (walls, goals, rocks))
  scorer = B.PlacementScorer(goal=goal, minorder=minorder)
  sep = scorer(wd, gd, rd, fused=False)
  try:
    values, actions, best = capi.score_f32(
      wd, gd if goal else None, rd, None if goal else gd.amax(dim=(1, 2)),
      level_mode=2 if goal else 1, minorder=minorder)
  except capi.SrlError as err:
    # big maps are outside the single-launch kernel; the scorer falls back
    assert err.code == capi.SRL_E_UNSUPPORTED and H >= 128
    default = scorer(wd, gd, rd, fused='full')
    assert torch.equal(default['actions'], sep['actions'])
    assert torch.equal(default['best'], sep['best'])
    return
  assert torch.equal(values, sep['values'])
  assert torch.equal(actions, sep['actions'])
  assert torch.equal(best, sep['best'])
  full = scorer(wd, gd, rd, fused='full')
  assert torch.equal(full['actions'], actions) and torch.equal(full['best'], best)
  default = scorer(wd, gd, rd, want_shown=True)   # max-plus + mask_select (2 launches)
  assert torch.equal(default['actions'], actions) and torch.equal(default['best'], best)
  sep_shown = scorer(wd, gd, rd, want_shown=True, fused=False)['shown']
  assert torch.equal(default['shown'], sep_shown)
  for e in (0, E - 1):
    wg = np.stack([walls[e], goals[e]], -1)
    for r in range(R):
      a, _ = S.baseline_call((wg, rocks[e, r][..., None]), goal=goal, minorder=minorder)
      assert int(actions[e, r]) == a
  # score maps are optional
  none_values, actions2, _ = capi.score_f32(wd, gd if goal else None, rd,
                                            None if goal else gd.amax(dim=(1, 2)),
                                            level_mode=2 if goal else 1, minorder=minorder,
                                            want_values=False)
  assert none_values is None and torch.equal(actions2, actions)


def test_fused_scoring_full_config2():
  """BASELINE config 2 at full size: fused == separate on every environment."""
  from stackrl_b200 import baselines, capi
  E, R, H, W, h = 4096, 8, 32, 32, 16
  walls, rocks, _ = synth.placement_batch(0, E, R, H, W, h)
  goals = synth.goals(7, E, H, W)
  dev = torch.device('cuda')
  wd, gd, rd = (torch.from_numpy(x).to(dev) for x in (walls, goals, rocks))
  sep = baselines.PlacementScorer()(wd, gd, rd, fused=False)
  values, actions, best = capi.score_f32(wd, gd, rd)
  assert torch.equal(values, sep['values'])
  assert torch.equal(actions, sep['actions']) and torch.equal(best, sep['best'])


def test_fused_scoring_rejects_unsupported_shapes():
  from stackrl_b200 import capi
  dev = torch.device('cuda')
  with pytest.raises(capi.SrlError) as err:                 # 17-wide rows: no 16-B rows
    capi.score_f32(torch.zeros((1, 17, 17), device=dev), torch.zeros((1, 17, 17), device=dev),
                   torch.zeros((1, 1, 5, 5), device=dev))
  assert err.value.code == capi.SRL_E_UNSUPPORTED


@pytest.mark.parametrize('dtype', ['float32', 'uint8'])
def test_placement_scorer_difference_batched(B, dtype):
  """PlacementScorer('difference') on a device batch == looping the oracle's
  Baseline(method='difference') over environments and views (float64 maps)."""
  E, R, H, W, h = 3, 4, 32, 32, 8
  dev = torch.device('cuda')
  obs = [synth.batched_observation(60 + e, H, W, h, R, dtype=dtype) for e in range(E)]
  walls = np.stack([o[0][0, ..., 0] for o in obs])
  goals = np.stack([o[0][0, ..., 1] for o in obs])
  rocks = np.stack([o[1][..., 0] for o in obs])
  out = B.PlacementScorer('difference')(torch.from_numpy(walls).to(dev),
                                        torch.from_numpy(goals).to(dev),
                                        torch.from_numpy(rocks).to(dev), want_shown=True)
  assert out['values'].dtype == torch.float64
  for e in range(E):
    (k, idx), v = S.greedy(
      obs[e], lambda o: S.baseline_call(o, method='difference'), value=True, batched=True,
      batchwise=True)
    assert tuple(out['best'][e].cpu().numpy()) == (k, idx)
    assert np.array_equal(out['shown'][e].cpu().numpy().reshape(R, -1), v)


@pytest.mark.parametrize('dtype', ['float32', 'uint8'])
@pytest.mark.parametrize('envs,chunks', [(10, 4), (3, 4), (1, 1)])
def test_host_pipeline_equals_oracle(B, dtype, envs, chunks):
  """numpy observations in, numpy actions out (the call bench.py times as `e2e`):
  chunked copies overlapping the kernels give the oracle's picks, for float32 and
  for the uint8 observations of the registered environments; ragged chunking
  (fewer environments than chunks) included."""
  E, R, H, W, h = envs, 8, 32, 32, 16
  walls, rocks, _ = synth.placement_batch(41, E, R, H, W, h)
  goals = synth.goals(42, E, H, W)
  walls, goals, rocks = (synth.to_dtype(x, dtype) for x in (walls, goals, rocks))
  pipe = B.HostPipeline(B.PlacementScorer(), E, R, H, W, h, chunks=chunks,
                        dtype=getattr(torch, dtype))
  assert pipe.h2d_bytes == (walls.nbytes + goals.nbytes + rocks.nbytes)
  for _ in range(2):                       # staging buffers are reused
    actions, best = pipe(walls, goals, rocks)
    for e in range(E):
      wg = np.stack([walls[e], goals[e]], -1)
      obs = (np.stack([wg] * R), rocks[e][..., None])
      k, idx = S.greedy(obs, lambda o: S.baseline_call(o, method='height'),
                        batched=True, batchwise=True)
      assert (best[e, 0], best[e, 1]) == (k, idx)
      assert actions[e, k] == idx
  with pytest.raises(TypeError):
    B.HostPipeline(B.PlacementScorer(), E, R, H, W, h, dtype=torch.float64)


@pytest.mark.parametrize('dtype', ['float32', 'uint8'])
def test_host_pipeline_cuda_graph_replay(B, dtype):
  """The captured step (one graph launch) gives the eager step's actions, also
  after the staged observations change (the graph reads the same pinned buffers)."""
  E, R, H, W, h = 12, 8, 32, 32, 16
  pipe = B.HostPipeline(B.PlacementScorer(), E, R, H, W, h, chunks=3,
                        dtype=getattr(torch, dtype))
  batches = []
  for seed in (51, 52, 53):
    walls, rocks, _ = synth.placement_batch(seed, E, R, H, W, h)
    goals = synth.goals(seed + 100, E, H, W)
    batches.append([synth.to_dtype(x, dtype) for x in (walls, goals, rocks)])
  eager = [tuple(a.copy() for a in pipe(*b)) for b in batches]
  pipe.capture()
  for b, (actions, best) in zip(batches, eager):
    got_actions, got_best = pipe(*b)
    assert np.array_equal(got_actions, actions) and np.array_equal(got_best, best)
