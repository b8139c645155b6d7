"""Parity of the CUDA max-plus path (through the C ABI) with the reference:
golden vectors produced by the reference's own ``height`` and the numpy oracle
on seeded inputs.  Bit-exact (integer/float32 max-plus is order independent and
every add is a single IEEE add)."""
import numpy as np
import pytest
import torch

from oracle import scoring_np as S
from stackrl_b200 import synth

pytestmark = pytest.mark.gpu

F32_CASES = ['c2like_f32', 'nonsquare_wall_f32', 'dense_rock_f32', 'empty_rock_f32',
             'ties_f32', 'flat_wall_f32', 'c4like_f32', 'c5like_f32',
             'full_rock_window_f32']


@pytest.fixture(scope='module')
def capi():
  from stackrl_b200 import capi
  return capi


def _oracle_maps(walls, rocks, level):
  """Loop the oracle's ``height`` over a planar batch -> [E,R,Ph,Pw] float32."""
  E, R = rocks.shape[:2]
  out = []
  for e in range(E):
    goal = np.full(walls.shape[1:], level[e], dtype='float32')
    wg = np.stack([walls[e], goal], axis=-1)
    out.append([S.height((wg, rocks[e, r][..., None])) for r in range(R)])
  return np.asarray(out).astype('float32')


@pytest.mark.parametrize('case', F32_CASES)
def test_height_matches_reference_golden(scoring_golden, case):
  from stackrl_b200 import baselines
  got = baselines.height(scoring_golden.obs(case))
  want = scoring_golden[case + '/height']
  assert got.dtype == want.dtype and got.shape == want.shape
  assert np.array_equal(got, want)


@pytest.mark.parametrize('variant', ['0', '1'])
@pytest.mark.parametrize('shape', [
  # E, R, H, W, h
  (5, 8, 32, 32, 16),      # config 2 geometry, ragged last CTA group
  (3, 3, 64, 64, 16),      # config 4 geometry
  (2, 2, 128, 128, 32),    # config 1/5 geometry
  (4, 1, 20, 28, 6),       # rock side not a multiple of 4 (no TMA rows)
  (3, 2, 17, 19, 5),       # wall width not a multiple of 4 (no TMA)
  (2, 5, 40, 24, 12),
  (1, 1, 9, 9, 9),         # single candidate position
  (7, 36, 48, 48, 16),     # many rotations (rotation chunking)
])
def test_batched_matches_oracle(capi, monkeypatch, shape, variant):
  monkeypatch.setenv('SRL_MAXPLUS_VARIANT', variant)
  E, R, H, W, h = shape
  walls, rocks, level = synth.placement_batch(11, E, R, H, W, h)
  level = (level * np.linspace(0.5, 1.5, E)).astype('float32')
  dev = torch.device('cuda')
  got = capi.maxplus_f32(torch.from_numpy(walls).to(dev),
                         torch.from_numpy(rocks).to(dev),
                         torch.from_numpy(level).to(dev)).cpu().numpy()
  want = _oracle_maps(walls, rocks, level)
  assert np.array_equal(got, want)


@pytest.mark.timeout(120)
@pytest.mark.parametrize('shape', [
  # Few items per environment: a 32-item unit of the stream kernel spans more
  # environments than the ring has slots unless the dispatcher refuses (a consumer
  # would wait for a slot only its own completion can free: a hang, not an error).
  (64, 1, 16, 16, 16),      # 1 item per environment, unit spans 32 environments
  (8, 1, 64, 64, 60),       # 5 items per environment, 3 ring slots
  (100, 2, 20, 16, 16),     # 10 items per environment
  (300, 1, 32, 32, 32),     # one candidate position per map
  (40, 1, 48, 48, 44),      # big rock, few positions
])
def test_tiny_maps_do_not_wrap_the_stream_ring(capi, shape):
  E, R, H, W, h = shape
  walls, rocks, level = synth.placement_batch(21, E, R, H, W, h)
  dev = torch.device('cuda')
  got = capi.maxplus_f32(torch.from_numpy(walls).to(dev), torch.from_numpy(rocks).to(dev),
                         torch.from_numpy(level).to(dev))
  torch.cuda.synchronize()
  assert np.array_equal(got.cpu().numpy(), _oracle_maps(walls, rocks, level))


def test_caller_buffers_are_validated(capi):
  """An undersized / misplaced `out` is refused before any kernel writes into it."""
  dev = torch.device('cuda')
  walls, rocks, level = synth.placement_batch(1, 4, 2, 32, 32, 16)
  w, r, l = (torch.from_numpy(x).to(dev) for x in (walls, rocks, level))
  with pytest.raises(ValueError):
    capi.maxplus_f32(w, r, l, out=torch.empty((4, 2, 17, 16), device=dev))
  with pytest.raises(ValueError):
    capi.maxplus_f32(w, r, l, out=torch.empty((4, 2, 17, 17)).pin_memory())
  with pytest.raises(TypeError):
    capi.maxplus_f32(w, r, l, out=torch.empty((4, 2, 17, 17), device=dev, dtype=torch.float64))
  w8 = (w * 600).to(torch.uint8)
  r8 = (r * 600).to(torch.uint8)
  l8 = torch.full((4,), 170, dtype=torch.uint8, device=dev)
  with pytest.raises(ValueError):
    capi.maxplus_u8(w8, r8, l8, out=torch.empty((4, 2, 17), dtype=torch.float64, device=dev))


def test_goal_level_kernel(capi):
  """get_inputs' goal.max() (baselines.py:23) without a torch reduction."""
  dev = torch.device('cuda')
  goals = synth.goals(3, 37, 40, 24)
  goals[5] = 0.
  got = capi.goal_level(torch.from_numpy(goals).to(dev)).cpu().numpy()
  assert np.array_equal(got, goals.reshape(37, -1).max(axis=1))
  g8 = synth.to_dtype(goals, 'uint8')
  got8 = capi.goal_level(torch.from_numpy(g8).to(dev)).cpu().numpy()
  assert np.array_equal(got8, g8.reshape(37, -1).max(axis=1))


def test_no_level_and_pose_threshold(capi):
  """level=None, threshold=1e-4: the Observer.pose mask (observer.py:405-409)."""
  E, R, H, W, h = 3, 2, 24, 24, 8
  walls, rocks, _ = synth.placement_batch(5, E, R, H, W, h)
  rocks[rocks > 0] -= np.float32(0.0624)    # push some cells under 1e-4
  rocks = np.maximum(rocks, 0).astype('float32')
  dev = torch.device('cuda')
  got = capi.maxplus_f32(torch.from_numpy(walls).to(dev),
                         torch.from_numpy(rocks).to(dev), None,
                         threshold=1e-4).cpu().numpy()
  for e in range(E):
    for r in range(R):
      live = rocks[e, r] > np.float32(1e-4)
      for i in (0, 7, H - h):
        for j in (0, 3, W - h):
          lifted = walls[e, i:i + h, j:j + h] + rocks[e, r]
          want = np.where(live, lifted, 0).max()
          assert got[e, r, i, j] == want


def test_full_size_config2_properties(capi):
  """BASELINE config 2 at full size (4096 envs x 8 rotations, 32x32 / 16x16):
  size-independent properties + an oracle check on a seeded subsample."""
  E, R, H, W, h = 4096, 8, 32, 32, 16
  walls, rocks, level = synth.placement_batch(0, E, R, H, W, h)
  dev = torch.device('cuda')
  w_d, r_d, l_d = (torch.from_numpy(x).to(dev) for x in (walls, rocks, level))
  out = capi.maxplus_f32(w_d, r_d, l_d)
  assert out.shape == (E, R, 17, 17)
  # (1) translation equivariance: shifting the wall by one row shifts the map.
  shifted = torch.roll(w_d, shifts=1, dims=1)
  out_s = capi.maxplus_f32(shifted, r_d, l_d)
  assert torch.equal(out_s[:, :, 1:, :], out[:, :, :-1, :])
  # (2) monotone in the wall: raising the wall never lowers a drop height.
  out_up = capi.maxplus_f32(w_d + 0.25, r_d, l_d)
  assert bool((out_up >= out).all())
  # (3) lower bound: at least the rock's own maximum over the level.
  rock_top = (r_d / l_d[:, None, None, None]).amax(dim=(2, 3))
  assert bool((out >= rock_top[:, :, None, None]).all())
  # (4) idempotent launch (no state between calls).
  assert torch.equal(out, capi.maxplus_f32(w_d, r_d, l_d))
  # (5) oracle on a subsample of environments.
  pick = np.random.default_rng(0).choice(E, 24, replace=False)
  want = _oracle_maps(walls[pick], rocks[pick], level[pick])
  assert np.array_equal(out[torch.from_numpy(pick).to(dev)].cpu().numpy(), want)


def test_empty_batch_and_bad_arguments(capi):
  dev = torch.device('cuda')
  out = capi.maxplus_f32(torch.empty((0, 8, 8), device=dev),
                         torch.empty((0, 1, 4, 4), device=dev))
  assert out.shape == (0, 1, 5, 5)
  with pytest.raises(TypeError):
    capi.maxplus_f32(torch.zeros((1, 8, 8)), torch.zeros((1, 1, 4, 4)))
  with pytest.raises(capi.SrlError):
    capi.maxplus_f32(torch.zeros((1, 4, 4), device=dev),
                     torch.zeros((1, 1, 8, 8), device=dev),
                     out=torch.zeros((1, 1, 1, 1), device=dev))


def _numpy_maps(walls, rocks, level, threshold=0.):
  """Vectorised restatement of baselines.py:21-43 for planar batches (float32
  division, float32 add, masked cells contribute 0)."""
  E, R, h = rocks.shape[0], rocks.shape[1], rocks.shape[2]
  out = []
  for e in range(E):
    o = walls[e] / level[e] if level is not None else walls[e]
    win = np.lib.stride_tricks.sliding_window_view(o, (h, h))
    maps = []
    for r in range(R):
      n = rocks[e, r] / level[e] if level is not None else rocks[e, r]
      maps.append(np.where(n > np.float32(threshold), win + n, np.float32(0)).max(axis=(2, 3)))
    out.append(maps)
  return np.asarray(out, dtype='float32')


@pytest.mark.parametrize('shape', [
  (301, 1, 32, 32, 16),    # 17 items per environment: a warp pass spans 2-3 environments
  (150, 3, 20, 20, 8),     # ring slots reused many times per CTA
  (1200, 8, 32, 32, 16),   # CTA ranges that start and end inside an environment
  (400, 2, 64, 64, 16),    # two strips per output row
  (40, 8, 48, 48, 16),
])
def test_stream_kernel_item_stream(capi, monkeypatch, shape):
  """maxplus_stream_kernel: the item stream is cut in 32-item units that ignore
  environment boundaries; every cut must give the oracle's maps, and the same
  bits as the barrier-synchronised staged kernel."""
  E, R, H, W, h = shape
  walls, rocks, level = synth.placement_batch(21, E, R, H, W, h)
  level = (level * np.linspace(0.6, 1.4, E)).astype('float32')
  dev = torch.device('cuda')
  args = [torch.from_numpy(x).to(dev) for x in (walls, rocks, level)]
  got = capi.maxplus_f32(*args)
  pick = np.random.default_rng(1).choice(E, min(E, 48), replace=False)
  want = _numpy_maps(walls[pick], rocks[pick], level[pick])
  assert np.array_equal(got[torch.from_numpy(pick).to(dev)].cpu().numpy(), want)
  monkeypatch.setenv('SRL_MP_MODE', '1')
  assert torch.equal(got, capi.maxplus_f32(*args))


def test_stream_kernel_negative_values(capi):
  """Tiles with negative heights take the FMNMX3 sweep (the integer-max sweep
  is only exact for non-negative operands); a batch mixes both kinds."""
  E, R, H, W, h = 64, 8, 32, 32, 16
  walls, rocks, level = synth.placement_batch(3, E, R, H, W, h)
  rng = np.random.default_rng(9)
  walls[::3] -= np.float32(0.2)                      # negative walls in every third env
  rocks[1::4] -= rng.uniform(0, 0.05, rocks[1::4].shape).astype('float32')
  dev = torch.device('cuda')
  for thr in (0., -0.02):
    got = capi.maxplus_f32(torch.from_numpy(walls).to(dev), torch.from_numpy(rocks).to(dev),
                           torch.from_numpy(level).to(dev), threshold=thr).cpu().numpy()
    assert np.array_equal(got, _numpy_maps(walls, rocks, level, thr))


@pytest.mark.parametrize('tile', ['5', '9', '13', '17', '21', '25'])
@pytest.mark.parametrize('mode', ['2', '1', '0'])
def test_every_tile_width_and_kernel(capi, monkeypatch, tile, mode):
  """All per-thread tile widths (T outputs per thread) of the three kernels --
  stream (2), staged (1), direct (0) -- on shapes with one and several strips per
  output row, including widths that do not divide the row."""
  monkeypatch.setenv('SRL_MP_T', tile)
  monkeypatch.setenv('SRL_MP_MODE', mode)
  dev = torch.device('cuda')
  for E, R, H, W, h in ((9, 3, 32, 32, 16), (5, 2, 48, 64, 16), (3, 1, 64, 40, 8)):
    walls, rocks, level = synth.placement_batch(77, E, R, H, W, h)
    got = capi.maxplus_f32(torch.from_numpy(walls).to(dev), torch.from_numpy(rocks).to(dev),
                           torch.from_numpy(level).to(dev)).cpu().numpy()
    assert np.array_equal(got, _numpy_maps(walls, rocks, level))


def _quantised_batch(seed, E, R, H, W, h, qlog2=-14, wall_top=0.3):
  """Heightmaps as the rasteriser leaves them: non-negative multiples of 2^qlog2
  (observer.py:259-260 evaluates the elevation at magnitude 1000 in float32)."""
  walls, rocks, level = synth.placement_batch(seed, E, R, H, W, h)
  q = np.float32(2.0 ** qlog2)
  walls = (np.round(walls * (wall_top / 0.3) / q) * q).astype('float32')
  rocks = (np.round(rocks / q) * q).astype('float32')
  return walls, rocks, level


@pytest.mark.parametrize('shape', [(64, 8, 32, 32, 16), (700, 8, 32, 32, 16), (150, 2, 64, 64, 16),
                                   (33, 3, 48, 40, 8),
                                   (450, 1, 64, 64, 16)])   # config-4 geometry, wide-tile batch
def test_fixed_point_sweep_is_bit_exact(capi, shape):
  """srl_maxplus_f32_q: quantised heightmaps with a power-of-two level are swept
  in 16-bit fixed point (VIADDMNMX.S16x2) -- same bits as numpy's float32."""
  E, R, H, W, h = shape
  walls, rocks, level = _quantised_batch(5, E, R, H, W, h)
  dev = torch.device('cuda')
  args = [torch.from_numpy(x).to(dev) for x in (walls, rocks, level)]
  got = capi.maxplus_f32(*args, quantum_log2=-14)
  assert torch.equal(got, capi.maxplus_f32(*args))           # float sweep, same bits
  pick = np.random.default_rng(2).choice(E, min(E, 40), replace=False)
  want = _numpy_maps(walls[pick], rocks[pick], level[pick])
  assert np.array_equal(got[torch.from_numpy(pick).to(dev)].cpu().numpy(), want)


def test_wide_tile_of_the_stream_kernel_equals_the_default_tile(capi, monkeypatch):
  """49 output columns (config 4): a batch big enough for the stream kernel is swept with two
  strips of 25 columns instead of three of 17 -- same bits as the forced 17-column tile, float
  and fixed-point sweeps."""
  dev = torch.device('cuda')
  walls, rocks, level = _quantised_batch(9, 600, 1, 64, 64, 16)
  fw, fr, _ = synth.placement_batch(10, 600, 1, 64, 64, 16)
  walls[::3], rocks[::3] = fw[::3], fr[::3]              # a third of the batch is not quantised
  args = [torch.from_numpy(x).to(dev) for x in (walls, rocks, level)]
  got_q, got_f = capi.maxplus_f32(*args, quantum_log2=-14), capi.maxplus_f32(*args)
  monkeypatch.setenv('SRL_MP_T', '17')
  assert torch.equal(got_q, capi.maxplus_f32(*args, quantum_log2=-14))
  assert torch.equal(got_f, capi.maxplus_f32(*args))
  assert torch.equal(got_q, got_f)
  pick = np.arange(0, 600, 25)
  assert np.array_equal(got_q[torch.from_numpy(pick).to(dev)].cpu().numpy(),
                        _numpy_maps(walls[pick], rocks[pick], level[pick]))


def test_fixed_point_sweep_falls_back_per_environment(capi):
  """The hint is only a hint: environments that are not quantised, too tall for
  14 bits, negative, or normalised by a level that is not a power of two take the
  float sweep; a batch can mix all of them."""
  E, R, H, W, h = 96, 8, 32, 32, 16
  walls, rocks, level = _quantised_batch(6, E, R, H, W, h)
  fw, fr, _ = synth.placement_batch(7, E, R, H, W, h)
  walls[1::6] = fw[1::6]                       # arbitrary float32 walls
  rocks[2::6] = fr[2::6]                       # arbitrary float32 rocks
  walls[3::6] += np.float32(1.5)               # counts above 2^14
  walls[4::6, 0, 0] = np.float32(-2.0 ** -14)  # one negative cell
  level = level.copy()
  level[5::6] = np.float32(0.3)                # not a power of two
  dev = torch.device('cuda')
  args = [torch.from_numpy(x).to(dev) for x in (walls, rocks, level)]
  got = capi.maxplus_f32(*args, quantum_log2=-14).cpu().numpy()
  assert np.array_equal(got, _numpy_maps(walls, rocks, level))
  # no level at all, pose threshold
  got = capi.maxplus_f32(args[0], args[1], None, threshold=1e-4, quantum_log2=-14)
  assert torch.equal(got, capi.maxplus_f32(args[0], args[1], None, threshold=1e-4))
