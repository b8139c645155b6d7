"""oracle/observe_np.py (depth->elevation, packing, IoU/OR) and the oracle's
pose restatement against episodes of the UNMODIFIED reference env recorded on
the fake pybullet backend (tests/golden/observe.npz).  CPU only."""
import numpy as np
import pytest

from oracle import observe_np as O
from oracle import scoring_np as S
from tests.conftest import GEOM, split_depths

EPISODES = [('stack_f32', 'float32', 1), ('stack_u8', 'uint8', 1), ('test_f32_rot8', 'float32', 8)]


@pytest.mark.parametrize('name,dtype,views', EPISODES)
def test_conversion_and_packing(observe_golden, name, dtype, views):
  g = observe_golden
  goal = g[name + '/goal']
  n_steps = int(g[name + '/n_recorded'])
  rock = None
  for k in range(n_steps + 1):
    key = '{}/s{}'.format(name, k)
    depths = split_depths(g, key)
    wall = O.wall_elevation(depths[0], GEOM['max_z'])
    assert np.array_equal(wall, g[key + '/overhead_map'])
    if len(depths) > 1:
      rock = [O.rock_elevation(d, GEOM['object_z']) for d in depths[1:]]
      assert len(rock) == views
    if k == n_steps:
      break
    want_rock = g[key + '/object_map']
    if views == 1:
      assert np.array_equal(rock[0], want_rock)
      obs = O.pack_obs(wall, goal, rock[0], dtype, GEOM['max_z'], GEOM['object_max_dimension'])
    else:
      assert np.array_equal(np.array(rock), want_rock)
      obs = O.pack_obs_batched(wall, goal, rock, dtype, GEOM['max_z'],
                               GEOM['object_max_dimension'])
    assert obs[0].dtype == g[key + '/obs0'].dtype
    assert np.array_equal(obs[0], g[key + '/obs0'])
    assert np.array_equal(obs[1], g[key + '/obs1'])


@pytest.mark.parametrize('name,dtype,views', EPISODES)
def test_pose_and_actions(observe_golden, name, dtype, views):
  g = observe_golden
  px = GEOM['pixel']
  size = (GEOM['h'] * px, GEOM['h'] * px, GEOM['object_z'])
  Pw = GEOM['W'] - GEOM['h'] + 1
  for k in range(int(g[name + '/n_recorded'])):
    key = '{}/s{}'.format(name, k)
    obs = (g[key + '/obs0'], g[key + '/obs1'])
    if views == 1:
      a, _ = S.baseline_call(obs, method='height')
      assert a == int(g[key + '/action'])
      view, flat = None, a
      rock = g[key + '/object_map']
    else:
      (view, flat), _ = S.greedy(obs, lambda o: S.baseline_call(o, method='height'),
                                 value=True, batched=True, batchwise=True)
      assert (view, flat) == tuple(g[key + '/action'])
      rock = g[key + '/object_map'][view]
    pos = S.pose(g[key + '/overhead_map'], rock, (flat // Pw, flat % Pw), (px, px), size)
    assert np.array_equal(np.array(pos, dtype='float64'), g[key + '/pose_position'])


@pytest.mark.parametrize('name', ['stack_f32', 'test_f32_rot8'])
def test_iou_and_or_rewards(observe_golden, name):
  g = observe_golden
  goal = g[name + '/goal']
  memory = {'iou': 0., 'or': 0.}
  for k in range(int(g[name + '/n_recorded'])):
    wall = g['{}/s{}/overhead_map'.format(name, k + 1)]
    want = g['{}/s{}/rewards'.format(name, k)]
    for col, metric in ((0, 'iou'), (1, 'or')):
      r, memory[metric] = O.reward(wall, goal, GEOM['goal_z'], metric, memory[metric])
      assert r == want[col], (k, metric)


def test_c1_stack_v0_episode_actions_and_final_packing(observe_golden):
  """BASELINE config 1: the 30-rock Stack-v0 episode (uint8 observations) of the unmodified
  reference.  The oracle's height policy picks the recorded action from every recorded
  observation, the final wall packs to the recorded observation, and the 30 step rewards
  telescope to the final wall's IoU / occupation ratio."""
  g = observe_golden
  name = 'c1_stack_v0_30'
  assert int(g[name + '/n_recorded']) == 30
  for k in range(30):
    key = '{}/s{}'.format(name, k)
    a, _ = S.baseline_call((g[key + '/obs0'], g[key + '/obs1']), method='height')
    assert a == int(g[key + '/action']), k
  wall = g[name + '/s30/overhead_map']
  obs = O.pack_obs(wall, g[name + '/goal'], np.zeros((32, 32), 'float32'), 'uint8',
                   GEOM['max_z'], GEOM['object_max_dimension'])
  assert np.array_equal(obs[0], g[name + '/s30/obs0'])
  tot_iou = sum(float(g['{}/s{}/rewards'.format(name, k)][0]) for k in range(30))
  tot_or = sum(float(g['{}/s{}/rewards'.format(name, k)][1]) for k in range(30))
  _, iou = O.reward(wall, g[name + '/goal'], GEOM['goal_z'], 'iou', 0.)
  _, occ = O.reward(wall, g[name + '/goal'], GEOM['goal_z'], 'or', 0.)
  assert abs(tot_iou - iou) < 1e-12 and abs(tot_or - occ) < 1e-12


def test_contact_precheck_known_cases():
  """The heightmap contact pre-check on configurations with a known answer."""
  h = 8
  rock = np.full((h, h), 0.05, 'float32')                  # flat underside
  wall = np.zeros((20, 20), 'float32')
  count, mask, ok = O.contact_precheck(wall, rock, (3, 4))
  assert (count, mask, ok) == (64, 0xff, True)             # rests on all of it
  wall[3:7, :] = np.float32(0.02)                          # a ledge under the first rows only
  count, mask, ok = O.contact_precheck(wall, rock, (3, 4))
  assert count == 32 and not ok and mask == 0b00111100     # one half-plane only: tips over
  rock2 = rock.copy()
  rock2[0, 0] = np.float32(0.08)                           # a single spike
  wall[:] = 0
  assert O.contact_precheck(wall, rock2, (0, 0)) == (1, 1 << 4, False)
  # background cells (rock <= 1e-4) never touch
  rock3 = np.zeros((h, h), 'float32')
  rock3[2:6, 2:6] = np.float32(0.03)
  assert O.contact_precheck(wall, rock3, (5, 5))[0] == 16


@pytest.mark.parametrize('n,seed', [(2, 0), (9, 1)])
def test_oracle_zbuffer_renders_a_tessellated_plane_analytically(n, seed):
  """The oracle z-buffer against an analytic known answer (the same scene as the GPU test
  test_tessellated_plane_analytic): every covered pixel holds the plane's height at the
  pixel centre to within the elevation formula's float32 rounding (1.5 x 2^-14 m)."""
  from oracle import raster_np as R
  from stackrl_b200 import camera
  geo = camera.ObserverGeometry(128, 32, 0.125 / 32, 0.375)
  rng = np.random.default_rng(seed)
  a, b, c = 0.3, -0.17, 0.16
  x0, x1, y0, y1 = 0.0813, 0.4191, 0.0779, 0.4233
  X, Y = np.meshgrid(np.linspace(x0, x1, n), np.linspace(y0, y1, n), indexing='ij')
  verts = np.stack([X, Y, c + a * X + b * Y], -1).reshape(-1, 3).astype('float32')
  tris = []
  for i in range(n - 1):
    for j in range(n - 1):
      p00, p01, p10, p11 = i * n + j, i * n + j + 1, (i + 1) * n + j, (i + 1) * n + j + 1
      if rng.random() < 0.5:
        tris += [(p00, p10, p11), (p00, p11, p01)]
      else:
        tris += [(p00, p10, p01), (p10, p11, p01)]
  depth = R.render_depth(geo.overhead_view, geo.overhead_projection, 128, 128,
                         [(verts, np.asarray(tris, dtype='int32'), np.identity(3), np.zeros(3))])
  m = O.wall_elevation(depth, 0.375)
  px = 0.125 / 32
  ctr = (np.arange(128) + 0.5) * px
  inx, iny = (ctr > x0 + px) & (ctr < x1 - px), (ctr > y0 + px) & (ctr < y1 - px)
  want = c + a * ctr[:, None] + b * ctr[None, :]
  assert np.abs(m - want)[np.ix_(inx, iny)].max() <= 1.5 * 2.0 ** -14 + 2e-6
  assert np.all(m[ctr < x0 - px, :] == 0) and np.all(m[:, ctr > y1 + px] == 0)


def test_oracle_rock_view_plane_analytic():
  """Rock view (the camera under the spawned rock, observer.py:262-293, with the column
  mirror of :277): a tilted plane through the spawn point maps to object_z / 2 - (a x + b
  y) at the pixel centres x = (i + 0.5 - 16) px, y = (j + 0.5 - 16) px, to within the
  elevation formula's float32 rounding -- pins the axes and the mirror analytically."""
  from oracle import raster_np as R
  from stackrl_b200 import camera
  geo = camera.ObserverGeometry(128, 32, 0.125 / 32, 0.375)
  spawn = ((0., 0., 0.5), (0., 0., 0., 1.))
  a, b, n = 0.21, -0.13, 9
  X, Y = np.meshgrid(np.linspace(-0.0571, 0.0593, n), np.linspace(-0.0589, 0.0577, n),
                     indexing='ij')
  verts = np.stack([X, Y, a * X + b * Y], -1).reshape(-1, 3).astype('float32')
  tris = []
  for i in range(n - 1):
    for j in range(n - 1):
      p00, p01, p10, p11 = i * n + j, i * n + j + 1, (i + 1) * n + j, (i + 1) * n + j + 1
      tris += [(p00, p10, p11), (p00, p11, p01)] if (i + j) % 2 else \
        [(p00, p10, p01), (p10, p11, p01)]
  bodies = [(verts, np.asarray(tris, dtype='int32'), np.identity(3), np.array(spawn[0]))]
  m = O.rock_elevation(R.render_depth(geo.object_view(spawn, 0), geo.object_projection, 32, 32,
                                       bodies), geo.object_z)
  ctr = (np.arange(32) + 0.5 - 16) * (0.125 / 32)
  inside = np.abs(ctr) < 0.05
  want = geo.object_z / 2 - (a * ctr[:, None] + b * ctr[None, :])
  assert np.abs(m - want)[np.ix_(inside, inside)].max() <= 1.5 * 2.0 ** -14 + 2e-6
