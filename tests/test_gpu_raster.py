"""Rasterisation + observation path on the GPU.

Parity target: the oracle's software z-buffer (oracle/csrc/oracle.c through
oracle.raster_np) -- NOT pybullet's TinyRenderer, which the reference uses but
does not contain (PARITY UNPINNED for the depth image itself; the arithmetic the
reference applies to it IS pinned by tests/golden/observe.npz).  CUDA and oracle
run the same IEEE op sequence, so depth images and elevations are compared
bit-for-bit, which is stricter than the 1e-5 relative bound of the task."""
import numpy as np
import pytest
import torch

from oracle import observe_np as O
from oracle import raster_np as R
from stackrl_b200 import synth
from tests.conftest import GEOM, split_depths

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
  from stackrl_b200 import camera, capi, envs, meshes, observer
  return dict(camera=camera, capi=capi, envs=envs, meshes=meshes, observer=observer)


def _random_scene(meshes, seed, n_rocks, subdivisions=2):
  rng = np.random.default_rng(seed)
  verts, tris = meshes.synthetic_rocks(seed, n_rocks, subdivisions, max_dimension=0.12)
  bodies = []
  for k in range(n_rocks):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    pos = np.array([rng.uniform(0.08, 0.42), rng.uniform(0.08, 0.42), rng.uniform(0.03, 0.3)])
    bodies.append((verts[k], tris, R.quat_matrix(q), pos))
  return bodies


def _gpu_render(mods, bodies, view, proj, rows, cols, mode, zrange):
  obs, capi = mods['observer'], mods['capi']
  verts, tris, inst = obs._instances(bodies)
  job = obs._job(view, proj, 0, len(inst), zrange)
  dev = torch.device('cuda')
  return capi.raster(torch.from_numpy(verts).to(dev), torch.from_numpy(tris).to(dev),
                     inst, job, rows, cols, mode)[0].cpu().numpy()


@pytest.mark.parametrize('seed,n_rocks,sub', [(0, 1, 2), (1, 6, 2), (2, 12, 1), (3, 3, 3)])
def test_wall_image_matches_oracle_bitwise(mods, seed, n_rocks, sub):
  geo = mods['camera'].ObserverGeometry(128, 32, 0.125 / 32, 0.375)
  bodies = _random_scene(mods['meshes'], seed, n_rocks, sub)
  want_d = R.render_depth(geo.overhead_view, geo.overhead_projection, 128, 128, bodies)
  got_d = _gpu_render(mods, bodies, geo.overhead_view, geo.overhead_projection, 128, 128,
                      mods['capi'].RASTER_DEPTH, 0.375)
  assert np.array_equal(got_d, want_d)
  assert (want_d < 1).sum() > 50                      # something was drawn
  got = _gpu_render(mods, bodies, geo.overhead_view, geo.overhead_projection, 128, 128,
                    mods['capi'].RASTER_WALL, 0.375)
  want = O.wall_elevation(want_d, 0.375)
  assert np.array_equal(got, want)
  np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)   # the bound the task states


@pytest.mark.parametrize('seed', [0, 5])
def test_rock_images_all_orientations(mods, seed):
  geo = mods['camera'].ObserverGeometry(128, 32, 0.16 / 32, 0.375, orientation_freedom=3)
  verts, tris = mods['meshes'].synthetic_rocks(seed, 1, 3, max_dimension=0.16)
  spawn = ((0., 0., 0.375 + 0.16), (0., 0., 0., 1.))
  bodies = [(verts[0], tris, np.identity(3), np.array(spawn[0]))]
  for k in range(8):
    view = geo.object_view(spawn, k)
    want_d = R.render_depth(view, geo.object_projection, 32, 32, bodies)
    got = _gpu_render(mods, bodies, view, geo.object_projection, 32, 32,
                      mods['capi'].RASTER_ROCK, geo.object_z)
    want = O.rock_elevation(want_d, geo.object_z)
    assert np.array_equal(got, want)
    assert (got == 0).sum() > 20 and got.max() > geo.object_z / 2   # exact-zero background


def test_config3_rock_2000_triangles(mods):
  """BASELINE config 3's mesh size: a 2 000-triangle / 1 002-vertex synthetic rock on a
  32x32 image at 0.005 m/px (vertex cache and index staging both nearly full), plus a
  5 120-triangle one that exceeds the vertex cache (corners projected per triangle)."""
  geo = mods['camera'].ObserverGeometry(128, 32, 0.005, 0.375)
  spawn = ((0., 0., 0.375 + 0.16), (0., 0., 0., 1.))
  for kwargs in (dict(frequency=10), dict(subdivisions=4)):
    verts, tris = mods['meshes'].synthetic_rocks(4, 2, max_dimension=0.16, **kwargs)
    for k in range(2):
      bodies = [(verts[k], tris, np.identity(3), np.array(spawn[0]))]
      view = geo.object_view(spawn, 0)
      want_d = R.render_depth(view, geo.object_projection, 32, 32, bodies)
      got = _gpu_render(mods, bodies, view, geo.object_projection, 32, 32,
                        mods['capi'].RASTER_ROCK, geo.object_z)
      assert np.array_equal(got, O.rock_elevation(want_d, geo.object_z))
      assert (got > 0).sum() > 100


def test_box_known_answer(mods, observe_golden):
  """The reference's perfect box 0_0.obj (half extents 0.0536 x 0.0268 x
  0.0179 m) resting on the ground: flat top at 2*hz, footprint 2hx x 2hy."""
  v, t = observe_golden['mesh/0_0/verts'], observe_golden['mesh/0_0/tris']
  hx, hy, hz = np.abs(v).max(axis=0)
  geo = mods['camera'].ObserverGeometry(128, 32, 0.125 / 32, 0.375)
  bodies = [(v, t, np.identity(3), np.array([0.25, 0.25, hz]))]
  m = _gpu_render(mods, bodies, geo.overhead_view, geo.overhead_projection, 128, 128,
                  mods['capi'].RASTER_WALL, 0.375)
  top = m[m > 0]
  assert np.all(np.abs(top - 2 * hz) <= 2 ** -14 + 1e-7)     # float32 quantum at 1000 m
  px = 0.125 / 32
  assert abs(len(top) - (2 * hx / px) * (2 * hy / px)) <= 2 * (2 * hx + 2 * hy) / px
  rows, cols = np.nonzero(m)
  assert abs((rows.min() + rows.max() + 1) / 2 * px - 0.25) <= px    # rows run along x
  assert rows.max() - rows.min() > cols.max() - cols.min()          # long side along x
  assert np.all(m[m <= 0] == 0)                                     # ground exactly 0.0


def test_sphere_underside_analytic(mods):
  geo = mods['camera'].ObserverGeometry(128, 32, 0.125 / 32, 0.375)
  v, t = mods['meshes'].icosphere(4)
  radius = 0.05
  spawn = ((0., 0., 0.5), (0., 0., 0., 1.))
  bodies = [((v * radius).astype('float32'), t, np.identity(3), np.array(spawn[0]))]
  m = _gpu_render(mods, bodies, geo.object_view(spawn, 0), geo.object_projection, 32, 32,
                  mods['capi'].RASTER_ROCK, geo.object_z)
  c = (np.arange(32) + 0.5 - 16) * (0.125 / 32)
  rho2 = c[:, None] ** 2 + c[None, :] ** 2
  inside = rho2 < (radius - 0.125 / 32) ** 2
  want = geo.object_z / 2 + np.sqrt(np.clip(radius ** 2 - rho2, 0, None))
  assert np.abs(m - want)[inside].max() < 4e-4         # faceting + 2^-14 quantisation
  assert np.all(m[rho2 > (radius + 0.125 / 32) ** 2] == 0)


@pytest.mark.parametrize('n,seed', [(2, 0), (9, 1), (33, 2)])
def test_tessellated_plane_analytic(mods, n, seed):
  """Known answer that does not involve the oracle: a tilted plane z = c + a x + b y,
  tessellated into 2 (n-1)^2 triangles with random diagonals, seen by the overhead
  camera.  Every covered pixel (i, j) must hold the plane's height at its CENTRE
  x = (i + 0.5) px, y = (j + 0.5) px to within the float32 rounding of the reference's
  depth -> elevation formula (observer.py:259-260: three roundings at magnitude 1000,
  1.5 x 2^-14 m) -- interpolation across triangle edges, pixel-centre sampling (half a
  pixel off would be 6e-4 m here) and the elevation formula in one check.  The grid is
  asymmetric on purpose: a pixel centre EXACTLY on a shared edge can be dropped by both
  triangles (separately rounded float32 edge functions)."""
  geo = mods['camera'].ObserverGeometry(128, 32, 0.125 / 32, 0.375)
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
  tris = np.asarray(tris, dtype='int32')
  bodies = [(verts, tris, np.identity(3), np.zeros(3))]
  m = _gpu_render(mods, bodies, geo.overhead_view, geo.overhead_projection, 128, 128,
                  mods['capi'].RASTER_WALL, 0.375)
  px = 0.125 / 32
  ctr = (np.arange(128) + 0.5) * px
  inx, iny = (ctr > x0 + px) & (ctr < x1 - px), (ctr > y0 + px) & (ctr < y1 - px)
  want = c + a * ctr[:, None] + b * ctr[None, :]
  err = np.abs(m - want)[np.ix_(inx, iny)].max()
  assert err <= 1.5 * 2.0 ** -14 + 2e-6, err
  # outside the plane: ground, exactly 0
  assert np.all(m[ctr < x0 - px, :] == 0) and np.all(m[ctr > x1 + px, :] == 0)
  assert np.all(m[:, ctr < y0 - px] == 0) and np.all(m[:, ctr > y1 + px] == 0)


def test_rock_view_plane_analytic(mods):
  """Rock view (the camera under the spawned rock, observer.py:262-293, with the column
  mirror of :277): a tilted plane through the spawn point maps to object_z / 2 - (a x + b
  y) at the pixel centres x = (i + 0.5 - 16) px, y = (j + 0.5 - 16) px, to within the
  elevation formula's float32 rounding -- pins the axes and the mirror analytically."""
  geo = mods['camera'].ObserverGeometry(128, 32, 0.125 / 32, 0.375)
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
  m = _gpu_render(mods, bodies, geo.object_view(spawn, 0), geo.object_projection, 32, 32,
                  mods['capi'].RASTER_ROCK, geo.object_z)
  ctr = (np.arange(32) + 0.5 - 16) * (0.125 / 32)
  inside = np.abs(ctr) < 0.05
  want = geo.object_z / 2 - (a * ctr[:, None] + b * ctr[None, :])
  assert np.abs(m - want)[np.ix_(inside, inside)].max() <= 1.5 * 2.0 ** -14 + 2e-6


def test_drop_lands_lowest_point_on_floor(mods):
  """A rock rendered from below and max-plus-dropped on an empty floor rests
  with its lowest vertex at z = 0 (known-answer 2)."""
  from stackrl_b200 import capi
  geo = mods['camera'].ObserverGeometry(128, 32, 0.125 / 32, 0.375)
  verts, tris = mods['meshes'].synthetic_rocks(9, 1, 3, max_dimension=0.12)
  spawn = ((0., 0., 0.5), (0., 0., 0., 1.))
  bodies = [(verts[0], tris, np.identity(3), np.array(spawn[0]))]
  rock = _gpu_render(mods, bodies, geo.object_view(spawn, 0), geo.object_projection, 32, 32,
                     capi.RASTER_ROCK, geo.object_z)
  dev = torch.device('cuda')
  z = capi.drop_height_f32(torch.zeros((1, 128, 128), device=dev),
                           torch.from_numpy(rock)[None, None].to(dev),
                           torch.tensor([[0, 40, 40]], dtype=torch.int32, device=dev))
  centre_z = z.item() - geo.object_z / 2
  lowest = centre_z + float(verts[0][:, 2].min())
  # Pixel-centre sampling can miss the very tip of a spiky rock (never the other
  # way round): the tip ends at most about one pixel (3.9 mm) below the floor.
  assert -4e-3 < lowest <= 1e-4


def _fixture_bank(mods, g):
  bank = mods['meshes'].MeshBank()
  for name in g['mesh_names']:
    bank.add(g['mesh/{}/verts'.format(name)], g['mesh/{}/tris'.format(name)],
             g['mesh/{}/com'.format(name)], name=str(name))
  return bank


@pytest.mark.parametrize('drawn', [False, True])
@pytest.mark.parametrize('name,dtype,freedom', [('stack_f32', 'float32', 0),
                                                ('stack_u8', 'uint8', 0),
                                                ('test_f32_rot8', 'float32', 3),
                                                ('c1_stack_v0_30', 'uint8', 0)])
def test_batched_env_replays_reference_episode(mods, observe_golden, name, dtype, freedom, drawn):
  """BatchedStackEnv + the GPU height policy reproduce, step for step, the
  episode the UNMODIFIED reference StackEnv / TestStackEnv produced with
  Baseline('height') on the static fake backend: same observations (bitwise),
  same actions, same four rewards (IoU / OR within 1e-6, DIoU / DOR exactly --
  rows a11, a12).  ``drawn``: the rock order and the goal come from the
  environment's own seed-7 streams instead of being handed over (row a13).
  ``c1_stack_v0_30`` is BASELINE config 1: the 30-rock Stack-v0 episode (the per-step float
  maps are not in that fixture, only the packed observations)."""
  g = observe_golden
  envs = mods['envs']
  bank = _fixture_bank(mods, g)
  steps = int(g[name + '/n_steps'])
  order = [bank.names[str(n)] for n in g[name + '/urdf_order']][::-1]
  lims = g[name + '/goal_lims']
  E = 3
  # reward parameters of the golden run: rewarder='all', defaults (scale 1, no discount)
  env = envs.BatchedStackEnv(bank, E, episode_length=steps, dtype=dtype, rewarder='all',
                             orientation_freedom=freedom, seed=7)
  policy = envs.HeightPolicy()
  if drawn:
    obs, _, _ = env.reset()
    assert np.array_equal(env.goal_lims[0], lims)          # environment 0 is seeded 7 + 0
    assert [int(x) for x in env._order[0]] == order[::-1]
    # the other environments (seeds 8, 9) replay environment 0's episode from here on
    obs, _, _ = env.reset(rock_orders=[order] * E, goal_lims=[lims] * E)
  else:
    obs, _, _ = env.reset(rock_orders=[order] * E, goal_lims=[lims] * E)
  assert np.array_equal(env.goals[0].cpu().numpy(), g[name + '/goal'])
  for k in range(steps):
    key = '{}/s{}'.format(name, k)
    if key + '/overhead_map' in g:
      assert np.array_equal(env.obs.walls[1].cpu().numpy(), g[key + '/overhead_map'])
      rock = env.obs.rocks[1].cpu().numpy()
      assert np.array_equal(rock if freedom else rock[0], g[key + '/object_map'])
    for e in range(E):
      assert np.array_equal(obs[0][e].cpu().numpy(), g[key + '/obs0'])
      assert np.array_equal(obs[1][e].cpu().numpy(), g[key + '/obs1'])
    action = policy(env)
    if freedom:
      got = (int(action[0][0]), int(action[1][0]))
      assert got == tuple(int(x) for x in g[key + '/action'])
    else:
      assert int(action[0]) == int(g[key + '/action'])
    obs, reward, terminal = env.step(action)
    assert np.array_equal(env._rest[0, k], g[key + '/pose_position'])
    if freedom:
      assert np.allclose(env._quats[0, k], g[key + '/pose_orientation'], atol=1e-15)
    want = g[key + '/rewards']                       # IoU, OR, DIoU, DOR (rewarder='all')
    assert sorted(reward) == sorted(['IoU', 'OR', 'DIoU', 'DOR'])
    for e in range(E):
      np.testing.assert_allclose(float(reward['IoU'][e]), want[0], rtol=1e-6, atol=1e-9)
      np.testing.assert_allclose(float(reward['OR'][e]), want[1], rtol=1e-6, atol=1e-9)
      assert float(reward['DIoU'][e]) == np.float32(want[2])
      assert float(reward['DOR'][e]) == np.float32(want[3])
    assert bool(terminal[0]) == (k == steps - 1)
  env.check_actions()
  assert np.array_equal(obs[0][0].cpu().numpy(), g['{}/s{}/obs0'.format(name, steps)])


def test_drop_in_observer_against_recorded_maps(mods, observe_golden):
  """gpu_observer_class: the single-environment Observer API driven by a
  scene()-exposing simulator reproduces the reference Observer's maps."""
  g = observe_golden
  name = 'test_f32_rot8'
  order = [str(n) for n in g[name + '/urdf_order']]

  class Sim(object):
    new_pose = ((0., 0., 0.375 + 0.125), (0., 0., 0., 1.))
    def __init__(self):
      self.bodies, self._new = [], True
    @property
    def has_new_object(self):
      new, self._new = self._new, False
      return new
    def add(self, mesh, pos, quat):
      v, t, com = (g['mesh/{}/{}'.format(mesh, k)] for k in ('verts', 'tris', 'com'))
      rot = R.quat_matrix(quat)
      self.bodies.append((v, t, rot, np.asarray(pos) - rot.dot(com)))
      self._new = True
    def scene(self):
      return list(self.bodies)

  sim = Sim()
  Obs = mods['observer'].gpu_observer_class(object)
  obs = Obs(sim, overhead_resolution=128, object_resolution=32, pixel_size=0.125 / 32,
            max_z=0.375, orientation_freedom=3)
  placed = []
  for k in range(3):
    sim.bodies = list(placed)
    sim.add(order[k], *Sim.new_pose)                     # the spawned, unplaced rock
    obs()
    wall, rocks = obs.state
    key = '{}/s{}'.format(name, k)
    assert np.array_equal(wall, g[key + '/overhead_map'])
    assert np.array_equal(np.array(rocks), g[key + '/object_map'])
    assert obs.num_objects == 8 and obs.shape == ((128, 128), (32, 32))
    view, flat = (int(x) for x in g[key + '/action'])
    pose = obs.pose([flat // 97, flat % 97], index=view)
    assert np.array_equal(np.array(pose['position'], dtype='float64'), g[key + '/pose_position'])
    assert np.allclose(pose['orientation'], g[key + '/pose_orientation'], atol=1e-15)
    v, t, com = (g['mesh/{}/{}'.format(order[k], n)] for n in ('verts', 'tris', 'com'))
    rot = R.quat_matrix(pose['orientation'])
    placed.append((v, t, rot, np.asarray(pose['position'], dtype='float64') - rot.dot(com)))


def test_pack_and_reward_kernels(mods):
  capi = mods['capi']
  E, R_, H, W, h = 5, 3, 32, 40, 8
  walls, rocks, _ = synth.placement_batch(3, E, R_, H, W, h)
  goals = synth.goals(4, E, H, W)
  dev = torch.device('cuda')
  wd, gd, rd = (torch.from_numpy(x).to(dev) for x in (walls, goals, rocks))
  for dtype in ('float32', 'uint8'):
    wg, rk = capi.pack_obs(wd, gd, rd, dtype=dtype, scale=0.375)
    wg2, rk2 = capi.pack_obs(wd, gd, rd, dtype=dtype, scale=0.375, repeat_wall=True)
    for e in range(E):
      want = O.pack_obs_batched(walls[e], goals[e], list(rocks[e]), dtype, 0.375, 0.125)
      assert np.array_equal(wg[e].cpu().numpy(), want[0][0])
      assert np.array_equal(wg2[e].cpu().numpy(), want[0])
      assert np.array_equal(rk[e].cpu().numpy(), want[1])
  inter, uni, vol = capi.reward_sums(wd, gd, torch.full((E,), 0.25, device=dev))
  for e in range(E):
    np.testing.assert_allclose(inter[e].item(), O.intersection(walls[e], goals[e], 0.25), rtol=1e-6)
    np.testing.assert_allclose(uni[e].item(), O.union(walls[e], goals[e]), rtol=1e-6)
    np.testing.assert_allclose(vol[e].item(), goals[e].sum(), rtol=1e-6)


def test_batched_env_random_rollout_and_settle_hook(mods, observe_golden):
  """Random-goal, random-order rollout with a settle hook that nudges every rock:
  episode bookkeeping (terminal flags, rock cursor, DOR reward against a plain
  Python evaluation of rewarder.py:261-295) stays consistent."""
  g = observe_golden
  envs = mods['envs']
  bank = _fixture_bank(mods, g)
  E, steps = 6, 5
  shift = np.array([0.004, -0.002, 0.])

  def settle(mesh_ids, positions, quats):
    return positions + shift, quats

  env = envs.BatchedStackEnv(bank, E, episode_length=steps, rewarder='dor', reward_params=2,
                             reward_scale=None, settle=settle, seed=3)
  policy = envs.HeightPolicy()
  obs, r, t = env.reset()
  assert obs[0].shape == (E, 128, 128, 2) and obs[1].shape == (E, 32, 32, 1)
  assert not bool(t.any()) and env.batch_size == E
  total = np.zeros(E)
  for k in range(steps):
    obs, r, t = env.step(policy(env))
    total += r.cpu().numpy()
    assert bool(t.all()) == (k == steps - 1)
  geo = env.obs.geo
  pmax = max(geo.object_h * geo.pixel_h, geo.object_w * geo.pixel_w)
  for e in range(E):
    (u0, v0), (u1, v1) = env.goal_lims[e]
    acc = 0.
    for k in range(steps):
      p = env._rest[e, k]
      u, v = p[0] // geo.pixel_h, p[1] // geo.pixel_w
      if u0 <= u < u1 and v0 <= v < v1:
        acc += max(0., 1 - (np.linalg.norm(shift) / pmax) ** 2) * 1.
    np.testing.assert_allclose(total[e], acc / steps * steps, rtol=1e-6, atol=1e-7)
  with pytest.raises(RuntimeError):
    env.step(policy(env))                      # finished environments must be reset
  a = env.sample()
  assert a.shape == (E,) and int(a.max()) < 97 * 97


def test_rasterised_maps_take_the_fixed_point_sweep_unchanged():
  """Heightmaps from raster_kernel are multiples of 2^-14 m (observer.py:259-260
  in float32): the quantum hint HeightPolicy passes must not change a bit."""
  from stackrl_b200 import baselines, envs, meshes
  from stackrl_b200.camera import HEIGHT_QUANTUM_LOG2
  dev = torch.device('cuda')
  bank = meshes.MeshBank()
  v, t = meshes.synthetic_rocks(5, 16, 1, max_dimension=0.12)
  for k in range(16):
    bank.add(v[k], t)
  env = envs.BatchedStackEnv(bank, 64, episode_length=6, observable_size_ratio=2,
                             resolution_factor=4, dtype='float32', seed=3,
                             orientation_freedom=3, device=dev)
  policy = envs.HeightPolicy()
  env.reset()
  for _ in range(3):
    env.step(policy(env))
  walls, goals, rocks = env.planes()
  q = 2.0 ** HEIGHT_QUANTUM_LOG2
  assert bool(((walls / q) == torch.round(walls / q)).all())
  assert bool(((rocks / q) == torch.round(rocks / q)).all())
  plain = baselines.PlacementScorer('height')(walls, goals, rocks)
  hinted = baselines.PlacementScorer('height', quantum_log2=HEIGHT_QUANTUM_LOG2)(
    walls, goals, rocks)
  assert torch.equal(plain['values'], hinted['values'])
  assert torch.equal(plain['actions'], hinted['actions'])
  assert torch.equal(plain['best'], hinted['best'])


# --------------------------------------------------------------------------- #
# round 2: device-side step, both rasterisers, CUDA graph
# --------------------------------------------------------------------------- #
def _synthetic_env(mods, E, dtype='float32', freedom=0, steps=6, seed=3, **kw):
  meshes, envs = mods['meshes'], mods['envs']
  bank = meshes.MeshBank()
  v, t = meshes.synthetic_rocks(5, 16, 1, max_dimension=0.12)
  for k in range(16):
    bank.add(v[k], t, com=(0.002 * k, -0.001 * k, 0.0015 * k))
  return envs.BatchedStackEnv(bank, E, episode_length=steps, observable_size_ratio=4,
                              resolution_factor=4, dtype=dtype, seed=seed,
                              orientation_freedom=freedom, **kw)


@pytest.mark.parametrize('mode', ['rock2k', 'walls', 'big_mesh', 'many_instances'])
def test_round2_rasteriser_equals_round1_and_oracle(mods, monkeypatch, mode):
  """raster.cu (shared-memory set-up records, head-flag owner search) against the
  round-1 kernel (raster_v1.cuh, SRL_RASTER_MODE=1) and the oracle's C z-buffer:
  same bits, for 2000-triangle rocks (BASELINE config 3), multi-instance wall
  images, a mesh larger than the vertex cache and more instances than a chunk."""
  camera, capi, meshes, obs_mod = (mods[k] for k in ('camera', 'capi', 'meshes', 'observer'))
  dev = torch.device('cuda')
  if mode == 'rock2k':
    geo = camera.ObserverGeometry(128, 32, 0.005, 0.375)
    verts, tris = meshes.synthetic_rocks(4, 6, max_dimension=0.16, frequency=10)
    spawn = ((0., 0., 0.375 + 0.16), (0., 0., 0., 1.))
    scenes = [[(verts[k], tris, np.identity(3), np.array(spawn[0]))] for k in range(6)]
    view, proj, rows, cols = geo.object_view(spawn, 0), geo.object_projection, 32, 32
    rmode, zr, hint = capi.RASTER_ROCK, geo.object_z, 1002
  elif mode == 'big_mesh':
    geo = camera.ObserverGeometry(128, 32, 0.005, 0.375)
    verts, tris = meshes.synthetic_rocks(9, 2, max_dimension=0.16, frequency=16)   # 2562 verts
    assert verts.shape[1] > 2048
    spawn = ((0., 0., 0.375 + 0.16), (0., 0., 0., 1.))
    scenes = [[(verts[k], tris, np.identity(3), np.array(spawn[0]))] for k in range(2)]
    view, proj, rows, cols = geo.object_view(spawn, 0), geo.object_projection, 32, 32
    rmode, zr, hint = capi.RASTER_ROCK, geo.object_z, 0
  else:
    geo = camera.ObserverGeometry(64, 16, 0.125 / 16, 0.375)
    n = 5 if mode == 'walls' else 40
    scenes = [_random_scene(meshes, 30 + s, n, 1 if n > 8 else 2) for s in range(3)]
    view, proj, rows, cols = geo.overhead_view, geo.overhead_projection, 64, 64
    rmode, zr, hint = capi.RASTER_WALL, 0.375, 300
  flat = [b for sc in scenes for b in sc]
  v, t, inst = obs_mod._instances(flat)
  jobs, at = [], 0
  for sc in scenes:
    jobs.append(obs_mod._job(view, proj, at, len(sc), zr))
    at += len(sc)
  jobs = np.concatenate(jobs)
  vd, td = torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev)
  got = {}
  for tag, env_mode in (('r2', '0'), ('r1', '1')):
    monkeypatch.setenv('SRL_RASTER_MODE', env_mode)
    got[tag] = capi.raster(vd, td, inst, jobs, rows, cols, rmode, max_cached_verts=hint).cpu().numpy()
    got[tag + 'd'] = capi.raster(vd, td, inst, jobs, rows, cols, capi.RASTER_DEPTH,
                                 max_cached_verts=hint).cpu().numpy()
  assert np.array_equal(got['r2d'], got['r1d'])
  assert np.array_equal(got['r2'], got['r1'])
  for k, sc in enumerate(scenes):
    want = R.render_depth(view, proj, rows, cols, sc)
    assert np.array_equal(got['r2d'][k], want)
    assert (want < 1).sum() > 20
  # the device-side instance-count override draws a prefix of each job's instances
  if mode == 'walls':
    monkeypatch.setenv('SRL_RASTER_MODE', '0')
    counts = torch.tensor([2, 0, 5], dtype=torch.int32, device=dev)
    part = capi.raster(vd, td, inst, jobs, rows, cols, capi.RASTER_DEPTH, inst_counts=counts,
                       max_cached_verts=hint).cpu().numpy()
    for k, c in enumerate([2, 0, 5]):
      assert np.array_equal(part[k], R.render_depth(view, proj, rows, cols, scenes[k][:c]))


def test_place_poses_kernel_and_invalid_actions(mods):
  """srl_place_poses_f32 = Observer.pose (observer.py:392-421) per environment, on
  strided action columns; out-of-range actions give NaN poses and a status flag
  instead of reading out of bounds (env.py:237 asserts)."""
  capi = mods['capi']
  dev = torch.device('cuda')
  E, R_, H, W, h = 9, 4, 40, 32, 8
  walls, rocks, _ = synth.placement_batch(8, E, R_, H, W, h)
  rocks = np.where(rocks < 2e-4, rocks * 0.3, rocks).astype('float32')   # cells under 1e-4
  rng = np.random.default_rng(1)
  Ph, Pw = H - h + 1, W - h + 1
  best = np.stack([rng.integers(0, R_, E), rng.integers(0, Ph * Pw, E)], 1).astype('int64')
  best[2] = (R_, 5)            # view out of range
  best[4] = (0, Ph * Pw)       # position out of range
  best[6] = (-1, 3)
  bd = torch.from_numpy(best).to(dev)
  orient = np.random.default_rng(2).normal(size=(R_, 4))
  geom = (0.01, 0.0125, 0.08, 0.1, 0.1)
  poses, status = capi.place_poses(torch.from_numpy(walls).to(dev), torch.from_numpy(rocks).to(dev),
                                   bd[:, 0], bd[:, 1], torch.from_numpy(orient).to(dev), geom)
  poses, status = poses.cpu().numpy(), status.cpu().numpy()
  assert list(status) == [0, 0, 1, 0, 1, 0, 1, 0, 0]
  for e in range(E):
    if status[e]:
      assert np.isnan(poses[e]).all()
      continue
    r, a = best[e]
    i, j = a // Pw, a % Pw
    n = rocks[e, r]
    z = (walls[e, i:i + h, j:j + h] + n)[n > np.float32(1e-4)].max()
    x = i * geom[0] + geom[2] / 2
    y = j * geom[1] + geom[3] / 2
    z = z - np.float32(geom[4] / 2)
    assert np.array_equal(poses[e], np.array([x, y, z, *orient[r]], dtype='float64'))


@pytest.mark.parametrize('dtype,freedom', [('float32', 0), ('uint8', 0), ('float32', 2)])
def test_step_graph_replays_the_eager_chain(mods, dtype, freedom):
  """capture(policy): policy + step as ONE CUDA graph gives the same observations,
  rewards, terminals and device state as the eager kernel chain."""
  envs = mods['envs']
  E, steps = 40, 5
  a = _synthetic_env(mods, E, dtype, freedom, steps)
  b = _synthetic_env(mods, E, dtype, freedom, steps)
  pa, pb = envs.HeightPolicy(), envs.HeightPolicy()
  a.reset()
  b.reset()
  oa, ra, ta = a.step(pa(a))                       # one eager step each (lazy attributes)
  ob, rb, tb = b.step(pb(b))
  b.capture(pb)
  for k in range(1, steps):
    oa, ra, ta = a.step(pa(a))
    ob, rb, tb = b.step_policy()
    assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1])
    assert torch.equal(ra, rb) and torch.equal(ta, tb)
    assert bool(ta.all()) == (k == steps - 1)
  assert torch.equal(a.obs.state.hist_rest, b.obs.state.hist_rest)
  assert torch.equal(a.obs.state.counts, b.obs.state.counts)
  a.check_actions()
  b.check_actions()
  with pytest.raises(RuntimeError):
    b.step_policy()
  # a new episode replays through the same graph
  a.reset()
  b.reset()
  oa, ra, ta = a.step(pa(a))
  ob, rb, tb = b.step_policy()
  assert torch.equal(oa[0], ob[0]) and torch.equal(ra, rb)
  # graph without a policy: the action is copied into the captured buffers
  c = _synthetic_env(mods, E, dtype, freedom, steps)
  d = _synthetic_env(mods, E, dtype, freedom, steps)
  c.reset()
  d.reset()
  act = pa(c)
  oc, rc, _ = c.step(act)
  od, rd, _ = d.step(act)
  d.capture()
  act = pa(c)
  oc, rc, _ = c.step(act)
  od, rd, _ = d.step(act)
  assert torch.equal(oc[0], od[0]) and torch.equal(oc[1], od[1]) and torch.equal(rc, rd)


@pytest.mark.parametrize('dtype,freedom', [('float32', 0), ('uint8', 0), ('float32', 2)])
def test_persistent_observation_equals_fresh_packing(mods, dtype, freedom):
  """persistent_observation=True: step() rewrites only the wall rows the appended rock
  changed (srl_raster_incremental_rows -> srl_pack_rewards_rows_f32) in fixed buffers;
  the observation equals the freshly packed one at every step, across partial resets
  (new goals: every row stale) and new episodes."""
  E, steps = 24, 6
  a = _synthetic_env(mods, E, dtype, freedom, steps, rewarder='all')
  b = _synthetic_env(mods, E, dtype, freedom, steps, rewarder='all',
                     persistent_observation=True)
  a.reset()
  b.reset()
  first = None
  for k in range(steps - 1):
    action = a.sample()
    (oa, ra, ta), (ob, rb, tb) = a.step(action), b.step(action)
    assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1])
    for name in ra:
      assert torch.equal(ra[name], rb[name])
    if first is None:
      first = ob[0]
    assert ob[0].data_ptr() == first.data_ptr()          # the same buffer every step
    rows = b.obs.wall_rows.cpu().numpy()
    assert (rows[:, 0] >= 0).all() and (rows[:, 1] <= a.obs.geo.overhead_h).all()
    assert (rows[:, 1] - rows[:, 0] < a.obs.geo.overhead_h).all()     # a rock, not the scene
    if k == 1:
      for env in (a, b):
        env.reset(env_ids=[1, 5, E - 1])
  a.reset()
  b.reset()
  action = a.sample()
  (oa, _, _), (ob, _, _) = a.step(action), b.step(action)
  assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1])


def test_uint8_policy_scores_the_packed_observation(mods):
  """The planar uint8 cast HeightPolicy scores equals the channels of the packed
  uint8 observation (env.py:171-178), and its level is the quantised goal height."""
  env = _synthetic_env(mods, 12, 'uint8', 1)
  env.reset()
  env.step(mods['envs'].HeightPolicy()(env))
  w8, g8, r8 = env.planes_u8()
  wall_goal, rock = env.observation
  assert torch.equal(w8, wall_goal[:, 0, ..., 0]) and torch.equal(g8, wall_goal[:, 0, ..., 1])
  assert torch.equal(r8, rock[..., 0])
  assert torch.equal(env._level8_d, g8.amax(dim=(1, 2)))


def test_discounted_rewards_and_settle_hook_moving_every_rock(mods):
  """DOR / DIoU (rewarder.py:261-295) with position and orientation discounts when
  the settle hook moves the new rock AND, through the all-poses return, the rocks
  placed earlier; checked against a plain Python evaluation of the reference's loop."""
  envs = mods['envs']
  E, steps = 5, 4
  rng = np.random.default_rng(0)
  log = {'rest': [[] for _ in range(E)], 'placed': [[] for _ in range(E)]}

  def settle(mesh_ids, positions, quats):
    d = rng.normal(scale=0.004, size=positions.shape)
    q = quats + rng.normal(scale=0.02, size=quats.shape)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    new = np.concatenate([positions + d, q], axis=1)
    for e in range(E):
      log['placed'][e].append(np.concatenate([positions[e], quats[e]]))
      # earlier rocks drift a little too
      log['rest'][e] = [p + np.r_[rng.normal(scale=0.001, size=3), np.zeros(4)]
                        for p in log['rest'][e]] + [new[e]]
    allp = np.stack([np.stack(log['rest'][e]) for e in range(E)])
    return new[:, :3], new[:, 3:], allp

  for metric in ('dor', 'diou'):
    for e in range(E):
      log['rest'][e], log['placed'][e] = [], []
    env = _synthetic_env(mods, E, steps=steps, rewarder=metric, reward_params=(2, 1),
                         reward_scale=None, settle=settle, freedom=1)
    policy = envs.HeightPolicy()
    env.reset()
    geo = env.obs.geo
    pmax = max(geo.object_h * geo.pixel_h, geo.object_w * geo.pixel_w)
    memory = np.zeros(E)
    for k in range(steps):
      _, r, _ = env.step(policy(env))
      r = r.cpu().numpy()
      hist = env.obs.state.hist_rest.cpu().numpy()
      for e in range(E):
        assert np.array_equal(hist[e, :k + 1], np.stack(log['rest'][e]))
        (u0, v0), (u1, v1) = env.goal_lims[e]
        acc, n_out = 0., 0
        for rest, placed in zip(log['rest'][e], log['placed'][e]):
          u, v = float(rest[0]) // geo.pixel_h, float(rest[1]) // geo.pixel_w
          if u0 <= u < u1 and v0 <= v < v1:
            perr = np.linalg.norm(np.subtract(placed[:3], rest[:3]))
            oerr = 2 * np.arccos(min(float(np.dot(placed[3:], rest[3:])), 1.))
            acc += max(0., 1 - (perr / pmax) ** 2) * max(0., 1 - (oerr / np.pi) ** 1)
          else:
            n_out += 1
        value = acc / steps if metric == 'dor' else acc / (steps + n_out)
        np.testing.assert_allclose(r[e], (value - memory[e]) * steps, rtol=1e-6, atol=1e-7)
        memory[e] = value
    # the wall image shows the drifted rocks: re-rasterising the history gives it back
    walls = env.obs.walls.clone()
    env.obs.observe_walls()
    assert torch.equal(walls, env.obs.walls)


def test_persistent_observation_with_a_settle_hook(mods):
  """A settle hook that moves the new rock only (appended image, a few rows rewritten) or
  every rock (whole-scene redraw: srl_raster_incremental_rows reports all rows): the
  persistent observation still equals the freshly packed one."""
  E, steps = 6, 4
  for move_all in (False, True):
    envs_ = []
    for persistent in (False, True):
      rng = np.random.default_rng(3)
      rest = [[] for _ in range(E)]

      def settle(mesh_ids, positions, quats, rng=rng, rest=rest):
        new = np.concatenate([positions + rng.normal(scale=0.003, size=positions.shape), quats], 1)
        for e in range(E):
          rest[e] = [p + (np.r_[rng.normal(scale=0.001, size=3), np.zeros(4)] if move_all
                          else 0.) for p in rest[e]] + [new[e]]
        if move_all:
          return new[:, :3], new[:, 3:], np.stack([np.stack(rest[e]) for e in range(E)])
        return new[:, :3], new[:, 3:]

      env = _synthetic_env(mods, E, steps=steps, settle=settle,
                           persistent_observation=persistent)
      env.reset()
      envs_.append(env)
    a, b = envs_
    for _ in range(steps):
      action = a.sample()
      (oa, ra, _), (ob, rb, _) = a.step(action), b.step(action)
      assert torch.equal(oa[0], ob[0]) and torch.equal(oa[1], ob[1]) and torch.equal(ra, rb)
      rows = b.obs.wall_rows.cpu().numpy()
      H = a.obs.geo.overhead_h
      assert ((rows[:, 1] - rows[:, 0] == H).all()) == (move_all and len(rows) > 0)


def test_incremental_wall_image_equals_the_full_redraw(mods):
  """A step draws only the appended rock onto the kept depth image
  (srl_raster_incremental); re-drawing every placed rock gives the same bits."""
  envs = mods['envs']
  for dtype, freedom in (('float32', 0), ('uint8', 2)):
    env = _synthetic_env(mods, 33, dtype, freedom, steps=7)
    policy = envs.HeightPolicy()
    env.reset()
    for k in range(7):
      env.step(policy(env))
      walls = env.obs.walls.clone()
      depth = env.obs._wall_depth.clone()
      env.obs._wall_depth_valid = False
      env.obs.observe_walls()                      # whole scene
      assert torch.equal(walls, env.obs.walls), (dtype, k)
      assert torch.equal(depth, env.obs._wall_depth)
      assert int(env.obs.counts[0]) == k + 1
    # a new episode starts from an empty image
    env.reset()
    assert float(env.obs.walls.abs().max()) == 0.


def test_device_side_episode_draws(mods):
  """vector_rng=True: rock lists and goal rectangles drawn by srl_env_draw obey the
  constraints of env.py:268-272 / rewarder.py:211-250 and change per episode."""
  for ratio in (.25, (.5, .25), None):
    env = _synthetic_env(mods, 500, steps=6, vector_rng=True, goal_size_ratio=ratio, seed=9)
    env.reset()
    lims, order = env.goal_lims.copy(), env._order.copy()
    h = lims[:, 1, 0] - lims[:, 0, 0]
    w = lims[:, 1, 1] - lims[:, 0, 1]
    assert (h >= 16).all() and (w >= 16).all() and (lims[:, 1] <= 64).all() and (lims >= 0).all()
    u_max, v_max = 64 - h, 64 - w
    assert (lims[:, 0, 0] >= u_max // 8).all() and (lims[:, 0, 0] <= 7 * u_max // 8).all()
    assert (lims[:, 0, 1] >= v_max // 8).all() and (lims[:, 0, 1] <= 7 * v_max // 8).all()
    if ratio == .25:
      assert (np.abs(h * w - 1024) <= np.maximum(h, w)).all()
      assert len(np.unique(h)) > 5                      # the beta draw spreads the aspect
    if ratio == (.5, .25):
      assert set(zip(h.tolist(), w.tolist())) == {(32, 16), (16, 32)}
    if ratio is None:
      assert (h == 16).all() and len(np.unique(w)) > 5  # quirk Q13
    srt = np.sort(order, axis=1)
    assert not (srt[:, 1:] == srt[:, :-1]).any()        # 16 rocks >= 6 per episode
    assert order.min() >= 0 and order.max() < 16 and len(np.unique(order[:, 0])) == 16
    goal = env.goals.cpu().numpy()
    for e in (0, 17, 499):
      (u0, v0), (u1, v1) = lims[e]
      want = np.zeros((64, 64), 'float32')
      want[u0:u1, v0:v1] = np.float32(0.25)
      assert np.array_equal(goal[e], want)
    env.step(mods['envs'].HeightPolicy()(env))
    env.reset(env_ids=[3, 4])                           # partial reset: a new draw for two
    lims2, order2 = env.goal_lims, env._order
    keep = np.ones(500, bool)
    keep[[3, 4]] = False
    assert np.array_equal(lims2[keep], lims[keep]) and np.array_equal(order2[keep], order[keep])
    assert not np.array_equal(order2[[3, 4]], order[[3, 4]])
    assert env.obs.counts.cpu().numpy()[[3, 4, 5]].tolist() == [0, 0, 1]
  # a bank smaller than the episode: drawn with replacement
  from stackrl_b200 import capi
  env = _synthetic_env(mods, 64, steps=40, vector_rng=True, seed=2)
  env.reset()
  assert env._order.max() < 16 and env._order.shape == (64, 40)


@pytest.mark.parametrize('dtype,freedom', [('float32', 0), ('uint8', 0), ('float32', 2)])
def test_pack_rewards_goal_rectangle_equals_goal_plane(mods, dtype, freedom):
  """srl_pack_rewards_f32 with goals == NULL derives the goal map from the goal
  limits (what the environment step does): same observation and rewards, bit for bit,
  as reading the filled goal plane."""
  from stackrl_b200 import capi
  outs = []
  for plane in (False, True):
    env = _synthetic_env(mods, 12, dtype=dtype, freedom=freedom, steps=4, rewarder='all')
    env.reset()
    for _ in range(2):
      env.step(env.sample())
    g = env.obs.geo
    wall_goal, rock, r = capi.pack_rewards(
      env.obs.state, env.obs.walls, env.goals if plane else None, env.obs.rocks, env._goal_z_d,
      env._rects_d, env.metric, env.scale, (g.pixel_h, g.pixel_w), env._pmax, env._pexp,
      env._oexp, dtype=env._dtype, obs_scale=env._scale, repeat_wall=env.R > 1)
    outs.append((wall_goal, rock, r))
  for a, b in zip(*outs):
    assert torch.equal(a, b)
  assert outs[0][0][..., 1].any()


@pytest.mark.parametrize('dtype,freedom', [('float32', 0), ('uint8', 2)])
def test_cached_rock_images_equal_the_per_step_rasterisation(mods, dtype, freedom):
  """The images of the spawned rocks fetched from the per-bank table (rasterised once,
  srl_gather_rows_f32) are the images rasterised every step: same observations, rewards
  and rock planes over an episode, bit for bit."""
  a = _synthetic_env(mods, 20, dtype=dtype, freedom=freedom, steps=5)
  b = _synthetic_env(mods, 20, dtype=dtype, freedom=freedom, steps=5, rock_cache_bytes=0)
  assert a.obs._cache_rocks and not b.obs._cache_rocks
  oa, ob = a.reset()[0], b.reset()[0]
  for _ in range(5):
    for x, y in zip(oa, ob):
      assert torch.equal(x, y)
    assert torch.equal(a.obs.rocks, b.obs.rocks) and bool((a.obs.rocks > 0).any())
    action = a.sample()
    (oa, ra, ta), (ob, rb, tb) = a.step(action), b.step(action)
    assert torch.equal(ra, rb) and torch.equal(ta, tb)
  assert a.obs._rock_cache is not None and b.obs._rock_cache is None


def test_gather_rows_out_of_range_index_is_nan(mods):
  from stackrl_b200 import capi
  dev = torch.device('cuda')
  for row in (7, 64, 1024, 3000):
    table = torch.arange(5 * row, dtype=torch.float32, device=dev).view(5, row)
    index = torch.tensor([4, 0, -1, 2, 5, 3], dtype=torch.int32, device=dev)
    out = torch.zeros((6, row), dtype=torch.float32, device=dev)
    capi.gather_rows(table, index, out)
    for k, src in enumerate(index.tolist()):
      if 0 <= src < 5:
        assert torch.equal(out[k], table[src])
      else:
        assert bool(torch.isnan(out[k]).all())


def test_parallel_env_contract_block_call_seed(mods):
  """ParallelEnv's surface (utils.py:393-428, 520-536): ``block=False`` returns a callable
  that delivers the same time step, ``env(action)`` is ``step``, ``seed`` returns one
  ``[seed_i, goal_seed_i]`` per environment with the reference's derivation."""
  a = _synthetic_env(mods, 6, steps=4, seed=11)
  b = _synthetic_env(mods, 6, steps=4, seed=11, block=False)
  ra, rb = a.reset(), b.reset()
  assert callable(rb) and not callable(ra)
  rb = rb()
  for _ in range(3):
    for x, y in zip(ra[0], rb[0]):
      assert torch.equal(x, y)
    assert torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2])
    action = a.sample()
    ra, rb = a(action), b.step(action)()
  assert not callable(b.step(b.sample(), block=True))
  seeds = a.seed(40)
  assert len(seeds) == 6
  for i, (s0, s1) in enumerate(seeds):
    assert s0 == 40 + i
    assert s1 == int(np.random.RandomState(40 + i).randint(2 ** 32))
  a.close()


def test_contact_precheck_matches_oracle(mods):
  """contact cells / octants / support verdict of the chosen
  placements against the numpy restatement (exact: counts and integer octants)."""
  capi = mods['capi']
  dev = torch.device('cuda')
  E, R_, H, W, h = 40, 3, 48, 40, 16
  walls, rocks, _ = synth.placement_batch(12, E, R_, H, W, h)
  q = np.float32(2.0 ** -14)
  walls = (np.round(walls / q) * q).astype('float32')       # heights as the rasteriser leaves them
  rocks = (np.round(rocks / q) * q).astype('float32')
  walls[:8] = 0.                                            # flat floor: wide contact patch
  walls[8:12, 20:, :] += np.float32(0.05)                   # a step: contacts on one side only
  rng = np.random.default_rng(3)
  Ph, Pw = H - h + 1, W - h + 1
  best = np.stack([rng.integers(0, R_, E), rng.integers(0, Ph * Pw, E)], 1).astype('int64')
  best[-1] = (R_, 0)                                        # invalid action
  bd = torch.from_numpy(best).to(dev)
  for eps in (0., 2.0 ** -13, 2e-3):
    c, o, s = capi.contact_precheck(torch.from_numpy(walls).to(dev), torch.from_numpy(rocks).to(dev),
                                    bd[:, 0], bd[:, 1], eps=eps)
    c, o, s = c.cpu().numpy(), o.cpu().numpy(), s.cpu().numpy()
    assert (c[-1], o[-1], bool(s[-1])) == (0, 0, False)
    for e in range(E - 1):
      r, a = best[e]
      want = O.contact_precheck(walls[e], rocks[e, r], (a // Pw, a % Pw), eps=eps)
      assert (int(c[e]), int(o[e]), bool(s[e])) == want, (eps, e)
    assert c[:-1].min() >= 1                                # the maximum itself always touches
  assert s[:8].all()                                        # flat floor under a convex underside


def test_gpu_camera_returns_the_recorded_depth_images(mods, observe_golden):
  """Seam b2: GpuCamera.getCameraImage(width, height, viewMatrix, projectionMatrix) --
  the call the reference's unmodified Observer makes (observer.py:252-257, 267-272) --
  returns, bit for bit, the depth images the reference env recorded on the fake
  backend (overhead camera first, then the eight object views)."""
  g = observe_golden
  name = 'test_f32_rot8'
  order = [str(n) for n in g[name + '/urdf_order']]
  spawn = ((0., 0., 0.375 + 0.125), (0., 0., 0., 1.))
  geo = mods['camera'].ObserverGeometry(128, 32, 0.125 / 32, 0.375, orientation_freedom=3)

  class Sim(object):
    def __init__(self):
      self.bodies = []
    def scene(self):
      return list(self.bodies)

  def body(mesh, pos, quat):
    v, t, com = (g['mesh/{}/{}'.format(mesh, k)] for k in ('verts', 'tris', 'com'))
    rot = R.quat_matrix(quat)
    return (v, t, rot, np.asarray(pos, dtype='float64') - rot.dot(com))

  sim = Sim()
  cam = mods['observer'].GpuCamera(sim)
  # the reference Simulator's floor (simulator.py:167-179): a 20 m x 20 m box of height 0
  from oracle.fake_pybullet import box_mesh
  fv, ft = box_mesh((10, 10, 0))
  placed = [(fv, ft, np.eye(3), np.zeros(3))]
  for k in range(3):
    key = '{}/s{}'.format(name, k)
    depths = split_depths(g, key)
    assert len(depths) == 9
    sim.bodies = placed + [body(order[k], *spawn)]
    w, h, rgb, depth, seg = cam.getCameraImage(
      width=128, height=128, viewMatrix=geo.overhead_view,
      projectionMatrix=geo.overhead_projection)
    assert (w, h, rgb, seg) == (128, 128, None, None)
    assert depth.dtype == np.float32 and np.array_equal(depth, depths[0])
    for r in range(8):
      _, _, _, depth, _ = cam.getCameraImage(
        width=32, height=32, viewMatrix=geo.object_view(spawn, r),
        projectionMatrix=geo.object_projection)
      assert np.array_equal(depth, depths[1 + r])
    placed.append(body(order[k], g[key + '/pose_position'], g[key + '/pose_orientation']))


@pytest.mark.parametrize('mode_name', ['RASTER_DEPTH', 'RASTER_WALL', 'RASTER_ROCK'])
def test_warp_per_image_kernel_equals_cta_kernel_and_redraw(mods, monkeypatch, mode_name):
  """The in-place incremental image of a small mesh is drawn by one warp per image into a
  window of the image (raster_warp_kernel).  Same state, image and reported rows as the
  CTA-per-image kernel (SRL_RASTER_WARP=0) and as re-drawing the whole scene, and the
  depth state is the oracle's -- for small rocks, a slab wider than one window pass (several
  passes by rows and columns), a rock half outside the image, a mesh too big for the warp's
  vertex cache (uncached path), a vertex far outside the image (window = whole image),
  jobs without instances, and a job count that does not fill the last CTA."""
  capi, obs_mod, meshes = mods['capi'], mods['observer'], mods['meshes']
  mode = getattr(capi, mode_name)
  geo = mods['camera'].ObserverGeometry(96, 32, 0.125 / 32, 0.375)
  rows = cols = 96
  rng = np.random.default_rng(11)
  small_v, small_t = meshes.synthetic_rocks(3, 40, 1, max_dimension=0.06)     # 42 vertices
  big_v, big_t = meshes.synthetic_rocks(4, 2, 3, max_dimension=0.1)           # 642 vertices
  box_t = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1],
                    [1, 5, 6], [1, 6, 2], [2, 6, 7], [2, 7, 3], [3, 7, 4], [3, 4, 0]], 'int32')
  def box(hx, hy, hz):
    return np.array([[sx * hx, sy * hy, sz * hz] for sz in (-1, 1)
                     for sx, sy in ((-1, -1), (1, -1), (1, 1), (-1, 1))], 'float32')
  def pose(lo=0.06, hi=0.31):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return R.quat_matrix(q), np.array([rng.uniform(lo, hi), rng.uniform(lo, hi),
                                       rng.uniform(0.03, 0.3)])
  scenes = []
  for k in range(37):
    n = int(rng.integers(0, 5))
    bodies = [(small_v[int(rng.integers(40))], small_t) + pose() for _ in range(n)]
    scenes.append(bodies)
  eye = np.identity(3)
  scenes[3] = scenes[3][:1] + [(box(0.16, 0.15, 0.01), box_t, eye, np.array([.19, .19, .2]))]
  scenes[5] = [(box(0.02, 0.17, 0.02), box_t) + (pose()[0], np.array([.19, .19, .1]))]
  scenes[8] = scenes[8][:2] + [(small_v[0], small_t, eye, np.array([0.37, 0.005, 0.1]))]
  scenes[13] = scenes[13][:1] + [(big_v[1], big_t) + pose()]
  far_vertex = small_v[1].copy()
  far_vertex[0] = [5.0e8, 0., 0.01]                       # projects beyond |1e9| px
  scenes[21] = [(far_vertex, small_t, eye, np.array([.2, .2, .1]))]
  scenes[30] = []
  bodies_all, begin = [], []
  for s in scenes:
    begin.append(len(bodies_all))
    bodies_all += s
  verts, tris, inst = obs_mod._instances(bodies_all)
  jobs = np.zeros(len(scenes), dtype=capi.JOB_DTYPE)
  jobs['view'], jobs['proj'] = geo.overhead_view, geo.overhead_projection
  jobs['inst_begin'] = begin
  jobs['inst_count'] = [len(s) for s in scenes]
  jobs['zrange'] = 0.375
  dev = torch.device('cuda')
  verts_d, tris_d = torch.from_numpy(verts).to(dev), torch.from_numpy(tris).to(dev)
  counts = torch.tensor([len(s) for s in scenes], dtype=torch.int32, device=dev)
  before = (counts - 1).clamp(min=0)

  def run(inst_counts, only_last, state, out, rows_out, hint):
    capi.raster(verts_d, tris_d, inst, jobs, rows, cols, mode, out=out, inst_counts=inst_counts,
                depth_state=state, only_last=only_last, rows_out=rows_out, max_cached_verts=hint)
    torch.cuda.synchronize()
  shape = (len(scenes), rows, cols)
  state0 = torch.empty(shape, dtype=torch.float32, device=dev)
  out0 = torch.empty(shape, dtype=torch.float32, device=dev)
  rows0 = torch.zeros((len(scenes), 2), dtype=torch.int32, device=dev)
  run(before, 0, state0, out0, rows0, 2048)               # every instance but the last
  results = []
  for warp in ('1', '0'):
    monkeypatch.setenv('SRL_RASTER_WARP', warp)
    state, out, r = state0.clone(), out0.clone(), torch.zeros_like(rows0)
    run(counts, 2, state, out, r, 64)
    results.append((state, out, r))
  monkeypatch.delenv('SRL_RASTER_WARP')
  for a, b in zip(results[0], results[1]):
    assert torch.equal(a, b)
  full_state, full_out = torch.empty_like(state0), torch.empty_like(out0)
  run(counts, 0, full_state, full_out, torch.zeros_like(rows0), 2048)
  assert torch.equal(results[0][0], full_state)
  assert torch.equal(results[0][1], full_out)
  changed = (full_state != state0).any(dim=2).cpu().numpy()
  r = results[0][2].cpu().numpy()
  for k in range(len(scenes)):
    idx = np.nonzero(changed[k])[0]
    if len(idx):
      assert r[k, 0] <= idx[0] and idx[-1] < r[k, 1], k
    if not scenes[k]:
      assert r[k, 0] >= r[k, 1]
  assert tuple(r[21]) == (0, rows) and r[3, 1] - r[3, 0] > 70 and changed.sum() > 300
  if mode == capi.RASTER_DEPTH:
    for k in (3, 5, 8, 13, 17, 21, 30):
      want = R.render_depth(geo.overhead_view, geo.overhead_projection, rows, cols, scenes[k])
      assert np.array_equal(full_state[k].cpu().numpy(), want), k


@pytest.mark.parametrize('scale', [0.375, 0.125, 0.16, 0.3, 1.0, 0.0078125])
def test_quantise_planes_is_the_exact_float32_cast(mods, scale):
  """srl_quantise_planes_u8 hoists the reciprocal part of the float32 division out of the
  per-pixel work; the bytes are those of numpy's `np.array(x * 255 / scale, 'uint8')` in
  float32 (env.py:171-178), also where x * 255 / scale lies within a few ulps of an integer,
  for zeros, subnormals and a value that needs the general division."""
  capi = mods['capi']
  s32 = np.float32(scale)
  k = np.arange(256, dtype='float64')
  exact = (k * float(s32) / 255.0).astype('float32')                # quotient ~ an integer
  near = np.concatenate([np.nextafter(exact, np.float32(np.inf), dtype='float32'),
                         np.nextafter(exact, np.float32(-np.inf), dtype='float32'), exact])
  for _ in range(3):
    near = np.concatenate([near, np.nextafter(near, np.float32(np.inf), dtype='float32'),
                           np.nextafter(near, np.float32(-np.inf), dtype='float32')])
  rng = np.random.default_rng(4)
  vals = np.concatenate([near, rng.uniform(0, float(s32), 20000).astype('float32'),
                         np.array([0., 1e-42, 1e-38, 3e-33, float(s32)], 'float32')])
  vals = vals[(vals >= 0) & (vals <= s32)]
  H = W = 64
  n = (len(vals) // (H * W)) * H * W
  walls = vals[:n].reshape(-1, H, W)
  E = len(walls)
  goals = np.ascontiguousarray(walls[::-1])
  rocks = np.ascontiguousarray(walls[:, :16, :16].reshape(E, 1, 16, 16))
  dev = torch.device('cuda')
  got = capi.quantise_planes(*(torch.from_numpy(x).to(dev) for x in (walls, goals, rocks)),
                             float(s32))
  with np.errstate(over='ignore'):
    want = [np.array(x * np.float32(255) / s32, dtype='uint8') for x in (walls, goals, rocks)]
  for g, w in zip(got, want):
    assert np.array_equal(g.cpu().numpy(), w)
