"""Host side of the observer seams, no GPU needed.

* a1: ObserverGeometry's camera matrices equal what the UNMODIFIED reference
  Observer asks its simulator for (observer.py:84-141) -- recorded from the fake
  pybullet's computeViewMatrix / computeProjectionMatrix (GL look-at / glFrustum).
* b1: ``gpu_observer_class(reference Observer)`` constructs without a CUDA device
  and passes Rewarder's isinstance gate (rewarder.py:52-63), also inside the
  reference StackEnv (env.py:138-168).
* b2: ``gpu_simulator_class(reference Simulator)`` is a Simulator (same gate).
* ``PybulletScene`` turns pybullet's scene queries into the rasteriser's instances.
"""
import os

import numpy as np
import pytest

from oracle import fake_pybullet, refload
from stackrl_b200 import camera, observer


def _urdfs(ns, names=('0_0', '50_000', '75_003')):
  return [os.path.join(ns.root, 'stackrl/envs/data/generated', n + '.urdf') for n in names]


@pytest.fixture(scope='module')
def ref():
  fb = fake_pybullet.FakeBullet()
  return refload.load(pybullet=fb), fb


@pytest.mark.needs_reference
@pytest.mark.parametrize('freedom', [0, 3])
def test_camera_matrices_equal_the_reference_observers(ref, freedom):
  ns, fb = ref
  sim = ns.simulator.Simulator(spawn_position=[0, 0, 0.5], spawn_orientation=[0, 0, 0, 1])
  sim.connect() if hasattr(sim, 'connect') else None
  kw = dict(overhead_resolution=128, object_resolution=32, pixel_size=0.125 / 32, max_z=0.375)
  calls = []
  real = fb.getCameraImage

  def spy(width, height, viewMatrix, projectionMatrix, **_):
    calls.append((width, height, tuple(viewMatrix), tuple(projectionMatrix)))
    return real(width, height, viewMatrix, projectionMatrix)
  fb.getCameraImage = spy
  try:
    sim.reset(_urdfs(ns)[1])
    obs = ns.observer.Observer(sim, orientation_freedom=freedom, **kw)
    obs()
  finally:
    fb.getCameraImage = real
  geo = camera.ObserverGeometry(128, 32, 0.125 / 32, 0.375, freedom)
  assert len(calls) == 1 + 2 ** freedom
  w, h, view, proj = calls[0]
  assert (w, h) == (128, 128)
  np.testing.assert_allclose(view, geo.overhead_view, rtol=0, atol=1e-12)
  np.testing.assert_allclose(proj, geo.overhead_projection, rtol=1e-15, atol=0)
  for k in range(2 ** freedom):
    w, h, view, proj = calls[1 + k]
    assert (w, h) == (32, 32)
    np.testing.assert_allclose(view, geo.object_view(sim.new_pose, k), rtol=0, atol=1e-9)
    np.testing.assert_allclose(proj, geo.object_projection, rtol=1e-15, atol=0)
  if freedom:
    np.testing.assert_allclose(np.array(obs._object_orientations), np.array(geo.orientations),
                               atol=1e-15)
  assert obs.shape == geo.shape and obs.size == geo.size and obs.max_z == geo.max_z


@pytest.mark.needs_reference
def test_gpu_observer_passes_the_rewarder_gate_without_a_device(ref):
  """b1: a GpuObserver derived from the reference Observer is accepted where the
  reference checks ``isinstance(observer, Observer)`` (rewarder.py:58-63)."""
  ns, fb = ref
  Obs = observer.gpu_observer_class(ns.observer.Observer)
  sim = ns.simulator.Simulator(spawn_position=[0, 0, 0.5], spawn_orientation=[0, 0, 0, 1])
  obs = Obs(sim, overhead_resolution=128, object_resolution=32, pixel_size=0.125 / 32,
            max_z=0.375)
  assert isinstance(obs, ns.observer.Observer)
  rew = ns.rewarder.Rewarder(simulator=sim, observer=obs, metric='iou', goal_size_ratio=.25,
                             n_objects=6, seed=3)
  rew.reset()                              # draws a goal from the observer's shapes
  assert rew.goal.shape == (128, 128) and rew._goal_z == obs.max_z == 0.25
  # a plain-object observer is refused by the same gate
  plain = observer.gpu_observer_class(object)(sim, 128, 32, 0.125 / 32, 0.375)
  with pytest.raises(TypeError):
    ns.rewarder.Rewarder(simulator=sim, observer=plain)
  # the reference env builds around it (the observer is only *called* at reset)
  env = ns.env.StackEnv(urdfs=_urdfs(ns), observer=Obs, seed=1, episode_length=3)
  assert isinstance(env._obs, Obs) and env._obs.shape == ((128, 128), (32, 32))
  assert env.action_space.n == 97 * 97
  env.close()


@pytest.mark.needs_reference
def test_gpu_simulator_is_a_reference_simulator(ref):
  """b2: the CUDA camera as a Simulator subclass keeps both isinstance gates and the
  pybullet forwarder for everything but getCameraImage."""
  ns, fb = ref
  Sim = observer.gpu_simulator_class(ns.simulator.Simulator)
  env = ns.env.StackEnv(urdfs=_urdfs(ns), simulator=Sim, seed=1, episode_length=3)
  assert isinstance(env._sim, ns.simulator.Simulator)
  assert type(env._sim).getCameraImage is not None
  assert 'getCameraImage' in type(env._sim).__dict__            # not the pb forwarder
  assert env._sim.computeViewMatrix.func == fb.computeViewMatrix  # still forwarded
  env.close()


def test_pybullet_scene_equals_the_fake_servers_scene():
  fb = fake_pybullet.FakeBullet()
  fb.connect()
  if not refload.available():
    pytest.skip('needs the reference rock files')
  root = refload.REF_ROOT
  sid = fb.createVisualShape(halfExtents=(0.5, 0.4, 0.), visualFramePosition=(0.25, 0.25, 0.))
  fb.createMultiBody(baseVisualShapeIndex=sid)
  fb.createMultiBody()                                   # invisible body: skipped
  q = np.array([0.1, -0.2, 0.3, 0.9])
  q /= np.linalg.norm(q)
  for name, pos in (('50_000', (0.1, 0.2, 0.05)), ('95_042', (0.3, 0.25, 0.12))):
    fb.loadURDF(os.path.join(root, 'stackrl/envs/data/generated', name + '.urdf'), pos,
                tuple(q))
  got = observer.PybulletScene(fb)()
  want = fb.scene()
  assert len(got) == len(want) == 3
  for (v, t, rot, pos), (wv, wt, wrot, wpos) in zip(got, want):
    assert np.array_equal(np.asarray(t), wt)
    np.testing.assert_allclose(rot, wrot, atol=1e-15)
    # world-space vertices agree (the box's frame offset is baked into the fake's mesh)
    a = np.asarray(v, dtype='float64').dot(np.asarray(rot).T) + np.asarray(pos)
    b = np.asarray(wv, dtype='float64').dot(wrot.T) + wpos
    np.testing.assert_allclose(a, b, atol=1e-7)
