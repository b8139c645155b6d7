"""Golden episodes of the UNMODIFIED reference env on the fake pybullet backend.

Called from make_golden.py (build container only).  Runs the reference's own
StackEnv / TestStackEnv (env.py), Observer, Rewarder and Baseline('height')
with oracle.fake_pybullet as the ``pybullet`` module ("reference code, fake
physics": bodies stay where they are placed) and records, per step, the depth
images the fake renderer produced, the Observer's maps, the poses, the rewards,
the packed observations and the actions.  The meshes used are stored too, so the
fixture is self-contained on the GPU box.
"""
import os

import numpy as np

from oracle import fake_pybullet, refload

HERE = os.path.dirname(os.path.abspath(__file__))


def _episode(ns, fb, env_cls, urdfs, dtype, steps, policy_kwargs, out, prefix, light=False,
             **env_kwargs):
  """``light``: a long episode (C1: 30 steps) keeps the packed observations, actions, poses
  and rewards of every step but not the float maps and depth images."""
  depth_log = []
  real = fb.getCameraImage

  def recording(*a, **k):
    res = real(*a, **k)
    depth_log.append(res[3].copy())
    return res
  fb.getCameraImage = recording

  env = env_cls(urdfs=urdfs, dtype=dtype, rewarder='all', seed=7, episode_length=steps,
                **env_kwargs)
  pol = ns.baselines.Baseline(method='height', value=True, **policy_kwargs)
  obs = env.reset()
  order = [env._sim._last_urdf] if hasattr(env._sim, '_last_urdf') else []
  out[prefix + '/goal'] = env._rew.goal.copy()
  out[prefix + '/goal_lims'] = np.array(env._rew._goal_lims)
  out[prefix + '/n_steps'] = np.int64(steps)
  k = 0
  done = False
  while not done:
    m, n = env._obs.state
    out['{}/s{}/obs0'.format(prefix, k)] = obs[0]
    out['{}/s{}/obs1'.format(prefix, k)] = obs[1]
    if not light:
      out['{}/s{}/overhead_map'.format(prefix, k)] = np.array(m)
      out['{}/s{}/object_map'.format(prefix, k)] = np.array(n)
      out['{}/s{}/depths'.format(prefix, k)] = np.concatenate(
        [d.ravel() for d in depth_log]) if depth_log else np.zeros(0, 'float32')
      out['{}/s{}/depth_shapes'.format(prefix, k)] = np.array([d.shape for d in depth_log])
    del depth_log[:]
    a, v = pol(obs)
    if isinstance(a, tuple):
      action = (int(a[0]), int(a[1]))
      out['{}/s{}/action'.format(prefix, k)] = np.array(action, dtype='int64')
      pose = env._obs.pose([action[1] // env._action_width, action[1] % env._action_width],
                           index=action[0])
      out['{}/s{}/pose_orientation'.format(prefix, k)] = np.array(pose['orientation'])
    else:
      action = int(a)
      out['{}/s{}/action'.format(prefix, k)] = np.int64(action)
      pose = env._obs.pose([action // env._action_width, action % env._action_width])
    out['{}/s{}/pose_position'.format(prefix, k)] = np.array(pose['position'], dtype='float64')
    obs, reward, done, info = env.step(action)
    out['{}/s{}/rewards'.format(prefix, k)] = np.array(
      [info[name] for name in ('IoU', 'OR', 'DIoU', 'DOR')], dtype='float64')
    k += 1
  # terminal observation
  m, n = env._obs.state
  out['{}/s{}/overhead_map'.format(prefix, k)] = np.array(m)
  out['{}/s{}/obs0'.format(prefix, k)] = obs[0]
  if not light:
    out['{}/s{}/depths'.format(prefix, k)] = np.concatenate([d.ravel() for d in depth_log])
    out['{}/s{}/depth_shapes'.format(prefix, k)] = np.array([d.shape for d in depth_log])
  out[prefix + '/n_recorded'] = np.int64(k)
  fb.getCameraImage = real
  env.close()


def make_goal_draws():
  """a13 fixture: what the reference's two RandomState streams (StackEnv._random for
  the episode list, env.py:268-272; Rewarder._random for the goal rectangle,
  rewarder.py:211-259, seeded through env.seed -> Rewarder.seed, env.py:340-346)
  draw for a given seed: rock order and goal limits of the first two episodes, for
  24 seeds and the three goal_size_ratio modes (scalar, tuple, None)."""
  fb = fake_pybullet.FakeBullet()
  ns = refload.load(pybullet=fb)
  names = ['0_0', '0_3', '50_000', '55_017', '60_250', '75_003', '80_499', '95_042']
  urdfs = [os.path.join(ns.root, 'stackrl/envs/data/generated', n + '.urdf') for n in names]
  index = {os.path.basename(u): k for k, u in enumerate(urdfs)}
  loaded = []
  real_load = fb.loadURDF

  def logging_load(fileName, *a, **k):
    loaded.append(index[os.path.basename(fileName)])
    return real_load(fileName, *a, **k)
  fb.loadURDF = logging_load
  out = {'n_meshes': np.int64(len(urdfs))}
  modes = [('scalar', .25, 6), ('tuple', (.5, .25), 12), ('none', None, 5)]
  seeds = list(range(20)) + [123, 4242, 2 ** 31 + 5, 2 ** 32 - 2]
  out['seeds'] = np.array(seeds, dtype='int64')
  for tag, ratio, length in modes:
    lims = np.zeros((len(seeds), 2, 2, 2), dtype='int64')
    orders = np.zeros((len(seeds), 2, length), dtype='int64')
    env = ns.env.StackEnv(urdfs=urdfs, seed=0, goal_size_ratio=ratio, episode_length=length)
    for k, s in enumerate(seeds):
      env.seed(s)
      for ep in range(2):
        del loaded[:]
        env.reset()
        lims[k, ep] = np.array(env._rew._goal_lims)
        # pop order: the rock loaded by reset, then the list from its end (env.py:245)
        rest = [index[os.path.basename(u)] for u in env._episode_list][::-1]
        orders[k, ep] = [loaded[0]] + rest
    env.close()
    out[tag + '/goal_lims'] = lims
    out[tag + '/orders'] = orders
    out[tag + '/episode_length'] = np.int64(length)
  path = os.path.join(HERE, 'goal_draws.npz')
  np.savez_compressed(path, **out)
  print('goal_draws.npz: {} arrays, {:.0f} KB'.format(len(out), os.path.getsize(path) / 1024))


def main(ns=None):
  fb = fake_pybullet.FakeBullet()
  ns = refload.load(pybullet=fb)
  # Log the order in which the simulator loads URDFs (the env's RNG decides it).
  loaded = []
  real_load = fb.loadURDF

  def logging_load(fileName, *a, **k):
    loaded.append(os.path.basename(fileName))
    return real_load(fileName, *a, **k)
  fb.loadURDF = logging_load

  names = ['0_0', '0_3', '50_000', '55_017', '60_250', '75_003', '80_499', '95_042']
  urdfs = [os.path.join(ns.root, 'stackrl/envs/data/generated', n + '.urdf') for n in names]
  out = {}
  for n, u in zip(names, urdfs):
    mesh, com = fake_pybullet.parse_urdf(u)
    v, t = fake_pybullet.load_obj(mesh)
    out['mesh/{}/verts'.format(n)] = v
    out['mesh/{}/tris'.format(n)] = t
    out['mesh/{}/com'.format(n)] = com
  out['mesh_names'] = np.array(names)

  episodes = [
    ('stack_f32', ns.env.StackEnv, 'float32', 6, {}, {}),
    ('stack_u8', ns.env.StackEnv, 'uint8', 6, {}, {}),
    ('test_f32_rot8', ns.env.TestStackEnv, 'float32', 5,
     dict(batched=True, batchwise=True), dict(orientation_freedom=3)),
  ]
  for prefix, cls, dtype, steps, pk, ek in episodes:
    del loaded[:]
    _episode(ns, fb, cls, urdfs, dtype, steps, pk, out, prefix, **ek)
    out[prefix + '/urdf_order'] = np.array([n[:-5] for n in loaded])
  # C1 (BASELINE config 1): the registered Stack-v0 defaults -- uint8 observations, 128x128
  # wall, 32x32 rock, 30 rocks per episode (env.py:20) -- with Baseline('height'); 30 > 8
  # meshes, so the episode list is drawn with replacement (env.py:268-272).
  del loaded[:]
  _episode(ns, fb, ns.env.StackEnv, urdfs, 'uint8', 30, {}, out, 'c1_stack_v0_30', light=True)
  out['c1_stack_v0_30/urdf_order'] = np.array([n[:-5] for n in loaded])
  path = os.path.join(HERE, 'observe.npz')
  np.savez_compressed(path, **out)
  print('observe.npz: {} arrays, {:.0f} KB'.format(len(out), os.path.getsize(path) / 1024))


if __name__ == '__main__':
  main()
  make_goal_draws()
