"""Load the UNMODIFIED reference modules from /root/reference behind stubs.

Test infrastructure (see oracle/__init__.py).  The reference cannot be
imported as a package in this image: ``stackrl/__init__.py`` pulls in
TensorFlow, gym, pybullet and gin, none of which are installed.  The files on
the observation / placement-scoring path, however, only need a handful of
names from those packages, so this loader executes the reference's own source
files (never copied into this repo) with small stand-in modules:

  stackrl/agents/policies.py        -> ns.policies   (needs ``tensorflow.nest``)
  stackrl/baselines.py              -> ns.baselines  (needs ``gin``, ``gym``)
  stackrl/envs/stack/observer.py    -> ns.observer   (numpy only)
  stackrl/envs/stack/simulator.py   -> ns.simulator  (needs ``pybullet``)
  stackrl/envs/stack/rewarder.py    -> ns.rewarder   (needs ``gym.utils.seeding``)
  stackrl/envs/stack/env.py         -> ns.env        (needs old-API ``gym``)
  stackrl/envs/data/__init__.py     -> ns.data

``/root/reference`` only exists in the build container, so everything that
uses this module (tests/golden/make_golden.py, tests marked ``needs_reference``)
is skipped on the GPU box; the committed fixtures travel instead.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

def _root():
  """The reference tree: $STACKRL_REFERENCE, /root/reference (build container), or the
  byte-for-byte staging of the hot-path files under oracle/_ref (oracle/make_ref.py;
  travels with the repo snapshot to the GPU box, never committed)."""
  if os.environ.get('STACKRL_REFERENCE'):
    return os.environ['STACKRL_REFERENCE']
  if os.path.isfile('/root/reference/stackrl/baselines.py'):
    return '/root/reference'
  return os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


REF_ROOT = _root()


def available():
  return os.path.isfile(os.path.join(REF_ROOT, 'stackrl', 'baselines.py'))


def _module(name, **attrs):
  m = types.ModuleType(name)
  m.__spec__ = importlib.machinery.ModuleSpec(name, None)
  for k, v in attrs.items():
    setattr(m, k, v)
  return m


# --------------------------------------------------------------------------- #
# gin stub: decorators that return the decorated object untouched.
# --------------------------------------------------------------------------- #
def _gin_stub():
  def configurable(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
      return args[0]
    return lambda obj: obj
  return _module(
    'gin',
    configurable=configurable,
    external_configurable=lambda obj, *a, **k: obj,
  )


# --------------------------------------------------------------------------- #
# tensorflow stub: the three names agents/policies.py touches at import time
# and in PyGreedy's batched path (policies.py:63-64).
# --------------------------------------------------------------------------- #
def _tf_stub():
  def flatten(x):
    if isinstance(x, (tuple, list)):
      out = []
      for i in x:
        out.extend(flatten(i))
      return out
    return [x]

  def pack_sequence_as(structure, flat):
    flat = list(flat)
    def build(s):
      if isinstance(s, (tuple, list)):
        return type(s)(build(i) for i in s)
      return flat.pop(0)
    return build(structure)

  nest = _module('tensorflow.nest', flatten=flatten,
                 pack_sequence_as=pack_sequence_as)
  return _module('tensorflow', Module=object, nest=nest)


# --------------------------------------------------------------------------- #
# gym stub: the pre-0.21 API surface env.py / rewarder.py use.
# --------------------------------------------------------------------------- #
def _gym_stub():
  class Space(object):
    def seed(self, seed=None):
      self._rng = np.random.RandomState(seed)
    @property
    def rng(self):
      if not hasattr(self, '_rng'):
        self._rng = np.random.RandomState()
      return self._rng

  class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
      self.dtype = np.dtype(dtype)
      self.shape = tuple(shape)
      self.low = np.full(self.shape, low, dtype=self.dtype)
      self.high = np.full(self.shape, high, dtype=self.dtype)
    def sample(self):
      if self.dtype.kind == 'f':
        return self.rng.uniform(self.low, self.high).astype(self.dtype)
      return self.rng.randint(
        self.low.astype('int64'), self.high.astype('int64') + 1
      ).astype(self.dtype)
    def contains(self, x):
      x = np.asarray(x)
      return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

  class Discrete(Space):
    def __init__(self, n):
      self.n = n
      self.shape = ()
      self.dtype = np.dtype('int64')
    def sample(self):
      return int(self.rng.randint(self.n))
    def contains(self, x):
      try:
        xi = int(x)
      except (TypeError, ValueError):
        return False
      return xi == x and 0 <= xi < self.n

  class MultiDiscrete(Space):
    def __init__(self, nvec):
      self.nvec = np.asarray(nvec, dtype='int64')
      self.shape = self.nvec.shape
      self.dtype = np.dtype('int64')
    def sample(self):
      return (self.rng.random_sample(self.nvec.shape) * self.nvec).astype('int64')
    def contains(self, x):
      x = np.asarray(x)
      return x.shape == self.shape and np.all(x >= 0) and np.all(x < self.nvec)

  class Tuple(Space):
    def __init__(self, spaces):
      self.spaces = tuple(spaces)
    def __getitem__(self, i):
      return self.spaces[i]
    def __len__(self):
      return len(self.spaces)
    def __iter__(self):
      return iter(self.spaces)
    def sample(self):
      return tuple(s.sample() for s in self.spaces)
    def contains(self, x):
      return len(x) == len(self.spaces) and all(
        s.contains(i) for s, i in zip(self.spaces, x))

  class Env(object):
    metadata = {}
    def render(self, mode='human'):
      raise NotImplementedError
    def close(self):
      pass

  class Error(Exception):
    pass

  class _Spec(object):
    def __init__(self, id, entry_point, kwargs):
      self.id, self.entry_point, self._kwargs = id, entry_point, dict(kwargs or {})

  class _Registry(object):
    def __init__(self):
      self.env_specs = {}

  registry = _Registry()

  def register(id, entry_point=None, kwargs=None, **_):
    if id in registry.env_specs:
      raise Error('Cannot re-register id: {}'.format(id))
    registry.env_specs[id] = _Spec(id, entry_point, kwargs)

  def make(id, **kwargs):
    spec = registry.env_specs[id]
    kw = dict(spec._kwargs)
    kw.update(kwargs)
    return spec.entry_point(**kw)

  def np_random(seed=None):
    # Old gym hashes the seed before feeding RandomState; the stream itself is
    # host-side and not part of any parity claim, so a plain RandomState does.
    if seed is None:
      seed = int(np.random.SeedSequence().generate_state(1)[0])
    return np.random.RandomState(int(seed) % 2**32), seed

  def create_seed(seed=None, max_bytes=8):
    if seed is None:
      return int(np.random.SeedSequence().generate_state(1)[0])
    return int(seed) % 2**(8 * max_bytes)

  spaces = _module('gym.spaces', Box=Box, Discrete=Discrete,
                   MultiDiscrete=MultiDiscrete, Tuple=Tuple, Space=Space)
  seeding = _module('gym.utils.seeding', np_random=np_random,
                    create_seed=create_seed)
  utils = _module('gym.utils', seeding=seeding)
  envs = _module('gym.envs', registry=registry)
  error = _module('gym.error', Error=Error)
  gym = _module('gym', Env=Env, Space=Space, spaces=spaces, utils=utils,
                envs=envs, error=error, register=register, make=make)
  return {
    'gym': gym, 'gym.spaces': spaces, 'gym.utils': utils,
    'gym.utils.seeding': seeding, 'gym.envs': envs, 'gym.error': error,
  }


def _exec(name, relpath, registry):
  path = os.path.join(REF_ROOT, relpath)
  spec = importlib.util.spec_from_file_location(name, path)
  mod = importlib.util.module_from_spec(spec)
  registry[name] = mod
  sys.modules[name] = mod
  spec.loader.exec_module(mod)
  return mod


_CACHE = {}


def load(pybullet=None):
  """Returns a namespace holding the reference modules on the hot path.

  Args:
    pybullet: module-like object installed as ``pybullet`` while
      ``simulator.py`` is executed (``oracle.fake_pybullet`` for end-to-end
      runs).  ``None`` installs an empty placeholder, enough for the class
      definitions (``Rewarder``'s isinstance gates) to exist.
  """
  key = id(pybullet)
  if key in _CACHE:
    return _CACHE[key]
  if not available():
    raise RuntimeError(
      'reference tree not found at {} (it only exists in the build '
      'container; use the committed tests/golden fixtures)'.format(REF_ROOT))

  stubs = {'gin': _gin_stub(), 'tensorflow': _tf_stub()}
  stubs.update(_gym_stub())
  stubs['pybullet'] = pybullet if pybullet is not None else _module(
    'pybullet', GUI=1, DIRECT=2, GEOM_BOX=3, COV_ENABLE_GUI=1)
  for pkg in ('stackrl', 'stackrl.agents', 'stackrl.envs', 'stackrl.envs.stack'):
    stubs[pkg] = _module(pkg)
    stubs[pkg].__path__ = []

  saved = {k: sys.modules.get(k) for k in stubs}
  loaded = {}
  sys.modules.update(stubs)
  try:
    policies = _exec('stackrl.agents.policies', 'stackrl/agents/policies.py', loaded)
    stubs['stackrl.agents'].PyGreedy = policies.PyGreedy
    stubs['stackrl'].agents = stubs['stackrl.agents']
    baselines = _exec('stackrl.baselines', 'stackrl/baselines.py', loaded)
    data = _exec('stackrl.envs.data', 'stackrl/envs/data/__init__.py', loaded)
    stubs['stackrl.envs'].data = data
    observer = _exec('stackrl.envs.stack.observer',
                     'stackrl/envs/stack/observer.py', loaded)
    simulator = _exec('stackrl.envs.stack.simulator',
                      'stackrl/envs/stack/simulator.py', loaded)
    rewarder = _exec('stackrl.envs.stack.rewarder',
                     'stackrl/envs/stack/rewarder.py', loaded)
    env = _exec('stackrl.envs.stack.env', 'stackrl/envs/stack/env.py', loaded)
  finally:
    # Leave no stub behind: other libraries probe sys.modules for these names.
    for k in list(loaded) + list(stubs):
      if saved.get(k) is not None:
        sys.modules[k] = saved[k]
      else:
        sys.modules.pop(k, None)

  ns = types.SimpleNamespace(
    policies=policies, baselines=baselines, data=data, observer=observer,
    simulator=simulator, rewarder=rewarder, env=env, gym=stubs['gym'],
    root=REF_ROOT,
  )
  _CACHE[key] = ns
  return ns
