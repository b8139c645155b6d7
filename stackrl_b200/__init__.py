"""stackrl_b200: B200-native observation + placement-scoring path of stackrl.

Host code is Python mirroring the reference's interfaces
(``stackrl.baselines``, ``stackrl.envs.stack.observer``); all arithmetic runs in
hand-written sm_100a CUDA kernels reached through the C-ABI library
``libstackrl_b200.so`` (include/stackrl_b200.h) via ctypes.  There is no CPU
fallback: importing any compute module without the built library raises.

Submodules are imported lazily so that the numpy-only helpers
(``stackrl_b200.synth``) stay usable on machines without the CUDA build.
"""
import importlib

__version__ = '0.1.0'

_LAZY = ('capi', 'baselines', 'observer', 'envs', 'episodes', 'meshes', 'nets',
         'sharding', 'synth', 'camera')


def __getattr__(name):
  if name in _LAZY:
    return importlib.import_module('stackrl_b200.' + name)
  raise AttributeError(name)
