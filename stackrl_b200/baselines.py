"""Host-side mirror of ``stackrl.baselines`` backed by the sm_100a kernels.

Same names, arguments and return conventions as the reference module
(/root/reference/stackrl/baselines.py) so that callers written against it --
``Baseline(method=...)``, ``stackrl.test.run``'s ``policy(o) -> (a, v)``, the
heat-map sweep -- keep working; the arithmetic runs on the GPU through the C
ABI (stackrl_b200.capi).  Two surfaces:

* drop-in functions on ONE reference-layout observation
  ``(wall_goal [H,W,2], rock [h,w,1])`` given as numpy arrays (host) -- they
  upload, run the kernel and return numpy arrays of the reference's dtype;
* ``PlacementScorer`` on device-resident batches (planar ``walls [E,H,W]``,
  ``goals [E,H,W]``, ``rocks [E,R,h,h]`` torch CUDA tensors), the form the
  batched environment and bench.py use.

Nothing here computes on the CPU; without a CUDA device every call raises.
"""
import numpy as np
import torch

from stackrl_b200 import capi


def _device():
  if not torch.cuda.is_available():
    raise RuntimeError('stackrl_b200 needs a CUDA device (no CPU fallback)')
  return torch.device('cuda', torch.cuda.current_device())


def _split(inputs):
  """Reference-layout observation -> (wall, goal, rock) numpy planes.
  Accepts one observation or the batched TestStackEnv layout."""
  wall_goal = np.asarray(inputs[0])
  rock = np.asarray(inputs[1])
  if wall_goal.shape[-1] != 2 or rock.shape[-1] != 1:
    raise ValueError('expected ([.., H, W, 2], [.., h, w, 1]) observations')
  if rock.shape[-2] != rock.shape[-3]:
    # baselines.py:39 slices both axes with rock.shape[0].
    raise ValueError('the reference only defines square rock maps')
  return wall_goal[..., 0], wall_goal[..., 1], rock[..., 0]


def _upload(x, dtype):
  return torch.from_numpy(np.ascontiguousarray(x)).to(_device()).to(dtype)


# ---- baselines.py:28-43 ------------------------------------------------------ #
def height(inputs, mask=None, **kwargs):
  """Height based heuristic (max-plus drop map), reference ``height``.

  Returns a float64 [H-h+1, W-w+1] array like the reference's np.zeros
  container.  float32 observations use the float32 kernel (bit-exact with
  numpy's float32 add/max), integer observations the float64 one."""
  wall, goal, rock = _split(inputs)
  if wall.ndim != 2:
    raise ValueError('height() takes one observation; use PlacementScorer for batches')
  if wall.dtype == np.float32:
    level = torch.tensor([goal.max()], dtype=torch.float32, device=_device())
    f = capi.maxplus_f32(_upload(wall[None], torch.float32),
                         _upload(rock[None, None], torch.float32), level)
    f = f[0, 0].cpu().numpy().astype('float64')
  else:
    raise NotImplementedError(
      'observation dtype {} is not wired to a kernel yet'.format(wall.dtype))
  if mask is not None:
    f = np.where(mask, f, 0.)
  return f


methods = {
  'height': height,
}
