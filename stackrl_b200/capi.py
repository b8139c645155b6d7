"""ctypes binding of libstackrl_b200.so (include/stackrl_b200.h).

PyTorch is used only as the device-buffer interchange: every wrapper takes
contiguous CUDA tensors, passes ``data_ptr()`` and the current stream to the
C ABI, and returns the output tensor.  There is no CPU fallback: if the library
is missing this module raises ImportError at import time, and every entry point
raises on a non-CUDA tensor.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libstackrl_b200.so')


class SrlError(RuntimeError):
  """A C-ABI call returned a negative code."""
  def __init__(self, code, message):
    super(SrlError, self).__init__('srl error {}: {}'.format(code, message))
    self.code = code


if not os.path.exists(LIB_PATH):
  raise ImportError(
    '{} is missing: build it with `python -m stackrl_b200.build` (nvcc, sm_100a). '
    'stackrl_b200 has no CPU fallback.'.format(LIB_PATH))

lib = ctypes.CDLL(LIB_PATH)

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int

_SIGNATURES = {
  'srl_version': (_I, []),
  'srl_last_error': (_c.c_char_p, []),
  'srl_device_sm_count': (_I, [_c.POINTER(_I)]),
  'srl_maxplus_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _c.c_float, _P]),
  'srl_maxplus_f32_q': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _c.c_float, _I, _P]),
  'srl_maxplus_u8': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
  'srl_drop_height_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _c.c_float, _P]),
  'srl_goal_overlap_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
  'srl_goal_overlap_u8': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
  'srl_select_f32': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _c.c_double, _P]),
  'srl_select_f64': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _c.c_double, _P]),
  'srl_mask_select_f32': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _c.c_double, _P]),
  'srl_mask_select_f64': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _c.c_double, _P]),
  'srl_mask_select_f64_u8': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _c.c_double, _P]),
  'srl_score_f32': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _c.c_double, _P]),
  'srl_difference_weights': (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
  'srl_difference_f32': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
  'srl_difference_weights_u8': (_I, [_P, _P, _I, _I, _I, _I, _P]),
  'srl_difference_u8': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
  'srl_corrcoef_localized_f32': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
  'srl_corrcoef_localized_u8': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
  'srl_correlate_f32': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
  'srl_siam_correlation_f32': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
  'srl_siam_correlation_grad_f32': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
  'srl_raster': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _c.c_double, _P]),
  'srl_raster_ex': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _c.c_double, _I, _P]),
  'srl_raster_incremental': (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _c.c_double,
                                   _I, _P]),
  'srl_raster_incremental_rows': (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I,
                                        _c.c_double, _I, _P]),
  'srl_reward_sums_f32': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
  'srl_pack_obs': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _c.c_float, _I, _P]),
  'srl_place_poses_f32': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _c.c_double,
                                _c.c_double, _c.c_double, _c.c_double, _c.c_double,
                                _c.c_float, _P]),
  'srl_contact_precheck_f32': (_I, [_P] * 7 + [_I] * 6 + [_c.c_float, _c.c_float, _P]),
  'srl_env_reset': (_I, [_P, _P, _I, _P]),
  'srl_env_advance': (_I, [_P, _P, _P, _P]),
  'srl_env_set_poses': (_I, [_P, _P, _I, _P]),
  'srl_env_draw': (_I, [_P, _P, _P, _P] + [_I] * 10 + [_c.c_uint64, _c.c_uint64, _P]),
  'srl_fill_goals_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
  'srl_goal_level_f32': (_I, [_P, _P, _I, _I, _P]),
  'srl_goal_level_u8': (_I, [_P, _P, _I, _I, _P]),
  'srl_rewards_f32': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _c.c_double, _c.c_double,
                            _c.c_double, _c.c_double, _c.c_double, _c.c_double, _P]),
  'srl_pack_rewards_f32': (_I, [_P] * 10 + [_I, _I, _I, _I, _I, _c.c_float, _I, _I] +
                           [_c.c_double] * 6 + [_P]),
  'srl_pack_rewards_rows_f32': (_I, [_P] * 12 + [_I, _I, _I, _I, _I, _c.c_float, _I, _I] +
                                [_c.c_double] * 6 + [_P]),
  'srl_quantise_planes_u8': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _c.c_float, _P]),
  'srl_gather_rows_f32': (_I, [_P, _P, _P, _I, _I, _I, _P]),
  'srl_microbench_addmax': (_I, [_I, _I, _c.POINTER(_c.c_double)]),
  'srl_microbench_fma': (_I, [_I, _I, _c.POINTER(_c.c_double)]),
}

for _name, (_res, _args) in _SIGNATURES.items():
  try:
    _fn = getattr(lib, _name)
  except AttributeError:
    continue          # tests/test_capi_symbols.py checks the header <-> export match
  _fn.restype = _res
  _fn.argtypes = _args


def _check(code):
  if code != 0:
    raise SrlError(code, lib.srl_last_error().decode('utf-8', 'replace'))


def _dev(t, dtype, name):
  if not isinstance(t, torch.Tensor) or not t.is_cuda:
    raise TypeError('{} must be a CUDA tensor (stackrl_b200 has no CPU path)'.format(name))
  if t.dtype != dtype:
    raise TypeError('{} must be {}, got {}'.format(name, dtype, t.dtype))
  if not t.is_contiguous():
    raise ValueError('{} must be contiguous'.format(name))
  return _P(t.data_ptr())


def _out(t, dtype, shape, like, name='out'):
  """A caller-provided output buffer: right dtype, element count and device (the
  kernels write ``prod(shape)`` values into it, partly with bulk TMA stores)."""
  n = 1
  for d in shape:
    n *= int(d)
  if t.numel() != n:
    raise ValueError('{} must hold {} values ({}), got {}'.format(
      name, n, tuple(shape), tuple(t.shape)))
  if t.device != like.device:
    raise ValueError('{} is on {}, the inputs on {}'.format(name, t.device, like.device))
  return _dev(t, dtype, name)


def _same_device(*tensors):
  dev = None
  for t in tensors:
    if t is None:
      continue
    if dev is None:
      dev = t.device
    elif t.device != dev:
      raise ValueError('all tensors of one call must live on one device ({} vs {})'.format(
        dev, t.device))


def _opt(t, dtype, name):
  return _P(None) if t is None else _dev(t, dtype, name)


def _stream():
  return _P(torch.cuda.current_stream().cuda_stream)


def version():
  return lib.srl_version()


def sm_count():
  n = _I(0)
  _check(lib.srl_device_sm_count(ctypes.byref(n)))
  return n.value


def maxplus_f32(walls, rocks, level=None, threshold=0., out=None, quantum_log2=None):
  """walls [E,H,W], rocks [E,R,h,h], level [E] or None -> [E,R,H-h+1,W-h+1].
  ``quantum_log2``: hint that the values are non-negative multiples of
  2**quantum_log2 (heightmaps from the rasteriser: -14); environments for which
  it holds are swept in exact 16-bit fixed point, same bits either way."""
  E, H, W = walls.shape
  E2, R, h, h2 = rocks.shape
  if E2 != E or h != h2:
    raise ValueError('rocks must be [E, R, h, h] matching walls [E, H, W]')
  _same_device(walls, rocks, level)
  args = (_dev(walls, torch.float32, 'walls'), _dev(rocks, torch.float32, 'rocks'),
          _opt(level, torch.float32, 'level'))
  shape = (E, R, H - h + 1, W - h + 1)
  if h > H or h > W:
    # no candidate position: the library's own argument check reports it
    with torch.cuda.device(walls.device):
      _check(lib.srl_maxplus_f32(*args, _P(None), E, R, H, W, h, float(threshold), _stream()))
  if out is None:
    out = torch.empty(shape, dtype=torch.float32, device=walls.device)
  _out(out, torch.float32, shape, walls)
  with torch.cuda.device(walls.device):
    if quantum_log2 is None:
      _check(lib.srl_maxplus_f32(*args, _dev(out, torch.float32, 'out'),
                                 E, R, H, W, h, float(threshold), _stream()))
    else:
      _check(lib.srl_maxplus_f32_q(*args, _dev(out, torch.float32, 'out'),
                                   E, R, H, W, h, float(threshold), int(quantum_log2),
                                   _stream()))
  return out


def maxplus_u8(walls, rocks, level, out=None):
  """uint8 walls [E,H,W], rocks [E,R,h,h], level [E] -> float64 [E,R,Ph,Pw]."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  _same_device(walls, rocks, level)
  args = (_dev(walls, torch.uint8, 'walls'), _dev(rocks, torch.uint8, 'rocks'),
          _dev(level, torch.uint8, 'level'))
  shape = (E, R, H - h + 1, W - h + 1)
  if out is None:
    out = torch.empty(shape, dtype=torch.float64, device=walls.device)
  _out(out, torch.float64, shape, walls)
  with torch.cuda.device(walls.device):
    _check(lib.srl_maxplus_u8(*args, _dev(out, torch.float64, 'out'),
                              E, R, H, W, h, _stream()))
  return out


def _batch_dims(walls, rocks):
  E, H, W = walls.shape
  E2, R, h, h2 = rocks.shape
  if E2 != E or h != h2:
    raise ValueError('rocks must be [E, R, h, h] matching walls [E, H, W]')
  return E, R, H, W, h


def drop_height_f32(walls, rocks, picks, threshold=1e-4):
  """picks [E,3] int32 (rotation, row, column) -> [E] float32 drop heights."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  args = (_dev(walls, torch.float32, 'walls'), _dev(rocks, torch.float32, 'rocks'),
          _dev(picks, torch.int32, 'picks'))
  out = torch.empty((E,), dtype=torch.float32, device=walls.device)
  with torch.cuda.device(walls.device):
    _check(lib.srl_drop_height_f32(*args, _dev(out, torch.float32, 'out'),
                                   E, R, H, W, h, float(threshold), _stream()))
  return out


def goal_overlap(walls, goals, rocks):
  """Integer overlap counts [E,R,Ph,Pw] (int32) for float32 or uint8 maps."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  dt = walls.dtype
  if dt == torch.float32:
    fn = lib.srl_goal_overlap_f32
  elif dt == torch.uint8:
    fn = lib.srl_goal_overlap_u8
  else:
    raise TypeError('goal_overlap takes float32 or uint8 maps, got {}'.format(dt))
  args = (_dev(walls, dt, 'walls'), _dev(goals, dt, 'goals'), _dev(rocks, dt, 'rocks'))
  out = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.int32,
                    device=walls.device)
  with torch.cuda.device(walls.device):
    _check(fn(*args, _dev(out, torch.int32, 'counts'), E, R, H, W, h, _stream()))
  return out


def select(values, counts=None, minorder=1, overlap_threshold=0.75,
           want_shown=True, want_best=True):
  """values [E,R,Ph,Pw] (float32/float64), counts like values (int32) or None
  -> (actions [E,R] int64, shown [E,R,Ph,Pw] float64 | None, best [E,2] int64 | None)."""
  E, R, Ph, Pw = values.shape
  if values.dtype == torch.float32:
    fn = lib.srl_select_f32
  elif values.dtype == torch.float64:
    fn = lib.srl_select_f64
  else:
    raise TypeError('select takes float32 or float64 values')
  v = _dev(values, values.dtype, 'values')
  c = _opt(counts, torch.int32, 'counts')
  if counts is not None and counts.shape != values.shape:
    raise ValueError('counts must have the shape of values')
  dev = values.device
  actions = torch.empty((E, R), dtype=torch.int64, device=dev)
  shown = torch.empty((E, R, Ph, Pw), dtype=torch.float64, device=dev) if want_shown else None
  best = torch.empty((E, 2), dtype=torch.int64, device=dev) if want_best else None
  with torch.cuda.device(dev):
    _check(fn(v, c, _dev(actions, torch.int64, 'actions'),
              _opt(shown, torch.float64, 'shown'), _opt(best, torch.int64, 'best'),
              E, R, Ph, Pw, int(minorder), float(overlap_threshold), _stream()))
  return actions, shown, best


def difference_weights(rocks, level=None, weights_exponent=2):
  """rocks [E,R,h,h] float32 or uint8 -> normalised radial weights [E,R,h,h] float64."""
  E, R, h, h2 = rocks.shape
  out = torch.empty((E, R, h, h), dtype=torch.float64, device=rocks.device)
  if rocks.dtype == torch.uint8:
    with torch.cuda.device(rocks.device):
      _check(lib.srl_difference_weights_u8(_dev(rocks, torch.uint8, 'rocks'),
                                           _dev(out, torch.float64, 'weights'),
                                           E, R, h, int(weights_exponent), _stream()))
    return out
  args = (_dev(rocks, torch.float32, 'rocks'), _opt(level, torch.float32, 'level'))
  with torch.cuda.device(rocks.device):
    _check(lib.srl_difference_weights(*args, _dev(out, torch.float64, 'weights'),
                                      E, R, h, int(weights_exponent), _stream()))
  return out


def difference_f32(walls, rocks, level, weights, difference_exponent=2, want_top=False):
  """-> (f [E,R,Ph,Pw] float64, h0 [E,R,Ph,Pw] float32 | None)."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  args = (_dev(walls, torch.float32, 'walls'), _dev(rocks, torch.float32, 'rocks'),
          _opt(level, torch.float32, 'level'), _dev(weights, torch.float64, 'weights'))
  shape = (E, R, H - h + 1, W - h + 1)
  out = torch.empty(shape, dtype=torch.float64, device=walls.device)
  # always hand over a `top` buffer: the library then takes h0 from its max-plus
  # kernel instead of a scalar pass per position
  top = torch.empty(shape, dtype=torch.float32, device=walls.device)
  with torch.cuda.device(walls.device):
    _check(lib.srl_difference_f32(*args, _dev(out, torch.float64, 'out'),
                                  _opt(top, torch.float32, 'top'), E, R, H, W, h,
                                  int(difference_exponent), _stream()))
  return out, (top if want_top else None)


def difference_u8(walls, rocks, level, weights, difference_exponent=2, want_top=False):
  """uint8 observations -> (f [E,R,Ph,Pw] float64, h0 [E,R,Ph,Pw] float64 | None)."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  args = (_dev(walls, torch.uint8, 'walls'), _dev(rocks, torch.uint8, 'rocks'),
          _opt(level, torch.uint8, 'level'), _dev(weights, torch.float64, 'weights'))
  shape = (E, R, H - h + 1, W - h + 1)
  out = torch.empty(shape, dtype=torch.float64, device=walls.device)
  top = torch.empty(shape, dtype=torch.float64, device=walls.device)
  with torch.cuda.device(walls.device):
    _check(lib.srl_difference_u8(*args, _dev(out, torch.float64, 'out'),
                                 _opt(top, torch.float64, 'top'), E, R, H, W, h,
                                 int(difference_exponent), _stream()))
  return out, (top if want_top else None)


def corrcoef_localized(walls, rocks, level=None):
  """corrcoef(localized=True) for float32 or uint8 planes -> [E,R,Ph,Pw] float64."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  u8 = walls.dtype == torch.uint8
  dt = torch.uint8 if u8 else torch.float32
  cdt = torch.float64 if u8 else torch.float32
  args = (_dev(walls, dt, 'walls'), _dev(rocks, dt, 'rocks'), _opt(level, dt, 'level'))
  work = torch.empty((E * R * (h * h + 2),), dtype=cdt, device=walls.device)
  out = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float64, device=walls.device)
  fn = lib.srl_corrcoef_localized_u8 if u8 else lib.srl_corrcoef_localized_f32
  with torch.cuda.device(walls.device):
    _check(fn(*args, _dev(work, cdt, 'work'), _dev(out, torch.float64, 'out'),
              E, R, H, W, h, _stream()))
  return out


# numpy mirrors of srl_raster_instance / srl_raster_job (include/stackrl_b200.h)
import numpy as _np
INSTANCE_DTYPE = _np.dtype([('rot', '<f8', (9,)), ('pos', '<f8', (3,)),
                            ('vert_begin', '<i4'), ('vert_count', '<i4'),
                            ('tri_begin', '<i4'), ('tri_count', '<i4')], align=True)
JOB_DTYPE = _np.dtype([('view', '<f8', (16,)), ('proj', '<f8', (16,)),
                       ('inst_begin', '<i4'), ('inst_count', '<i4'),
                       ('zrange', '<f8')], align=True)
assert INSTANCE_DTYPE.itemsize == 112 and JOB_DTYPE.itemsize == 272
RASTER_DEPTH, RASTER_WALL, RASTER_ROCK = 0, 1, 2


def raster(verts, tris, instances, jobs, rows, cols, mode, far_plane=1000., out=None,
           inst_counts=None, max_cached_verts=0, depth_state=None, only_last=False,
           rows_out=None):
  """verts [NV,3] f32 / tris [NT,3] i32 CUDA tensors (the mesh bank),
  instances / jobs numpy structured arrays (INSTANCE_DTYPE / JOB_DTYPE) or
  CUDA uint8 tensors holding them -> [njobs, rows, cols] float32.
  ``inst_counts`` [njobs] int32 (device) overrides the jobs' instance counts;
  ``max_cached_verts`` sizes the shared-memory vertex cache (srl_raster_ex).
  ``depth_state`` [njobs, rows, cols] float32: the GL depth image kept between calls
  (srl_raster_incremental); with ``only_last`` only the last instance of every job
  is drawn onto it -- same bits as re-drawing the whole scene; ``only_last=2``: ``out``
  still holds the image of ``depth_state`` and is updated in place.  ``rows_out``
  [njobs, 2] int32 (with ``depth_state``): receives the image rows each job may have
  changed (srl_raster_incremental_rows)."""
  dev = verts.device
  def as_bytes(a, dtype):
    if isinstance(a, torch.Tensor):
      return a
    a = _np.ascontiguousarray(a, dtype=dtype)
    return torch.from_numpy(a.view(_np.uint8).reshape(-1)).to(dev, non_blocking=True)
  njobs = len(jobs) if not isinstance(jobs, torch.Tensor) else jobs.numel() // JOB_DTYPE.itemsize
  inst_t = as_bytes(instances, INSTANCE_DTYPE)
  jobs_t = as_bytes(jobs, JOB_DTYPE)
  if inst_t.numel() == 0:                      # a scene may be empty (ground only)
    inst_t = torch.zeros(INSTANCE_DTYPE.itemsize, dtype=torch.uint8, device=dev)
  if verts.numel() == 0:
    verts = torch.zeros((1, 3), dtype=torch.float32, device=dev)
  if tris.numel() == 0:
    tris = torch.zeros((1, 3), dtype=torch.int32, device=dev)
  if out is None:
    out = torch.empty((njobs, rows, cols), dtype=torch.float32, device=dev)
  _out(out, torch.float32, (njobs, rows, cols), verts)
  _same_device(verts, tris, inst_t, jobs_t)
  args = (_dev(verts, torch.float32, 'verts'), _dev(tris, torch.int32, 'tris'),
          _dev(inst_t, torch.uint8, 'instances'), _dev(jobs_t, torch.uint8, 'jobs'),
          _dev(out, torch.float32, 'out'))
  with torch.cuda.device(dev):
    if depth_state is not None and rows_out is not None:
      _check(lib.srl_raster_incremental_rows(
        *args[:4], _opt(inst_counts, torch.int32, 'inst_counts'),
        _out(depth_state, torch.float32, (njobs, rows, cols), verts, 'depth_state'),
        int(only_last), args[4], _out(rows_out, torch.int32, (njobs, 2), verts, 'rows_out'),
        int(njobs), int(rows), int(cols), int(mode), float(far_plane), int(max_cached_verts),
        _stream()))
    elif depth_state is not None:
      _check(lib.srl_raster_incremental(
        *args[:4], _opt(inst_counts, torch.int32, 'inst_counts'),
        _out(depth_state, torch.float32, (njobs, rows, cols), verts, 'depth_state'),
        int(only_last), args[4], int(njobs), int(rows), int(cols), int(mode),
        float(far_plane), int(max_cached_verts), _stream()))
    else:
      _check(lib.srl_raster_ex(*args[:4], _opt(inst_counts, torch.int32, 'inst_counts'), args[4],
                               int(njobs), int(rows), int(cols), int(mode), float(far_plane),
                               int(max_cached_verts), _stream()))
  return out


def reward_sums(walls, goals, goal_z):
  """-> (intersection [E], union [E], goal volume [E]) float32."""
  E, H, W = walls.shape
  dev = walls.device
  args = (_dev(walls, torch.float32, 'walls'), _dev(goals, torch.float32, 'goals'),
          _dev(goal_z, torch.float32, 'goal_z'))
  inter = torch.empty((E,), dtype=torch.float32, device=dev)
  uni = torch.empty_like(inter)
  vol = torch.empty_like(inter)
  with torch.cuda.device(dev):
    _check(lib.srl_reward_sums_f32(*args, _P(inter.data_ptr()), _P(uni.data_ptr()),
                                   _P(vol.data_ptr()), E, H, W, _stream()))
  return inter, uni, vol


def pack_obs(walls, goals, rocks, dtype='float32', scale=1., repeat_wall=False, out=None):
  """Planar maps -> reference-layout observation tensors
  ([E,(R,)H,W,2], [E,R,h,h,1]) in float32 or uint8.  ``out``: (wall_goal, rock)
  buffers to fill instead of new tensors."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  dev = walls.device
  tdt = {'float32': torch.float32, 'uint8': torch.uint8}[str(dtype)]
  code = 0 if tdt == torch.float32 else 1
  wg_shape = (E, R, H, W, 2) if repeat_wall else (E, H, W, 2)
  args = (_dev(walls, torch.float32, 'walls'), _dev(goals, torch.float32, 'goals'),
          _dev(rocks, torch.float32, 'rocks'))
  if out is None:
    wall_goal = torch.empty(wg_shape, dtype=tdt, device=dev)
    rock = torch.empty((E, R, h, h, 1), dtype=tdt, device=dev)
  else:
    wall_goal, rock = out
    _out(wall_goal, tdt, wg_shape, walls, 'wall_goal')
    _out(rock, tdt, (E, R, h, h, 1), walls, 'rock')
  with torch.cuda.device(dev):
    _check(lib.srl_pack_obs(*args, _P(wall_goal.data_ptr()), _P(rock.data_ptr()),
                            E, R, H, W, h, code, float(scale), int(bool(repeat_wall)),
                            _stream()))
  return wall_goal, rock


SRL_E_UNSUPPORTED = -2


def mask_select(values, walls, goals, rocks, minorder=1, overlap_threshold=0.75,
                want_shown=True, want_best=True):
  """goal_overlap + select in one launch (counts stay in shared memory).
  values [E,R,Ph,Pw] f32/f64; walls/goals [E,H,W], rocks [E,R,h,h] f32 or u8 (raw)."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  key = (values.dtype, walls.dtype)
  fn = {(torch.float32, torch.float32): lib.srl_mask_select_f32,
        (torch.float64, torch.float32): lib.srl_mask_select_f64,
        (torch.float64, torch.uint8): lib.srl_mask_select_f64_u8}.get(key)
  if fn is None:
    raise TypeError('mask_select: unsupported dtypes {}'.format(key))
  dev = values.device
  args = (_dev(values, values.dtype, 'values'), _dev(walls, walls.dtype, 'walls'),
          _dev(goals, walls.dtype, 'goals'), _dev(rocks, walls.dtype, 'rocks'))
  actions = torch.empty((E, R), dtype=torch.int64, device=dev)
  shown = torch.empty(tuple(values.shape), dtype=torch.float64, device=dev) if want_shown else None
  best = torch.empty((E, 2), dtype=torch.int64, device=dev) if want_best else None
  with torch.cuda.device(dev):
    _check(fn(*args, _dev(actions, torch.int64, 'actions'),
              _opt(shown, torch.float64, 'shown'), _opt(best, torch.int64, 'best'),
              E, R, H, W, h, int(minorder), float(overlap_threshold), _stream()))
  return actions, shown, best


def score_f32(walls, goals, rocks, level=None, level_mode=2, minorder=1,
              overlap_threshold=0.75, want_values=True, want_best=True):
  """Fused scoring (one launch): -> (values [E,R,Ph,Pw] f32 | None, actions [E,R],
  best [E,2] | None).  Raises SrlError with code SRL_E_UNSUPPORTED for shapes the
  fused kernel does not cover."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  dev = walls.device
  args = (_dev(walls, torch.float32, 'walls'), _opt(goals, torch.float32, 'goals'),
          _dev(rocks, torch.float32, 'rocks'), _opt(level, torch.float32, 'level'))
  values = torch.empty((E, R, H - h + 1, W - h + 1), dtype=torch.float32, device=dev) \
    if want_values else None
  actions = torch.empty((E, R), dtype=torch.int64, device=dev)
  best = torch.empty((E, 2), dtype=torch.int64, device=dev) if want_best else None
  with torch.cuda.device(dev):
    _check(lib.srl_score_f32(*args, _opt(values, torch.float32, 'values'),
                             _dev(actions, torch.int64, 'actions'),
                             _opt(best, torch.int64, 'best'), E, R, H, W, h,
                             int(level_mode), int(minorder), float(overlap_threshold),
                             _stream()))
  return values, actions, best


def siam_correlation_f32(x, w, out=None):
  """Siamese correlation layer (nets/layers.py:21-38): x [B,H,W,C], w [B,h,wd,C]
  float32 channels-last -> [B,H-h+1,W-wd+1,1] float32 (per-sample VALID cross-
  correlation summed over the channels)."""
  if x.dim() != 4 or w.dim() != 4 or x.shape[0] != w.shape[0] or x.shape[3] != w.shape[3]:
    raise ValueError('correlation expects [B,H,W,C] and [B,h,w,C], got {} and {}'.format(
      tuple(x.shape), tuple(w.shape)))
  B, H, W, C = x.shape
  h, wd = w.shape[1], w.shape[2]
  if h > H or wd > W:
    raise ValueError('the second input must not be larger than the first')
  args = (_dev(x, torch.float32, 'x'), _dev(w, torch.float32, 'w'))
  if out is None:
    out = torch.empty((B, H - h + 1, W - wd + 1, 1), dtype=torch.float32, device=x.device)
  elif out.numel() != B * (H - h + 1) * (W - wd + 1):
    raise ValueError('out must hold [B, H-h+1, W-w+1, 1] values, got {}'.format(tuple(out.shape)))
  with torch.cuda.device(x.device):
    _check(lib.srl_siam_correlation_f32(*args, _dev(out, torch.float32, 'out'),
                                        B, H, W, C, h, wd, _stream()))
  return out


def siam_correlation_grad_f32(x, w, grad_out, want_x=True, want_w=True):
  """Vector-Jacobian products of the correlation layer: grad_out [B,Ph,Pw(,1)] ->
  (grad_x like x | None, grad_w like w | None)."""
  B, H, W, C = x.shape
  h, wd = w.shape[1], w.shape[2]
  if grad_out.numel() != B * (H - h + 1) * (W - wd + 1):
    raise ValueError('grad_out must hold [B, H-h+1, W-w+1] values, got {}'.format(
      tuple(grad_out.shape)))
  _same_device(x, w, grad_out)
  gx = torch.empty_like(x) if want_x else None
  gw = torch.empty_like(w) if want_w else None
  with torch.cuda.device(x.device):
    _check(lib.srl_siam_correlation_grad_f32(
      _dev(x, torch.float32, 'x'), _dev(w, torch.float32, 'w'),
      _dev(grad_out, torch.float32, 'grad_out'), _opt(gx, torch.float32, 'grad_x'),
      _opt(gw, torch.float32, 'grad_w'), B, H, W, C, h, wd, _stream()))
  return gx, gw


def correlate_f32(walls, rocks, level=None, want_corr=True, want_coef=True):
  """-> (correlate [E,R,Ph,Pw] f32 | None, corrcoef [E,R,Ph,Pw] f32 | None)."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  dev = walls.device
  args = (_dev(walls, torch.float32, 'walls'), _dev(rocks, torch.float32, 'rocks'),
          _opt(level, torch.float32, 'level'))
  shape = (E, R, H - h + 1, W - h + 1)
  corr = torch.empty(shape, dtype=torch.float32, device=dev) if want_corr else None
  coef = torch.empty(shape, dtype=torch.float32, device=dev) if want_coef else None
  with torch.cuda.device(dev):
    _check(lib.srl_correlate_f32(*args, _opt(corr, torch.float32, 'corr'),
                                 _opt(coef, torch.float32, 'coef'), E, R, H, W, h, _stream()))
  return corr, coef


# ---- device-side environment step (include/stackrl_b200.h: srl_env_state) ---------- #
class EnvStateStruct(ctypes.Structure):
  _fields_ = [('E', _c.c_int32), ('capacity', _c.c_int32), ('length', _c.c_int32),
              ('reserved', _c.c_int32)] + [(name, _P) for name in (
                'mesh_ranges', 'mesh_coms', 'spawn_rows', 'instances', 'counts', 'order',
                'cursor', 'current', 'hist_rest', 'hist_placed', 'hist_mesh', 'n_placed',
                'done', 'rock_instances', 'memory')]


class EnvState(object):
  """Owner of the device buffers behind one ``srl_env_state`` (E environments,
  ``capacity`` instance slots each, episodes of ``length`` rocks)."""

  def __init__(self, E, capacity, length, mesh_ranges, mesh_coms, spawn_rows, device):
    isz = INSTANCE_DTYPE.itemsize
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
    self.E, self.capacity, self.length = int(E), int(capacity), int(length)
    self.mesh_ranges = mesh_ranges          # [M,4] i32
    self.mesh_coms = mesh_coms              # [M,3] f64
    self.spawn_rows = spawn_rows            # [M,14] f64 (instance structs)
    self.instances = z((self.E * self.capacity * isz,), torch.uint8)
    self.counts = z((self.E,), torch.int32)
    self.order = z((self.E, self.length), torch.int32)
    self.cursor = z((self.E,), torch.int32)
    self.current = z((self.E,), torch.int32)
    self.hist_rest = z((self.E, self.length, 7), torch.float64)
    self.hist_placed = z((self.E, self.length, 7), torch.float64)
    self.hist_mesh = z((self.E, self.length), torch.int32)
    self.n_placed = z((self.E,), torch.int32)
    self.done = torch.ones((self.E,), dtype=torch.uint8, device=device)
    self.rock_instances = z((self.E * isz,), torch.uint8)
    self.memory = z((self.E, 4), torch.float64)
    self.device = device
    self.struct = EnvStateStruct(
      self.E, self.capacity, self.length, 0,
      *[getattr(self, name).data_ptr() for name, _ in EnvStateStruct._fields_[4:]])

  def ref(self):
    return ctypes.byref(self.struct)


def place_poses(walls, rocks, views, flat, orientations, geometry, threshold=1e-4,
                poses=None, status=None):
  """Observer.pose for a batch (observer.py:392-421).  ``views`` [E] int64 or None,
  ``flat`` [E] int64 device tensors; ``orientations`` [R,4] float64; ``geometry`` =
  (pixel_h, pixel_w, object_x, object_y, object_z).  -> (poses [E,7] float64,
  status [E] int32: set to 1 where the action was out of range, env.py:237; a
  caller-provided ``status`` is never cleared, so it accumulates over steps)."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  dev = walls.device
  _same_device(walls, rocks, views, flat, orientations, poses, status)
  if poses is None:
    poses = torch.empty((E, 7), dtype=torch.float64, device=dev)
  if status is None:
    status = torch.zeros((E,), dtype=torch.int32, device=dev)
  if flat.shape != (E,) or (views is not None and views.shape != (E,)):
    raise ValueError('actions must be [E] tensors')
  if orientations.shape != (R, 4):
    raise ValueError('orientations must be [R, 4]')
  # [E] tensors, or equally strided column views of one table (the `best` [E,2]
  # tensor of the selection kernels): the kernel reads them in place.
  stride = flat.stride(0) if E > 1 else 1
  for name, t in (('views', views), ('flat', flat)):
    if t is None:
      continue
    if not t.is_cuda or t.dtype != torch.int64:
      raise TypeError('{} must be an int64 CUDA tensor'.format(name))
    if E > 1 and t.stride(0) != stride:
      raise ValueError('views and flat must have the same stride')
  if stride < 1:
    raise ValueError('actions must have a positive stride')
  with torch.cuda.device(dev):
    _check(lib.srl_place_poses_f32(
      _dev(walls, torch.float32, 'walls'), _dev(rocks, torch.float32, 'rocks'),
      _P(None) if views is None else _P(views.data_ptr()), _P(flat.data_ptr()),
      _dev(orientations, torch.float64, 'orientations'),
      _out(poses, torch.float64, (E, 7), walls, 'poses'),
      _out(status, torch.int32, (E,), walls, 'status'), E, R, H, W, h, int(stride),
      *[float(x) for x in geometry], float(threshold), _stream()))
  return poses, status


def contact_precheck(walls, rocks, views, flat, eps=2. ** -13, threshold=1e-4):
  """Heightmap contact pre-check of the chosen placements (srl_contact_precheck_f32):
  -> (contacts [E] int32, octants [E] int32 bit mask, supported [E] bool)."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  dev = walls.device
  stride = flat.stride(0) if E > 1 else 1
  for name, t in (('views', views), ('flat', flat)):
    if t is not None and (not t.is_cuda or t.dtype != torch.int64 or
                          (E > 1 and t.stride(0) != stride)):
      raise TypeError('{} must be an int64 CUDA tensor of the common stride'.format(name))
  contacts = torch.empty((E,), dtype=torch.int32, device=dev)
  octants = torch.empty((E,), dtype=torch.int32, device=dev)
  supported = torch.empty((E,), dtype=torch.uint8, device=dev)
  with torch.cuda.device(dev):
    _check(lib.srl_contact_precheck_f32(
      _dev(walls, torch.float32, 'walls'), _dev(rocks, torch.float32, 'rocks'),
      _P(None) if views is None else _P(views.data_ptr()), _P(flat.data_ptr()),
      _P(contacts.data_ptr()), _P(octants.data_ptr()), _P(supported.data_ptr()), E, R, H, W, h,
      int(stride), float(threshold), float(eps), _stream()))
  return contacts, octants, supported.view(torch.bool)


def env_reset(state, env_ids=None):
  """Start episodes for ``env_ids`` (int32 device tensor; None: all)."""
  n = state.E if env_ids is None else env_ids.numel()
  with torch.cuda.device(state.device):
    _check(lib.srl_env_reset(state.ref(), _opt(env_ids, torch.int32, 'env_ids'), int(n),
                             _stream()))


def env_advance(state, rest, placed=None):
  with torch.cuda.device(state.device):
    _check(lib.srl_env_advance(state.ref(), _out(rest, torch.float64, (state.E, 7), state.done,
                                                 'rest'),
                               _P(None) if placed is None else
                               _out(placed, torch.float64, (state.E, 7), state.done, 'placed'),
                               _stream()))


def env_set_poses(state, poses):
  """poses [E, n, 7] float64: rest poses of the first n placed rocks."""
  n = int(poses.shape[1])
  with torch.cuda.device(state.device):
    _check(lib.srl_env_set_poses(state.ref(), _out(poses, torch.float64, (state.E, n, 7),
                                                   state.done, 'poses'), n, _stream()))


def env_draw(state, rects, n_meshes, shape, object_shape, goal_size_ratio, seed, episode,
             env_ids=None):
  """Device-side episode draws (srl_env_draw): fills ``state.order`` and ``rects``."""
  H, W = shape
  if not goal_size_ratio:
    mode, size, sh, sw = 0, 0, 0, 0
  elif isinstance(goal_size_ratio, (int, float)):
    mode, size, sh, sw = 1, int(goal_size_ratio * H * W), 0, 0
  else:
    mode, size = 2, 0
    sh, sw = int(goal_size_ratio[0] * H), int(goal_size_ratio[1] * W)
  n = state.E if env_ids is None else env_ids.numel()
  with torch.cuda.device(state.device):
    _check(lib.srl_env_draw(state.ref(), _dev(state.order, torch.int32, 'order'),
                            _out(rects, torch.int32, (state.E, 4), state.done, 'rects'),
                            _opt(env_ids, torch.int32, 'env_ids'), int(n), int(n_meshes),
                            int(H), int(W), int(object_shape[0]), int(object_shape[1]), mode,
                            size, sh, sw, int(seed) & (2 ** 64 - 1), int(episode), _stream()))


def fill_goals(rects, goal_z, goals, env_ids=None):
  """goals[e, u0:u1, v0:v1] = goal_z[e] (rewarder.py:252-258); rects [n,4] int32."""
  n = rects.shape[0]
  E, H, W = goals.shape
  _same_device(rects, goal_z, goals, env_ids)
  if env_ids is None and n != E:
    raise ValueError('rects must be [E, 4] when env_ids is not given')
  with torch.cuda.device(goals.device):
    _check(lib.srl_fill_goals_f32(_dev(rects, torch.int32, 'rects'),
                                  _out(goal_z, torch.float32, (E,), goals, 'goal_z'),
                                  _opt(env_ids, torch.int32, 'env_ids'),
                                  _dev(goals, torch.float32, 'goals'), int(n), H, W, _stream()))
  return goals


def goal_level(goals, out=None):
  """goal.max() per environment (baselines.py:23), float32 or uint8 planes."""
  E = goals.shape[0]
  HW = goals[0].numel()
  if out is None:
    out = torch.empty((E,), dtype=goals.dtype, device=goals.device)
  fn = {torch.float32: lib.srl_goal_level_f32, torch.uint8: lib.srl_goal_level_u8}.get(goals.dtype)
  if fn is None:
    raise TypeError('goal_level takes float32 or uint8 planes, got {}'.format(goals.dtype))
  with torch.cuda.device(goals.device):
    _check(fn(_dev(goals, goals.dtype, 'goals'), _out(out, goals.dtype, (E,), goals, 'level'),
              E, HW, _stream()))
  return out


METRICS = {'iou': 0, 'or': 1, 'diou': 2, 'dor': 3, 'all': 4}


def rewards(state, walls, goals, goal_z, rects, metric, scale, pixel, pmax, pexp, oexp,
            reward=None, value=None):
  """Rewarder.call on the device (rewarder.py:162-179): -> reward [E] (or [E,4] for
  'all') float32; updates ``state.memory``."""
  m = METRICS[metric]
  E, H, W = walls.shape
  shape = (E, 4) if m == 4 else (E,)
  if reward is None:
    reward = torch.empty(shape, dtype=torch.float32, device=walls.device)
  with torch.cuda.device(walls.device):
    _check(lib.srl_rewards_f32(
      state.ref(), _dev(walls, torch.float32, 'walls'), _dev(goals, torch.float32, 'goals'),
      _dev(goal_z, torch.float32, 'goal_z'), _dev(rects, torch.int32, 'rects'),
      _out(reward, torch.float32, shape, walls, 'reward'), _opt(value, torch.float64, 'value'),
      H, W, m, float(scale), float(pixel[0]), float(pixel[1]), float(pmax),
      -1.0 if pexp is None else float(pexp), -1.0 if oexp is None else float(oexp), _stream()))
  return reward


def pack_rewards(state, walls, goals, rocks, goal_z, rects, metric, scale, pixel, pmax, pexp,
                 oexp, dtype='float32', obs_scale=1., repeat_wall=False, out=None, rows=None,
                 full=None):
  """pack_obs + rewards of one step in one launch -> (wall_goal, rock, reward).
  ``goals=None``: the goal maps are the rectangles ``rects`` at ``goal_z`` (what
  ``fill_goals`` writes) and are not read from memory.  ``rows`` [E,2] int32 (+ ``full``
  [E] uint8, + ``out``): ``out`` is persistent and only the wall rows that changed are
  rewritten (srl_pack_rewards_rows_f32)."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  dev = walls.device
  m = METRICS[metric]
  tdt = {'float32': torch.float32, 'uint8': torch.uint8}[str(dtype)]
  wg_shape = (E, R, H, W, 2) if repeat_wall else (E, H, W, 2)
  if out is None:
    if rows is not None:
      raise ValueError('row-incremental packing needs the persistent buffers as `out`')
    wall_goal = torch.empty(wg_shape, dtype=tdt, device=dev)
    rock = torch.empty((E, R, h, h, 1), dtype=tdt, device=dev)
  else:
    wall_goal, rock = out
    _out(wall_goal, tdt, wg_shape, walls, 'wall_goal')
    _out(rock, tdt, (E, R, h, h, 1), walls, 'rock')
  reward = torch.empty((E, 4) if m == 4 else (E,), dtype=torch.float32, device=dev)
  head = (state.ref(), _dev(walls, torch.float32, 'walls'),
          _P(None) if goals is None else _dev(goals, torch.float32, 'goals'),
          _dev(rocks, torch.float32, 'rocks'), _dev(goal_z, torch.float32, 'goal_z'),
          _dev(rects, torch.int32, 'rects'))
  tail = (_P(wall_goal.data_ptr()), _P(rock.data_ptr()), _P(reward.data_ptr()), _P(None), R, H,
          W, h, 0 if tdt == torch.float32 else 1, float(obs_scale), int(bool(repeat_wall)), m,
          float(scale), float(pixel[0]), float(pixel[1]), float(pmax),
          -1.0 if pexp is None else float(pexp), -1.0 if oexp is None else float(oexp), _stream())
  with torch.cuda.device(dev):
    if rows is not None:
      _check(lib.srl_pack_rewards_rows_f32(
        *head, _out(rows, torch.int32, (E, 2), walls, 'rows'),
        _P(None) if full is None else _out(full, torch.uint8, (E,), walls, 'full'), *tail))
    else:
      _check(lib.srl_pack_rewards_f32(*head, *tail))
  return wall_goal, rock, reward


def gather_rows(table, index, out):
  """out[e] = table[index[e]] (float32 rows; int32 index on the device): the cached image
  of the rock every environment spawned."""
  n = int(table.shape[0])
  row = table[0].numel()
  E = int(index.numel())
  _same_device(table, index, out)
  with torch.cuda.device(table.device):
    _check(lib.srl_gather_rows_f32(
      _dev(table, torch.float32, 'table'), _dev(index, torch.int32, 'index'),
      _out(out, torch.float32, (E, row), table), E, row, n, _stream()))
  return out


def quantise_planes(walls, goals, rocks, scale, out=None):
  """StackEnv._return's uint8 cast (env.py:171-178) on planar maps -> uint8 planes."""
  E, R, H, W, h = _batch_dims(walls, rocks)
  dev = walls.device
  if out is None:
    out = (torch.empty((E, H, W), dtype=torch.uint8, device=dev),
           torch.empty((E, H, W), dtype=torch.uint8, device=dev),
           torch.empty((E, R, h, h), dtype=torch.uint8, device=dev))
  with torch.cuda.device(dev):
    _check(lib.srl_quantise_planes_u8(
      _dev(walls, torch.float32, 'walls'), _dev(goals, torch.float32, 'goals'),
      _dev(rocks, torch.float32, 'rocks'), _out(out[0], torch.uint8, (E, H, W), walls),
      _out(out[1], torch.uint8, (E, H, W), walls), _out(out[2], torch.uint8, (E, R, h, h), walls),
      E, R, H, W, h, float(scale), _stream()))
  return out


def microbench_addmax(variant, iters=2000):
  """(add, max) cells per second of the issue-rate micro-benchmark."""
  v = _c.c_double(0.)
  _check(lib.srl_microbench_addmax(int(variant), int(iters), ctypes.byref(v)))
  return v.value


def microbench_fma(variant, iters=2000):
  """FP32 FMAs per second of the FMA micro-benchmark (0 FFMA, 1 FFMA2 with a shared
  multiplicand, 2 FFMA2 with distinct operands)."""
  v = ctypes.c_double(0.)
  _check(lib.srl_microbench_fma(int(variant), int(iters), ctypes.byref(v)))
  return v.value
