"""Siamese correlation layer (stackrl/nets/layers.py:21-38) through
the C ABI against the float64 oracle.  Tolerance 1e-5 of the largest output (float32
products; TensorFlow's summation order is outside the reference tree)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope='module')
def mods():
  import torch
  from oracle import nets_np
  from stackrl_b200 import capi, nets
  return torch, nets, capi, nets_np


def _check(mods, B, H, W, C, h, w, seed=0, relu=False):
  torch, nets, capi, nets_np = mods
  rng = np.random.default_rng(seed)
  x = rng.standard_normal((B, H, W, C)).astype('float32')
  f = rng.standard_normal((B, h, w, C)).astype('float32')
  if relu:
    x, f = np.maximum(x, 0), np.maximum(f, 0)
  got = nets.correlation(torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda())
  assert tuple(got.shape) == (B, H - h + 1, W - w + 1, 1) and got.dtype == torch.float32
  want = nets_np.correlation(x, f)
  err = np.abs(got.cpu().numpy().astype('float64') - want).max()
  assert err <= TOL * np.abs(want).max(), (err, np.abs(want).max())
  return got


# the DQN's default geometry (config.gin:55: 16 channels, 128x128 wall, 32x32 rock), the
# PseudoSiamFCN branch output (64 channels), odd / ragged shapes, channel counts that
# are not a multiple of 4, a filter as large as the image, 1x1 filters, several column
# blocks (wide images) and row bands
@pytest.mark.parametrize('shape', [
  (3, 128, 128, 16, 32, 32), (2, 64, 64, 64, 16, 16), (2, 32, 32, 8, 16, 16),
  (4, 19, 23, 3, 5, 7), (2, 20, 20, 1, 20, 20), (3, 9, 9, 6, 1, 1),
  (1, 40, 260, 5, 9, 12), (1, 150, 30, 2, 3, 3), (2, 33, 47, 7, 10, 3),
])
def test_matches_oracle(mods, shape):
  _check(mods, *shape)


def test_relu_features_and_large_batch(mods):
  # non-negative feature maps (what follows a ReLU): sums of 16384 positive products
  _check(mods, 40, 128, 128, 16, 32, 32, seed=3, relu=True)


def test_known_answer_and_no_flip(mods):
  torch, nets, capi, nets_np = mods
  x = torch.arange(9, dtype=torch.float32).reshape(1, 3, 3, 1).cuda()
  f = torch.ones((1, 2, 2, 1), dtype=torch.float32).cuda()
  assert nets.correlation(x, f)[0, :, :, 0].cpu().tolist() == [[8, 12], [20, 24]]
  f[0, 0, 0, 0] = 0
  assert nets.correlation(x, f)[0, :, :, 0].cpu().tolist() == [[8, 11], [17, 20]]


def test_linearity_full_size(mods):
  # size-independent property: corr(x, a f + g) = a corr(x, f) + corr(x, g)
  torch, nets, capi, nets_np = mods
  g = torch.Generator(device='cuda').manual_seed(5)
  x = torch.randn((64, 128, 128, 16), device='cuda', generator=g)
  f = torch.randn((64, 32, 32, 16), device='cuda', generator=g)
  h = torch.randn((64, 32, 32, 16), device='cuda', generator=g)
  lhs = nets.correlation(x, 2 * f + h)
  rhs = 2 * nets.correlation(x, f) + nets.correlation(x, h)
  assert (lhs - rhs).abs().max().item() <= 1e-4 * rhs.abs().max().item()


def test_empty_batch_and_bad_arguments(mods):
  torch, nets, capi, nets_np = mods
  out = nets.correlation(torch.empty((0, 8, 8, 4), device='cuda'),
                         torch.empty((0, 3, 3, 4), device='cuda'))
  assert tuple(out.shape) == (0, 6, 6, 1)
  x = torch.zeros((1, 8, 8, 4), device='cuda')
  with pytest.raises(ValueError):
    nets.correlation(x, torch.zeros((1, 9, 3, 4), device='cuda'))     # filter taller than image
  with pytest.raises(ValueError):
    nets.correlation(x, torch.zeros((1, 3, 3, 5), device='cuda'))     # channel mismatch
  with pytest.raises(TypeError):
    nets.correlation(x.cpu(), torch.zeros((1, 3, 3, 4)))              # no CPU path
  with pytest.raises(TypeError):
    nets.correlation(x.double(), torch.zeros((1, 3, 3, 4), device='cuda').double())


# ---- round 2: tensor-core path (tcgen05, 3xTF32) and gradients ------------------------ #
TC_SHAPES = [
  (3, 128, 128, 16, 32, 32),     # config.gin geometry, one CTA band per sample... or more
  (1, 128, 128, 16, 32, 32),     # a single sample: several row bands
  (2, 32, 32, 8, 16, 16),
  (1, 40, 300, 8, 9, 12),        # three 128-pixel column blocks, filter rows padded to 16
  (5, 64, 64, 16, 16, 16),
  (2, 50, 45, 24, 20, 7),        # ragged everything, 6 chunks per pixel
  (1, 33, 33, 8, 32, 2),         # two output rows
]


@pytest.mark.parametrize('shape', TC_SHAPES)
def test_tensor_core_path_matches_oracle_and_fp32_path(mods, monkeypatch, shape):
  """srl_siam_correlation_f32 on tcgen05 (hi/lo split, three TF32 products) against the
  float64 oracle at the layer's tolerance, and against the FP32 FMA kernel."""
  torch, nets, capi, nets_np = mods
  B, H, W, C, h, w = shape
  rng = np.random.default_rng(7)
  x = rng.standard_normal((B, H, W, C)).astype('float32')
  f = rng.standard_normal((B, h, w, C)).astype('float32')
  xd, fd = torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda()
  monkeypatch.setenv('SRL_SIAM_MODE', '2')
  tc = nets.correlation(xd, fd).cpu().numpy().astype('float64')
  monkeypatch.setenv('SRL_SIAM_MODE', '0')
  fp = nets.correlation(xd, fd).cpu().numpy().astype('float64')
  want = nets_np.correlation(x, f)
  scale = np.abs(want).max()
  assert np.abs(tc - want).max() <= TOL * scale, (np.abs(tc - want).max(), scale)
  assert np.abs(fp - want).max() <= TOL * scale
  assert np.abs(tc - fp).max() <= TOL * scale
  # second launch: the pipeline leaves no state behind
  monkeypatch.setenv('SRL_SIAM_MODE', '2')
  assert np.array_equal(nets.correlation(xd, fd).cpu().numpy().astype('float64'), tc)


def test_tensor_core_path_exact_on_small_integers(mods, monkeypatch):
  """Integer-valued features: every product and sum is exact in TF32 / float32, so the
  tensor-core result must equal the oracle bit for bit (catches layout / descriptor
  errors that a tolerance could hide)."""
  torch, nets, capi, nets_np = mods
  monkeypatch.setenv('SRL_SIAM_MODE', '2')
  rng = np.random.default_rng(1)
  for shape in [(2, 128, 128, 16, 32, 32), (1, 40, 300, 8, 9, 12), (3, 24, 24, 8, 5, 5)]:
    B, H, W, C, h, w = shape
    x = rng.integers(-3, 4, (B, H, W, C)).astype('float32')
    f = rng.integers(-3, 4, (B, h, w, C)).astype('float32')
    got = nets.correlation(torch.from_numpy(x).cuda(), torch.from_numpy(f).cuda())
    assert np.array_equal(got.cpu().numpy().astype('float64'), nets_np.correlation(x, f)), shape


@pytest.mark.parametrize('shape', [(2, 24, 20, 8, 5, 7), (1, 40, 36, 3, 9, 4),
                                   (2, 64, 64, 16, 16, 16), (1, 19, 140, 2, 3, 12)])
def test_gradients_match_oracle(mods, shape):
  torch, nets, capi, nets_np = mods
  B, H, W, C, h, w = shape
  rng = np.random.default_rng(5)
  x = rng.standard_normal((B, H, W, C)).astype('float32')
  f = rng.standard_normal((B, h, w, C)).astype('float32')
  g = rng.standard_normal((B, H - h + 1, W - w + 1, 1)).astype('float32')
  xd = torch.from_numpy(x).cuda().requires_grad_(True)
  fd = torch.from_numpy(f).cuda().requires_grad_(True)
  out = nets.correlation(xd, fd)
  out.backward(torch.from_numpy(g).cuda())
  w0, w1 = nets_np.correlation_grads(x, f, g)
  for got, want in ((xd.grad, w0), (fd.grad, w1)):
    err = np.abs(got.cpu().numpy().astype('float64') - want).max()
    assert err <= TOL * np.abs(want).max(), (err, np.abs(want).max())
  # only one input needs a gradient
  xd2 = torch.from_numpy(x).cuda()
  fd2 = torch.from_numpy(f).cuda().requires_grad_(True)
  nets.correlation(xd2, fd2).sum().backward()
  want1 = nets_np.correlation_grads(x, f, np.ones_like(g))[1]
  assert np.abs(fd2.grad.cpu().numpy() - want1).max() <= TOL * np.abs(want1).max()
  assert xd2.grad is None


def test_gradients_full_size_adjoint_identity(mods):
  """config.gin geometry: <corr(x, f), g> = <x, grad_x(g)> = <f, grad_f(g)> (the layer is
  bilinear), a size-independent check of both backward kernels."""
  torch, nets, capi, nets_np = mods
  gen = torch.Generator(device='cuda').manual_seed(2)
  x = torch.randn((6, 128, 128, 16), device='cuda', generator=gen)
  f = torch.randn((6, 32, 32, 16), device='cuda', generator=gen)
  g = torch.randn((6, 97, 97, 1), device='cuda', generator=gen)
  gx, gf = capi.siam_correlation_grad_f32(x, f, g)
  lhs = (nets.correlation(x, f).double() * g.double()).sum().item()
  assert abs((x.double() * gx.double()).sum().item() - lhs) <= 2e-5 * abs(lhs) + 1e-2
  assert abs((f.double() * gf.double()).sum().item() - lhs) <= 2e-5 * abs(lhs) + 1e-2
