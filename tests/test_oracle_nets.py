"""Oracle of the Siamese correlation layer (stackrl/nets/layers.py:21-38) against the
definition of tf.nn.conv2d(VALID) written as loops and against scipy's correlate2d
(CPU; TensorFlow itself is not installable here: parity unpinned, see oracle/nets_np.py)."""
import numpy as np
import pytest
from scipy import signal

from oracle import nets_np


@pytest.mark.parametrize('shape', [(2, 9, 11, 3, 4, 5), (1, 8, 8, 1, 8, 8), (3, 12, 7, 5, 1, 1)])
def test_einsum_form_equals_loops_and_scipy(shape):
  B, H, W, C, h, w = shape
  rng = np.random.default_rng(11)
  x = rng.standard_normal((B, H, W, C)).astype('float32')
  f = rng.standard_normal((B, h, w, C)).astype('float32')
  got = nets_np.correlation(x, f)
  assert got.shape == (B, H - h + 1, W - w + 1, 1)
  np.testing.assert_allclose(got, nets_np.correlation_loops(x, f), rtol=0, atol=1e-12)
  for b in range(B):
    ref = sum(signal.correlate2d(x[b, :, :, c].astype('float64'), f[b, :, :, c].astype('float64'),
                                 mode='valid') for c in range(C))
    np.testing.assert_allclose(got[b, :, :, 0], ref, rtol=0, atol=1e-12)


def test_known_answer():
  # one channel, 3x3 ramp with a 2x2 filter of ones = 2x2 box sums; no kernel flip
  x = np.arange(9, dtype='float32').reshape(1, 3, 3, 1)
  f = np.ones((1, 2, 2, 1), dtype='float32')
  np.testing.assert_array_equal(nets_np.correlation(x, f)[0, :, :, 0], [[8, 12], [20, 24]])
  f[0, 0, 0, 0] = 0                      # drops the top-left pixel of every window
  np.testing.assert_array_equal(nets_np.correlation(x, f)[0, :, :, 0], [[8, 11], [17, 20]])


def test_mirror_has_no_cpu_path():
  """stackrl_b200.nets.correlation keeps the reference's name and arguments
  (nets/layers.py:21) and refuses anything that is not a float32 CUDA tensor
  instead of computing on the host."""
  import inspect
  import torch
  from stackrl_b200 import nets
  assert list(inspect.signature(nets.correlation).parameters) == [
    'in0', 'in1', 'parallel_iterations']
  x = np.zeros((1, 8, 8, 4), dtype='float32')
  f = np.zeros((1, 3, 3, 4), dtype='float32')
  with pytest.raises(TypeError):
    nets.correlation(x, f)
  with pytest.raises(TypeError):
    nets.correlation(torch.from_numpy(x), torch.from_numpy(f))
  with pytest.raises(TypeError):
    nets.correlation(torch.from_numpy(x).double(), torch.from_numpy(f).double())


def test_gradients_match_finite_differences():
  """The oracle's vector-Jacobian products against central differences of the oracle's
  forward (the layer is bilinear, so the difference quotient is exact up to rounding)."""
  rng = np.random.default_rng(3)
  x = rng.standard_normal((2, 7, 9, 3))
  f = rng.standard_normal((2, 3, 4, 3))
  g = rng.standard_normal((2, 5, 6))
  g0, g1 = nets_np.correlation_grads(x, f, g)
  loss = lambda a, b: float((nets_np.correlation(a, b)[..., 0] * g).sum())
  for idx in [(0, 0, 0, 0), (1, 6, 8, 2), (0, 3, 4, 1)]:
    d = np.zeros_like(x)
    d[idx] = 1e-3
    assert abs((loss(x + d, f) - loss(x - d, f)) / 2e-3 - g0[idx]) < 1e-9
  for idx in [(0, 0, 0, 0), (1, 2, 3, 2), (0, 1, 2, 1)]:
    d = np.zeros_like(f)
    d[idx] = 1e-3
    assert abs((loss(x, f + d) - loss(x, f - d)) / 2e-3 - g1[idx]) < 1e-9


@pytest.mark.parametrize('shape', [(2, 9, 11, 3, 4, 5), (3, 12, 7, 5, 2, 3)])
def test_oracle_equals_an_independent_framework_conv2d_and_its_autograd(shape):
  """TensorFlow is absent, PyTorch is not: torch.nn.functional.conv2d is, like
  tf.nn.conv2d, a cross-correlation (no kernel flip), so `correlation(x, w)[b] =
  conv2d(x[b] as NCHW, w[b] as one output channel)` is the layer of nets/layers.py:21-38
  computed by an independent implementation -- and its autograd gives independent
  gradients for the two inputs (what DQN.train back-propagates, nets/models.py:89, 182).
  float64 on the CPU."""
  import torch
  B, H, W, C, h, w = shape
  rng = np.random.default_rng(5)
  x = rng.standard_normal((B, H, W, C)).astype('float32')
  f = rng.standard_normal((B, h, w, C)).astype('float32')
  g = rng.standard_normal((B, H - h + 1, W - w + 1, 1)).astype('float32')
  xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
  ft = torch.tensor(f, dtype=torch.float64, requires_grad=True)
  outs = [torch.nn.functional.conv2d(xt[b].permute(2, 0, 1)[None], ft[b].permute(2, 0, 1)[None])
          for b in range(B)]
  out = torch.cat(outs, 0).permute(0, 2, 3, 1)              # [B, Ph, Pw, 1]
  np.testing.assert_allclose(nets_np.correlation(x, f), out.detach().numpy(), rtol=0,
                             atol=1e-12)
  (out * torch.tensor(g, dtype=torch.float64)).sum().backward()
  gx, gf = nets_np.correlation_grads(x, f, g)
  np.testing.assert_allclose(gx, xt.grad.numpy(), rtol=0, atol=1e-12)
  np.testing.assert_allclose(gf, ft.grad.numpy(), rtol=0, atol=1e-12)
