"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the
Siamese correlation layer, stackrl/nets/layers.py:21-38:

    tf.map_fn(lambda inps: tf.squeeze(tf.nn.conv2d(inps[0][None], inps[1][..., None],
                                                   strides=1, padding='VALID'), 0), ...)

tf.nn.conv2d is a cross-correlation (no kernel flip) with filter layout
[h, w, in_channels, out_channels]; the rock features become a one-output-channel
filter, so  out[b,i,j,0] = sum_{u,v,c} in0[b,i+u,j+v,c] * in1[b,u,v,c].

PARITY UNPINNED against TensorFlow itself: tensorflow is not installable in the
build container and the reference holds no golden values for this layer; the
restatement follows the published definition of conv2d above and is checked
against scipy.signal.correlate2d and -- forward and both gradients -- against
torch.nn.functional.conv2d + autograd, an independent framework's implementation of the
same cross-correlation (tests/test_oracle_nets.py).  Sums in float64.
"""
import numpy as np


def correlation(in0, in1):
  """in0 [B,H,W,C], in1 [B,h,w,C] -> float64 [B,H-h+1,W-w+1,1]."""
  in0 = np.asarray(in0, dtype='float64')
  in1 = np.asarray(in1, dtype='float64')
  B, H, W, C = in0.shape
  _, h, w, _ = in1.shape
  out = np.zeros((B, H - h + 1, W - w + 1, 1))
  for b in range(B):
    win = np.lib.stride_tricks.sliding_window_view(in0[b], (h, w), axis=(0, 1))  # [Ph,Pw,C,h,w]
    out[b, :, :, 0] = np.einsum('ijcuv,uvc->ij', win, in1[b], optimize=True)
  return out


def correlation_loops(in0, in1):
  """The same by the definition, four nested loops (small cases only)."""
  in0 = np.asarray(in0, dtype='float64')
  in1 = np.asarray(in1, dtype='float64')
  B, H, W, C = in0.shape
  _, h, w, _ = in1.shape
  out = np.zeros((B, H - h + 1, W - w + 1, 1))
  for b in range(B):
    for i in range(H - h + 1):
      for j in range(W - w + 1):
        out[b, i, j, 0] = (in0[b, i:i + h, j:j + w, :] * in1[b]).sum()
  return out


def correlation_grads(in0, in1, grad_out):
  """Vector-Jacobian products of ``correlation`` by the definition (float64):
  grad_in1[b,u,v,c] = sum_{i,j} g[b,i,j] in0[b,i+u,j+v,c];
  grad_in0[b,r,s,c] = sum_{u,v} g[b,r-u,s-v] in1[b,u,v,c]."""
  in0 = np.asarray(in0, dtype='float64')
  in1 = np.asarray(in1, dtype='float64')
  B, H, W, C = in0.shape
  _, h, w, _ = in1.shape
  g = np.asarray(grad_out, dtype='float64').reshape(B, H - h + 1, W - w + 1)
  g0 = np.zeros_like(in0)
  g1 = np.zeros_like(in1)
  for b in range(B):
    for u in range(h):
      for v in range(w):
        patch = in0[b, u:u + H - h + 1, v:v + W - w + 1, :]            # [Ph, Pw, C]
        g1[b, u, v, :] = np.einsum('ij,ijc->c', g[b], patch)
        g0[b, u:u + H - h + 1, v:v + W - w + 1, :] += g[b][:, :, None] * in1[b, u, v, :]
  return g0, g1
