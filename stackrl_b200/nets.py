"""Host-side mirror of the one ``stackrl.nets`` function on the widened path
(SURVEY 8f rank 2): the Siamese correlation layer that scores every placement
of the rock features over the wall features inside the DQN.

Reference: ``stackrl.nets.correlation`` (stackrl/nets/layers.py:21-38), a Keras
Lambda around ``tf.map_fn`` of one ``tf.nn.conv2d`` per sample, called from
PseudoSiamFCN / DeepQSiamFCN (stackrl/nets/models.py:89, 182).  Here it is one
launch of ``srl_siam_correlation_f32`` over the whole batch (tensor cores: tcgen05,
3xTF32) and, for training, ``srl_siam_correlation_grad_f32`` behind
torch.autograd.  The rest of the networks (U-Nets, position layers, dueling head)
stays the reference's.
"""
import torch

from stackrl_b200 import capi


class _Correlation(torch.autograd.Function):
  """The layer with its two vector-Jacobian products, so that a network trains
  through it like through the reference's Lambda(tf.map_fn(conv2d)) (the DQN's
  gradient tape, stackrl/agents/dqn.py, differentiates nets/models.py:89, 182)."""

  @staticmethod
  def forward(ctx, in0, in1):
    in0, in1 = in0.contiguous(), in1.contiguous()
    ctx.save_for_backward(in0, in1)
    return capi.siam_correlation_f32(in0, in1)

  @staticmethod
  def backward(ctx, grad_out):
    in0, in1 = ctx.saved_tensors
    g0, g1 = capi.siam_correlation_grad_f32(
      in0, in1, grad_out.contiguous(), want_x=ctx.needs_input_grad[0],
      want_w=ctx.needs_input_grad[1])
    return g0, g1


def correlation(in0, in1, parallel_iterations=None):
  """``in0`` [B,H,W,C], ``in1`` [B,h,w,C] (float32 CUDA tensors, channels-last
  like the reference's) -> [B,H-h+1,W-w+1,1]: for every sample the VALID
  cross-correlation of in0 with in1 used as the filter, summed over channels.
  Differentiable with respect to both inputs (torch.autograd).

  ``parallel_iterations`` is the reference's ``tf.map_fn`` knob; the batch is
  always one launch here, the argument is accepted and ignored."""
  del parallel_iterations
  if not (isinstance(in0, torch.Tensor) and isinstance(in1, torch.Tensor)):
    raise TypeError('correlation takes CUDA tensors (stackrl_b200 has no CPU path)')
  if in0.dtype != torch.float32 or in1.dtype != torch.float32:
    raise TypeError('correlation is float32 like the reference layer, got {} and {}'.format(
      in0.dtype, in1.dtype))
  if in0.requires_grad or in1.requires_grad:
    return _Correlation.apply(in0, in1)
  return capi.siam_correlation_f32(in0.contiguous(), in1.contiguous())


def benchmark(torch_module, device, samples=148, channels=16, side=128, rock=32, reps=10,
              tensor_peak_tflops=None, tensor_peak_source=''):
  """Throughput of the correlation layer at the config.gin geometry (config.gin:55):
  ``samples`` x (side^2 x channels (*) rock^2 x channels).  Returned as the dict
  bench.py puts under ``extra.siam_correlation``."""
  gen = torch_module.Generator(device=device).manual_seed(0)
  x = torch_module.randn((samples, side, side, channels), device=device, generator=gen)
  w = torch_module.randn((samples, rock, rock, channels), device=device, generator=gen)
  out = torch_module.empty((samples, side - rock + 1, side - rock + 1, 1), device=device)
  for _ in range(3):
    capi.siam_correlation_f32(x, w, out=out)
  a = torch_module.cuda.Event(enable_timing=True)
  b = torch_module.cuda.Event(enable_timing=True)
  torch_module.cuda.synchronize()
  a.record()
  for _ in range(reps):
    capi.siam_correlation_f32(x, w, out=out)
  b.record()
  torch_module.cuda.synchronize()
  ms = a.elapsed_time(b) / reps
  P = side - rock + 1
  flops = 2.0 * samples * P * P * rock * rock * channels
  fma_peak = 2 * max(capi.microbench_fma(v, 400) for v in (0, 1, 2)) / 1e12
  # Tensor-core work the kernel issues (csrc/siam_tc.cu): three products (hi*hi, hi*lo,
  # lo*hi) of every (image row, 128-pixel tile, filter row) -- all `side` image rows meet all
  # `rock` filter rows, and the pixel tiles are padded to 128.
  tiles = (P + 127) // 128
  tensor_flops = 3 * 2.0 * samples * side * (128 * tiles) * rock * rock * channels
  out = {
    'workload': '{} samples, {}x{}x{} wall features * {}x{}x{} rock features, float32 '
                '(stackrl.nets.correlation, config.gin geometry)'.format(
                  samples, side, side, channels, rock, rock, channels),
    'ms': ms, 'samples_per_s': samples / (ms * 1e-3),
    'fp32_equivalent_tflops': flops / (ms * 1e-3) / 1e12,
    'fp32_fma_peak_tflops': fma_peak,
    'fp32_fma_peak_source': 'srl_microbench_fma, best of FFMA / FFMA2 measured in this run '
                            '(nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4): what a '
                            'float32 kernel of this layer is bound by'}
  if tensor_peak_tflops:
    out['roofline'] = {
      'bound': 'tensor', 'kernel': 'siam_tc_kernel (tcgen05, FP16 hi/lo operand pairs)',
      'achieved': tensor_flops / (ms * 1e-3) / 1e12, 'peak': tensor_peak_tflops,
      'unit': 'TFLOP/s', 'frac': tensor_flops / (ms * 1e-3) / 1e12 / tensor_peak_tflops,
      'peak_source': tensor_peak_source,
      'note': 'achieved counts the three error-compensation products and the padded rows / '
              'pixels the kernel really multiplies; the layer itself is fp32_equivalent_tflops. '
              'The limiter is the operand read of an N = 64 MMA from shared memory, not the '
              'tensor pipe (DESIGN 3.8)'}
  return out
