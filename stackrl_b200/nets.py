"""Host-side mirror of the one ``stackrl.nets`` function on the widened path
(SURVEY 8f rank 2): the Siamese correlation layer that scores every placement
of the rock features over the wall features inside the DQN.

Reference: ``stackrl.nets.correlation`` (stackrl/nets/layers.py:21-38), a Keras
Lambda around ``tf.map_fn`` of one ``tf.nn.conv2d`` per sample, called from
PseudoSiamFCN / DeepQSiamFCN (stackrl/nets/models.py:89, 182).  Here it is one
launch of ``srl_siam_correlation_f32`` over the whole batch.  The rest of the
networks (U-Nets, position layers, dueling head) stays the reference's.
"""
import torch

from stackrl_b200 import capi


def correlation(in0, in1, parallel_iterations=None):
  """``in0`` [B,H,W,C], ``in1`` [B,h,w,C] (float32 CUDA tensors, channels-last
  like the reference's) -> [B,H-h+1,W-w+1,1]: for every sample the VALID
  cross-correlation of in0 with in1 used as the filter, summed over channels.

  ``parallel_iterations`` is the reference's ``tf.map_fn`` knob; the batch is
  always one launch here, the argument is accepted and ignored."""
  del parallel_iterations
  if not (isinstance(in0, torch.Tensor) and isinstance(in1, torch.Tensor)):
    raise TypeError('correlation takes CUDA tensors (stackrl_b200 has no CPU path)')
  if in0.dtype != torch.float32 or in1.dtype != torch.float32:
    raise TypeError('correlation is float32 like the reference layer, got {} and {}'.format(
      in0.dtype, in1.dtype))
  return capi.siam_correlation_f32(in0.contiguous(), in1.contiguous())
