"""Multi-GPU sharding of the environment batch.

Every environment, rock, rotation and candidate map is independent (the
reference already runs environments in unrelated processes,
stackrl/envs/utils.py:424-448), so the batch is split into contiguous blocks of
the environment index, one block per rank, and the hot path runs with NO
collective.  The only communication is a final gather of per-shard statistics
(timings, counts, checksums) with one all-gather -- NCCL over NVLink on GPUs,
gloo in the CPU tests.
"""
import contextlib
import os
import sys

import torch


@contextlib.contextmanager
def _stdout_to_stderr():
  """NCCL prints its version banner on file descriptor 1 when the first
  communicator is created; callers that print machine-readable lines on stdout
  (bench.py) want that on stderr."""
  sys.stdout.flush()
  saved = os.dup(1)
  try:
    os.dup2(2, 1)
    yield
  finally:
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)


def world():
  """(rank, local_rank, world_size) from the torchrun environment."""
  return (int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')),
          int(os.environ.get('WORLD_SIZE', '1')))


def shard_range(total, rank, world_size):
  """Contiguous [begin, end) block of ``total`` items owned by ``rank``; the
  first ``total % world_size`` ranks get one extra item."""
  if not 0 <= rank < world_size:
    raise ValueError('rank {} outside world of {}'.format(rank, world_size))
  base, extra = divmod(int(total), int(world_size))
  begin = rank * base + min(rank, extra)
  return begin, begin + base + (1 if rank < extra else 0)


def init(backend=None, device=None):
  """Initialise torch.distributed from the torchrun environment (no-op for a
  single process).  Returns the process group module or None."""
  rank, local_rank, size = world()
  if size == 1:
    return None
  import torch.distributed as dist
  if not dist.is_initialized():
    if backend is None:
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    kwargs = {}
    if backend == 'nccl' and device is not None:
      kwargs['device_id'] = device
    with _stdout_to_stderr():
      dist.init_process_group(backend, **kwargs)
      if backend == 'nccl':
        # create the communicator now (its banner goes to stderr with the rest)
        t = torch.zeros(1, device=device if device is not None else
                        torch.device('cuda', torch.cuda.current_device()))
        dist.all_reduce(t)
        torch.cuda.synchronize()
  return dist


def gather_stats(values, device=None):
  """All-gather a small 1-D float64 vector of per-shard statistics.
  Returns a [world_size, n] CPU tensor on every rank."""
  import torch.distributed as dist
  t = torch.as_tensor(values, dtype=torch.float64).reshape(-1)
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
    return t[None].clone()
  if dist.get_backend() == 'nccl':
    t = t.to(device if device is not None else torch.device('cuda', torch.cuda.current_device()))
  out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
  dist.all_gather(out, t)
  return torch.stack(out).cpu()


def checksum(tensor):
  """Order-independent 64-bit checksum of a tensor's bytes (sum of the raw
  32-bit words, wrapped): equal shards give equal sums on any rank layout."""
  flat = tensor.contiguous().view(-1)
  raw = flat.view(torch.uint8)
  pad = (-raw.numel()) % 4
  if pad:
    raw = torch.cat([raw, raw.new_zeros(pad)])
  return int(raw.view(torch.int32).to(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF
