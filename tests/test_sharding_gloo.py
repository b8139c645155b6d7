"""Host-side multi-rank logic on CPU: world_size-2 gloo processes shard a batch,
gather per-shard statistics, and the shards tile the batch exactly."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from stackrl_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch
from stackrl_b200 import sharding, synth
dist = sharding.init(backend='gloo')
rank, _, size = sharding.world()
E = 37
lo, hi = sharding.shard_range(E, rank, size)
walls = synth.walls(0, E, 8, 8)[lo:hi]
stats = sharding.gather_stats([hi - lo, float(walls.sum()), sharding.checksum(torch.from_numpy(walls)) % 2**40])
if rank == 0:
  print(json.dumps(stats.tolist()))
dist.destroy_process_group()
'''


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def test_shard_ranges_tile_the_batch():
  for total in (0, 1, 7, 4096, 65536, 262144):
    for size in (1, 2, 3, 4, 8):
      blocks = [sharding.shard_range(total, r, size) for r in range(size)]
      assert blocks[0][0] == 0 and blocks[-1][1] == total
      assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
      sizes = [b - a for a, b in blocks]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    sharding.shard_range(10, 2, 2)


def test_two_rank_gloo_stats_gather(tmp_path):
  import json
  script = tmp_path / 'worker.py'
  script.write_text(WORKER.format(root=ROOT))
  env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT=str(_free_port()),
             WORLD_SIZE='2', OMP_NUM_THREADS='1')
  procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE) for r in range(2)]
  outs = [p.communicate(timeout=240) for p in procs]
  assert all(p.returncode == 0 for p in procs), outs
  stats = np.array(json.loads(outs[0][0].decode().strip().splitlines()[-1]))
  assert stats.shape == (2, 3)
  assert stats[:, 0].tolist() == [19, 18]
  from stackrl_b200 import synth
  walls = synth.walls(0, 37, 8, 8)
  np.testing.assert_allclose(stats[:, 1], [walls[:19].sum(), walls[19:].sum()], rtol=1e-6)


def test_single_process_gather_is_identity():
  out = sharding.gather_stats([1., 2., 3.])
  assert out.shape == (1, 3) and out[0].tolist() == [1., 2., 3.]
