"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU
paths (block partition, position all-gather assembly)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from principia_b200 import sharding


def test_block_ranges_tile_the_index_space():
  for n in (1, 7, 100000, 2**20, 1000003):
    for world in (1, 2, 3, 4, 8):
      ranges = [sharding.block_range(n, r, world) for r in range(world)]
      assert ranges[0][0] == 0 and ranges[-1][1] == n
      for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
        assert e0 == b1
      sizes = [e - b for b, e in ranges]
      assert max(sizes) - min(sizes) <= 1
  assert sharding.equal_block_range(2**20, 3, 8) == (3 * 2**17, 4 * 2**17)
  with pytest.raises(ValueError):
    sharding.equal_block_range(10, 0, 4)


def _worker(rank, world, port, n, results):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    rng = np.random.default_rng(5)
    truth = torch.from_numpy(rng.normal(size=4 * n))
    begin, end = sharding.equal_block_range(n, rank, world)
    # Each rank only knows its own rows; the rest is garbage.
    full = torch.full((4 * n,), float('nan'), dtype=torch.float64)
    full[4 * begin:4 * end] = truth[4 * begin:4 * end]
    sharding.gather_sources(full, n, begin, end, dist)
    results[rank] = bool(torch.equal(full, truth))
    # Probe sharding needs no collective: the union of the blocks is the batch.
    b, e = sharding.block_range(1001, rank, world)
    counts = torch.tensor([e - b])
    dist.all_reduce(counts)
    results[world + rank] = int(counts.item()) == 1001
  finally:
    dist.destroy_process_group()


def test_position_all_gather_over_gloo():
  world = 2
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  manager = mp.Manager()
  results = manager.dict()
  mp.spawn(_worker, args=(world, port, 64, results), nprocs=world, join=True)
  assert all(results[k] for k in range(2 * world)), dict(results)
