"""Host-side sharding of the path across the GPUs of one box (one process per
GPU, torch.distributed for the plumbing).

* Massless probes and ensemble members are independent: contiguous blocks per
  rank, no data-path collective.
* Large-N all-pairs: targets split into blocks; every force evaluation needs
  all positions, so the owned slices are all-gathered (NCCL over NVLink) after
  every position update: `AllGatherExchange`.
"""
import numpy as np


def block_range(n, rank, world):
  """Contiguous block [begin, end) of rank `rank`: sizes differ by at most 1
  and the blocks tile [0, n)."""
  base, extra = divmod(n, world)
  begin = rank * base + min(rank, extra)
  return begin, begin + base + (1 if rank < extra else 0)


def equal_block_range(n, rank, world):
  """Block ranges for collectives that need equal contributions: n must be a
  multiple of the world size."""
  if n % world != 0:
    raise ValueError('n = %d is not a multiple of the world size %d' %
                     (n, world))
  size = n // world
  return rank * size, (rank + 1) * size


def gather_planes(full, local, n, begin, end, dist, group=None):
  """All-gathers the [begin, end) slices of the 3 coordinate planes of `full`
  (a flat tensor of 3 n values) in place.  Works on CPU (gloo) and CUDA (nccl)
  tensors; equal block sizes."""
  import torch
  world = dist.get_world_size(group)
  size = end - begin
  assert size * world == n
  for c in range(3):
    plane = full[c * n:(c + 1) * n]
    dist.all_gather_into_tensor(plane, plane[begin:end].clone(), group=group)
  return full


class AllGatherExchange:
  """The pb_exchange_positions_t callback of an i-block sharded all-pairs
  system: an NCCL all-gather of the owned position slices, issued on the
  stream the force kernel runs on."""

  def __init__(self, n, dist, device):
    import torch
    self.n = n
    self.dist = dist
    self.device = device
    self.torch = torch
    self.calls = 0

  def __call__(self, positions_pointer, n, begin, end, stream):
    torch = self.torch
    # Wrap the library's device buffer without copying.
    from torch.utils import dlpack  # noqa: F401
    full = _tensor_from_pointer(torch, positions_pointer, 3 * n, self.device)
    # The library launches on the legacy default stream (stream == NULL) and so
    # does torch's default stream: the collective is ordered after the kernel
    # that wrote the slices and before the force kernel.
    gather_planes(full, None, n, begin, end, self.dist)
    self.calls += 1


def _tensor_from_pointer(torch, pointer, count, device):
  """A float64 CUDA tensor aliasing `count` doubles at `pointer`."""
  class _Holder:
    pass
  holder = _Holder()
  holder.__cuda_array_interface__ = {
      'shape': (count,), 'typestr': '<f8', 'data': (int(pointer), False),
      'version': 3, 'strides': None}
  return torch.as_tensor(holder, device=device)
