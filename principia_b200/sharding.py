"""Host-side sharding of the path across the GPUs of one box (one process per
GPU, torch.distributed for the plumbing).

* Massless probes and ensemble members are independent: contiguous blocks per
  rank, no data-path collective.
* Large-N all-pairs: targets split into blocks; every force evaluation needs
  all positions, so the owned slices are all-gathered (NCCL over NVLink) after
  every position update.  The library does it itself -- one in-place
  ncclAllGather on the force stream -- once it has a communicator:
  `NcclCommunicator` creates one through the library's own run-time binding of
  NCCL, the unique id travelling over torch.distributed.  `AllGatherExchange`
  is the callback form of the same exchange over torch.distributed (any
  backend: the gloo tests use it).
"""
import ctypes

import numpy as np

SOURCE_DOUBLES = 4  # x, y, z, mu per body in the library's source array.


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


def gather_sources(full, n, begin, end, dist, group=None):
  """All-gathers rows [begin, end) of the [n][4] source array `full` (a flat
  tensor of 4 n values) in place.  Works on CPU (gloo) and CUDA (nccl)
  tensors; equal block sizes."""
  world = dist.get_world_size(group)
  size = end - begin
  assert size * world == n, (size, world, n)
  owned = full[SOURCE_DOUBLES * begin:SOURCE_DOUBLES * end].clone()
  dist.all_gather_into_tensor(full, owned, group=group)
  return full


class AllGatherExchange:
  """The pb_exchange_positions_t callback of an i-block sharded all-pairs
  system over torch.distributed: one all-gather of the owned rows of the
  source array, issued on the stream the force kernel runs on."""

  def __init__(self, n, dist, device):
    import torch
    self.n = n
    self.dist = dist
    self.device = device
    self.torch = torch
    self.calls = 0

  def __call__(self, sources_pointer, n, begin, end, stream):
    torch = self.torch
    # Wrap the library's device buffer without copying.
    full = _tensor_from_pointer(torch, sources_pointer, SOURCE_DOUBLES * n,
                                self.device)
    # The library launches on the legacy default stream (stream == NULL) and so
    # does torch's default stream: the collective is ordered after the kernel
    # that wrote the rows and before the force kernel.
    gather_sources(full, n, begin, end, self.dist)
    self.calls += 1


class NcclCommunicator:
  """An ncclComm_t created by the library (pb_nccl_comm_create) for the ranks
  of torch.distributed's default group: rank 0 draws the unique id, the others
  receive it through a broadcast of the default group (any backend)."""

  def __init__(self, dist, device_index):
    from . import _capi
    import torch
    self.lib = _capi.library()
    self.handle = None
    rank, world = dist.get_rank(), dist.get_world_size()
    unique_id = ctypes.create_string_buffer(128)
    if rank == 0:
      code = self.lib.pb_nccl_get_unique_id(unique_id)
      if code != _capi.PB_OK:
        raise RuntimeError(_capi.last_error())
    payload = torch.frombuffer(bytearray(unique_id.raw), dtype=torch.uint8)
    if dist.get_backend() == 'nccl':
      payload = payload.to(torch.device('cuda', device_index))
    dist.broadcast(payload, src=0)
    unique_id = ctypes.create_string_buffer(bytes(payload.cpu().tolist()), 128)
    handle = ctypes.c_void_p()
    code = self.lib.pb_nccl_comm_create(unique_id, world, rank, device_index,
                                        ctypes.byref(handle))
    if code != _capi.PB_OK:
      raise RuntimeError(_capi.last_error())
    self.handle = handle

  def close(self):
    if self.handle:
      self.lib.pb_nccl_comm_destroy(self.handle)
      self.handle = None

  def __del__(self):
    self.close()


def _tensor_from_pointer(torch, pointer, count, device):
  """A float64 CUDA tensor aliasing `count` doubles at `pointer`."""
  class _Holder:
    pass
  holder = _Holder()
  holder.__cuda_array_interface__ = {
      'shape': (count,), 'typestr': '<f8', 'data': (int(pointer), False),
      'version': 3, 'strides': None}
  return torch.as_tensor(holder, device=device)
