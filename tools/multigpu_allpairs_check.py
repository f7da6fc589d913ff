#!/usr/bin/env python3
"""Run under torchrun with N >= 2 ranks (one per GPU): the i-block sharded
all-pairs flow must reproduce, bit for bit, the slices of the unsharded flow
computed on each rank's own GPU -- with the position all-gather issued by the
library itself (pb_nbody_allpairs_set_nccl: one ncclAllGather per force
evaluation) and with the callback form over torch.distributed.  Run by
tests/test_gpu_full_size.py::test_all_pairs_multi_rank_bit_equality."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))


def main():
  import torch
  import torch.distributed as dist
  from principia_b200 import _capi, sharding
  from principia_b200.nbody import NBodyAllPairs
  from test_gpu_nbody import random_cluster
  rank = int(os.environ['RANK'])
  world = int(os.environ['WORLD_SIZE'])
  local_rank = int(os.environ['LOCAL_RANK'])
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  dist.init_process_group('nccl', device_id=device)
  n = 4096
  mu, q, v = random_cluster(n, seed=11)
  begin, end = sharding.equal_block_range(n, rank, world)
  ok = True
  communicator = sharding.NcclCommunicator(dist, local_rank)
  for method, t_final in ((_capi.BLANES_MOAN_2002_SRKN_14A, 2 * 600.0),
                          (_capi.QUINLAN_1999_ORDER_8A, 10 * 600.0)):
    whole = NBodyAllPairs(method, 600.0, mu, q, v, 0.0, device=local_rank)
    steps_w, forces_w = whole.solve(t_final)
    qw, vw, tw = whole.state()
    for exchange_kind in ('nccl in the library', 'callback'):
      exchange = None
      if exchange_kind == 'callback':
        exchange = sharding.AllGatherExchange(n, dist, device)
      sharded = NBodyAllPairs(method, 600.0, mu, q, v, 0.0, i_begin=begin,
                              i_end=end, device=local_rank, exchange=exchange)
      if exchange is None:
        sharded.set_nccl(communicator)
      steps, forces = sharded.solve(t_final)
      qs, vs, ts = sharded.state()
      same = (steps == steps_w and forces == forces_w and
              np.array_equal(qs.view(np.int64),
                             qw[:, begin:end].view(np.int64))
              and np.array_equal(vs.view(np.int64),
                                 vw[:, begin:end].view(np.int64))
              and np.array_equal(ts, tw) and
              (exchange is None or exchange.calls == forces))
      print('rank %d method %d %s: steps %d forces %d exchange %.3f ms force '
            '%.3f ms bitwise %s' %
            (rank, method, exchange_kind, steps, forces,
             sharded.last_exchange_ms(), sharded.last_force_ms(), same),
            flush=True)
      ok = ok and same
      sharded.close()
    whole.close()
  communicator.close()
  flag = torch.tensor([1.0 if ok else 0.0], device=device)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  dist.destroy_process_group()
  if flag.item() != 1.0:
    raise SystemExit('MISMATCH')
  if rank == 0:
    print('multi-GPU all-pairs check OK on %d ranks' % world)


if __name__ == '__main__':
  main()
