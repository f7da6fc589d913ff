#!/usr/bin/env python3
"""Secondary measurements: BASELINE.json configs 3, 4 and 5 (bench.py is the
headline, config 2).  One JSON line per config.

  python bench_configs.py [--config adaptive|history|ensemble|allpairs|all]
  torchrun --nproc-per-node N ... bench_configs.py      (weak scaling; the
      all-pairs config is i-block sharded with an NCCL all-gather)
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))


def dist_setup():
  import torch
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  if world > 1 and not dist.is_initialized():
    dist.init_process_group('nccl',
                            device_id=torch.device('cuda', local_rank))
  return torch, dist, world, rank, local_rank


def max_over_ranks(torch, dist, world, value, device):
  t = torch.tensor([value], dtype=torch.float64, device=device)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())


def sum_over_ranks(torch, dist, world, value, device):
  t = torch.tensor([value], dtype=torch.float64, device=device)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return float(t.item())


def bench_adaptive(args):
  """Config 3: eccentric probes, DormandElMikkawyPrince1986RKN434FM,
  1 mm / 1 mm/s, per-probe step control; probes sharded by rank."""
  torch, dist, world, rank, local_rank = dist_setup()
  from principia_b200 import workloads
  from principia_b200.ephemeris import Ephemeris
  device = torch.device('cuda', local_rank)
  system = workloads.sol_system()
  ephemeris = Ephemeris(system, device=local_rank)
  workloads.upload_sol_ephemeris(ephemeris)
  earth = system.bodies[system.index('Earth')]
  n = args.probes
  state0 = workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                      42 + rank)
  t_final = args.span
  results = []
  for repeat in range(2):
    state, times = state0.copy(), np.zeros((2, n))
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sink = None
    if args.sink_capacity > 0:
      # The plugin's default downsampling (10 m) applied inside the kernel.
      from principia_b200.trajectory import Downsampler
      sink = Downsampler(n, 10.0, args.sink_capacity, device=local_rank)
      sink.append([0.0], np.concatenate([state[0:3], state[6:9]])[None])
    result = ephemeris.flow_with_adaptive_step(t_final, state, times, 1e-3,
                                               1e-3, sink=sink)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    kernel_ms, _ = ephemeris.last_flow_kernel_ms()
    results.append((wall, kernel_ms, result))
  wall, kernel_ms, result = results[-1]
  assert np.all(result['status'] == 0)
  timeline_points = None
  if sink is not None:
    counts = sink.timelines()[0]
    timeline_points = float(counts.mean())
  kernel_s = max_over_ranks(torch, dist, world, kernel_ms * 1e-3, device)
  wall = max_over_ranks(torch, dist, world, wall, device)
  steps = sum_over_ranks(torch, dist, world, float(result['accepted'].sum()),
                         device)
  evaluations = sum_over_ranks(torch, dist, world,
                               float(result['evaluations'].sum()), device)
  if rank == 0:
    print(json.dumps({
        'config': 'config 3: eccentric probes, adaptive RKN434FM, 1 mm, 1 mm/s',
        'n_gpus': world, 'probes_per_gpu': n, 'span_seconds': t_final,
        'accepted_steps': steps, 'rhs_evaluations': evaluations,
        'kernel_seconds': kernel_s, 'e2e_seconds': wall,
        'probe_steps_per_second': steps / kernel_s,
        'probe_rkn_stages_per_second': evaluations / kernel_s,
        'e2e_probe_rkn_stages_per_second': evaluations / wall,
        'mean_steps_per_probe': steps / (n * world),
        'max_steps_per_probe': int(result['accepted'].max()),
        'sink_mean_timeline_points': timeline_points}))


def bench_single(args):
  """Config 1 (the reference's own CPU-runnable case, ephemeris_benchmark.cpp
  :181-248): the Sol ephemeris prolonged over `--span` seconds (100 days in
  the benchmark) and ONE LEO probe flowed through it, SRKN14A at 10 s.  The
  ephemeris is produced by the product (massive-body flow and Newhall fit on
  the device) and compared bit for bit with the oracle's.  Sequential in time:
  it does not shard (replicas only) and is latency-bound on a GPU; the oracle
  on one host core is timed beside it."""
  torch, dist, world, rank, local_rank = dist_setup()
  if rank != 0:
    return
  from principia_b200 import _capi
  from principia_b200 import workloads
  from principia_b200.ephemeris import Ephemeris
  from oracle_lib import OracleEphemeris
  system = workloads.sol_system()
  span = args.span
  ephemeris = Ephemeris(system, device=local_rank)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  prolong_steps = workloads.prolong_ephemeris(ephemeris, span)
  torch.cuda.synchronize()
  gpu_prolong_seconds = time.perf_counter() - t0
  oracle = OracleEphemeris(system)
  t0 = time.perf_counter()
  assert oracle.prolong(span) == 0
  cpu_prolong_seconds = time.perf_counter() - t0
  assert ephemeris.t_max() >= span
  n_segments = sum(len(oracle.segments(body)['t_max'])
                   for body in range(len(system.bodies)))
  # The product's polynomials against the oracle's, through EvaluateAllPositions
  # at random times over the whole span, bit for bit.
  rng = np.random.Generator(np.random.MT19937(1))
  ephemeris_equal = True
  for t in rng.uniform(0.0, span, 64):
    ephemeris_equal = ephemeris_equal and bool(np.array_equal(
        ephemeris.evaluate_all_positions(t).view(np.int64),
        oracle.evaluate_all_positions(t).view(np.int64)))
  earth = system.bodies[system.index('Earth')]
  # Radius 6371 km + 100 nautical miles, circular, in the equatorial plane of
  # the frame (the benchmark's probe).
  r = 6371e3 + 100 * 1852.0
  state0 = np.zeros((12, 1))
  state0[0:3, 0] = np.array(earth['q']) + [r, 0.0, 0.0]
  state0[6:9, 0] = np.array(earth['v']) + [0.0, math.sqrt(earth['mu'] / r), 0.0]
  method = _capi.BLANES_MOAN_2002_SRKN_14A
  steps = int(span / 10.0)
  ephemeris.flow_with_fixed_step(method, 10.0, 100.0, state0.copy(),
                                 np.zeros(2))  # Warm-up.
  state = state0.copy()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  result = ephemeris.flow_with_fixed_step(method, 10.0, span, state,
                                          np.zeros(2))
  gpu_seconds = time.perf_counter() - t0
  assert result['n_steps'] == steps, result
  expected = state0.copy()
  t0 = time.perf_counter()
  oracle.flow_with_fixed_step(method, 10.0, span, expected, np.zeros(2),
                              threads=1)
  cpu_seconds = time.perf_counter() - t0
  bitwise = bool(np.array_equal(state.view(np.int64), expected.view(np.int64)))
  print(json.dumps({
      'config': 'config 1: Sol ephemeris prolonged by the product, one LEO '
                'probe, SRKN14A at 10 s (sequential: replicas only)',
      'span_seconds': span, 'span_days': span / 86400.0,
      'prolong_steps_quinlan_tremaine_12_at_10_min': prolong_steps,
      'gpu_prolong_seconds': gpu_prolong_seconds,
      'cpu_oracle_prolong_seconds': cpu_prolong_seconds,
      'ephemeris_segments_per_body': n_segments / len(system.bodies),
      'ephemeris_bitwise_equal_to_oracle_at_64_times': ephemeris_equal,
      'probe_steps': steps,
      'gpu_seconds': gpu_seconds,
      'gpu_probe_stages_per_second': steps * 14 / gpu_seconds,
      'cpu_oracle_one_core_seconds': cpu_seconds,
      'cpu_probe_stages_per_second': steps * 14 / cpu_seconds,
      'probe_state_bitwise_equal_to_oracle': bitwise,
      'max_position_difference_m':
          float(np.abs(state[0:3] - expected[0:3]).max())}))
  assert bitwise and ephemeris_equal


def bench_history(args):
  """The plugin's history integrator on the config-2 probes: 10^5 LEO probes,
  Quinlan1999Order8A at 10 s (ksp_plugin/integrators.cpp:66-72), one
  right-hand-side evaluation per step; probes sharded by rank."""
  torch, dist, world, rank, local_rank = dist_setup()
  from principia_b200 import _capi
  from principia_b200 import workloads
  from principia_b200.ephemeris import Ephemeris
  device = torch.device('cuda', local_rank)
  system = workloads.sol_system()
  ephemeris = Ephemeris(system, device=local_rank)
  workloads.upload_sol_ephemeris(ephemeris)
  earth = system.bodies[system.index('Earth')]
  n = args.probes
  state = workloads.leo_probes(n, earth['q'], earth['v'], earth['mu'],
                               42 + rank)
  instance = ephemeris.new_instance(_capi.QUINLAN_1999_ORDER_8A, 10.0, state,
                                    np.zeros(2))
  instance.flow(100.0)  # Startup and warm-up.
  t = 100.0
  results = []
  for repeat in range(3):
    t += 10.0 * args.steps
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    result = instance.flow(t)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    kernel_ms, launches = ephemeris.last_flow_kernel_ms()
    results.append((wall, kernel_ms, result['n_steps']))
  wall, kernel_ms, steps = results[-1]
  assert steps == args.steps
  kernel_s = max_over_ranks(torch, dist, world, kernel_ms * 1e-3, device)
  wall = max_over_ranks(torch, dist, world, wall, device)
  if rank == 0:
    from principia_b200 import flops
    per_step = flops.massless_probe_stage_sol_leo()
    print(json.dumps({
        'config': 'history integrator: LEO probes, Quinlan1999Order8A, 10 s',
        'n_gpus': world, 'probes_per_gpu': n, 'steps': steps,
        'kernel_seconds': kernel_s, 'seconds': wall,
        'probe_steps_per_second': n * world * steps / kernel_s,
        'e2e_probe_steps_per_second': n * world * steps / wall,
        'algorithmic_tflops_rhs_only':
            n * world * steps * per_step / kernel_s / 1e12}))


def bench_ensemble(args):
  """Config 4: perturbed TRAPPIST-1 systems, Quinlan1999Order8A, 30 min."""
  torch, dist, world, rank, local_rank = dist_setup()
  from principia_b200 import _capi
  from principia_b200.ephemeris import Ephemeris
  from principia_b200.nbody import NBodyEnsemble
  from principia_b200.solar_system import SolarSystem
  from test_gpu_nbody import ensemble_state
  device = torch.device('cuda', local_rank)
  system = SolarSystem.from_fixture('trappist_jd2457000')
  ephemeris = Ephemeris(system, device=local_rank)
  order, _ = ephemeris.internal_order()
  n = args.systems
  state0 = ensemble_state(system, order, n, 1e-6, seed=42 + rank)
  t0 = system.epoch
  step = 1800.0
  ensemble = NBodyEnsemble(ephemeris, _capi.QUINLAN_1999_ORDER_8A, step,
                           state0, t0)
  torch.cuda.synchronize()
  start = time.perf_counter()
  startup_steps = ensemble.solve(t0 + 7 * step)  # The SRKN14A startup.
  torch.cuda.synchronize()
  startup = time.perf_counter() - start
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  start = time.perf_counter()
  steps = ensemble.solve(t0 + (7 + args.steps) * step)
  torch.cuda.synchronize()
  seconds = max_over_ranks(torch, dist, world, time.perf_counter() - start,
                           device)
  pairs = 28
  if rank == 0:
    print(json.dumps({
        'config': 'config 4: TRAPPIST-1 ensemble, Quinlan1999Order8A, 30 min',
        'n_gpus': world, 'systems_per_gpu': n, 'steady_steps': steps,
        'startup_steps': startup_steps, 'startup_seconds': startup,
        'seconds': seconds,
        'system_steps_per_second': world * n * steps / seconds,
        'pair_interactions_per_second': world * n * steps * pairs / seconds}))


def bench_trajectories(args):
  """ContinuousTrajectory production on the device.  (a) The
  fit alone: n trajectories appended from device memory, one Newhall fit per 8
  appends; HBM-bound (per trajectory and segment: 8 appended points copied
  into the ring, 96 B read + 96 B written each; the fit reads 9 points = 432 B
  and 20 B of state, writes 18 x 3 coefficients + degree = 436 B and the state).
  (b) The TRAPPIST ensemble flow with and without the trajectories as its
  append_state sink."""
  torch, dist, world, rank, local_rank = dist_setup()
  from principia_b200 import _capi
  from principia_b200.ephemeris import Ephemeris
  from principia_b200.nbody import ContinuousTrajectories, NBodyEnsemble
  from principia_b200.solar_system import SolarSystem
  from test_gpu_nbody import ensemble_state
  device = torch.device('cuda', local_rank)
  torch.cuda.set_device(device)
  system = SolarSystem.from_fixture('trappist_jd2457000')
  ephemeris = Ephemeris(system, device=local_rank)
  order, _ = ephemeris.internal_order()
  n_systems = args.systems
  n = n_systems * len(order)
  step = 1800.0
  segments = 8
  # (a) Synthetic circular motions of varied period, already on the device.
  generator = torch.Generator(device=device).manual_seed(42 + rank)
  period = step * 10.0**(1.3 + 3.0 * torch.rand(n, device=device,
                                                dtype=torch.float64,
                                                generator=generator))
  radius = 10.0**(7 + 4 * torch.rand(n, device=device, dtype=torch.float64,
                                     generator=generator))
  omega = 2 * math.pi / period
  trajectories = ContinuousTrajectories(n, step, 1e-3, segments + 1,
                                        device=local_rank)
  points = []
  for k in range(8 * segments + 1):
    angle = omega * (k * step)
    q = torch.stack([radius * torch.cos(angle), radius * torch.sin(angle),
                     torch.zeros_like(angle)]).contiguous()
    v = torch.stack([-radius * omega * torch.sin(angle),
                     radius * omega * torch.cos(angle),
                     torch.zeros_like(angle)]).contiguous()
    points.append((q, v))
  start_event = torch.cuda.Event(enable_timing=True)
  stop_event = torch.cuda.Event(enable_timing=True)
  for k in range(9):  # Warm-up: the first segment.
    q, v = points[k]
    trajectories.append_device(k * step, q.data_ptr(), v.data_ptr())
  torch.cuda.synchronize()
  start_event.record()
  for k in range(9, 8 * segments + 1):
    q, v = points[k]
    trajectories.append_device(k * step, q.data_ptr(), v.data_ptr())
  stop_event.record()
  torch.cuda.synchronize()
  fit_seconds = max_over_ranks(torch, dist, world,
                               start_event.elapsed_time(stop_event) / 1e3,
                               device)
  assert trajectories.status() == 0
  fits = (segments - 1) * n
  bytes_per_fit = 8 * 192 + 432 + 436 + 2 * 20
  degrees = trajectories.segments(0)['degree']
  trajectories.close()
  # (b) The ensemble flow into the trajectories.
  state0 = ensemble_state(system, order, n_systems, 1e-6, seed=42 + rank)
  t0 = system.epoch
  seconds = {}
  for sink in (False, True):
    ensemble = NBodyEnsemble(ephemeris, _capi.QUINLAN_1999_ORDER_8A, step,
                             state0, t0)
    sunk = (ContinuousTrajectories(n, step, 1e-3, 4, device=local_rank)
            if sink else None)
    if sink:
      ensemble.solve_to_trajectories(t0 + 8 * step, sunk)
    else:
      ensemble.solve(t0 + 8 * step)  # Startup and the first segment.
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    start = time.perf_counter()
    if sink:
      steps = ensemble.solve_to_trajectories(t0 + 24 * step, sunk)
    else:
      steps = ensemble.solve(t0 + 24 * step)
    torch.cuda.synchronize()
    seconds[sink] = max_over_ranks(torch, dist, world,
                                   time.perf_counter() - start, device)
    assert steps == 16
    ensemble.close()
  if rank == 0:
    print(json.dumps({
        'config': 'SURVEY 8f rank 1: ContinuousTrajectory production on the '
                  'device (Newhall fit every 8 appended points)',
        'n_gpus': world, 'trajectories_per_gpu': n,
        'timed_segments': segments - 1, 'fit_seconds': fit_seconds,
        'trajectory_fits_per_second': world * fits / fit_seconds,
        'appended_points_per_second': world * fits * 8 / fit_seconds,
        'algorithmic_bytes_per_fit': bytes_per_fit,
        'hbm_gb_per_second': fits * bytes_per_fit / fit_seconds / 1e9,
        'hbm_fraction_of_measured_peak_6553':
            fits * bytes_per_fit / fit_seconds / 1e9 / 6553.0,
        'degrees_of_trajectory_0': [int(d) for d in degrees],
        'ensemble_16_steps_seconds_without_sink': seconds[False],
        'ensemble_16_steps_seconds_with_sink': seconds[True],
        'ensemble_system_steps_per_second_with_sink':
            world * n_systems * 16 / seconds[True]}))


def bench_allpairs(args):
  """Config 5: N point masses, all pairs; i-block sharded, NCCL all-gather of
  the positions before every force evaluation."""
  torch, dist, world, rank, local_rank = dist_setup()
  from principia_b200 import _capi, sharding
  from principia_b200.nbody import NBodyAllPairs
  from test_gpu_nbody import random_cluster
  device = torch.device('cuda', local_rank)
  n = args.bodies
  mu, q, v = random_cluster(n, seed=42)
  begin, end = sharding.equal_block_range(n, rank, world)
  system = NBodyAllPairs(_capi.BLANES_MOAN_2002_SRKN_14A, 3600.0, mu, q, v,
                         0.0, i_begin=begin, i_end=end, device=local_rank)
  communicator = None
  if world > 1:
    communicator = sharding.NcclCommunicator(dist, local_rank)
    system.set_nccl(communicator)
  if args.arithmetic == 'contracted':
    system.set_arithmetic(_capi.PB_ARITHMETIC_CONTRACTED)
  system.accelerations()  # Warm-up.
  torch.cuda.synchronize()
  force_ms = []
  exchange_ms = []
  walls = []
  for _ in range(args.steps):
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    system.accelerations()
    torch.cuda.synchronize()
    walls.append(time.perf_counter() - t0)
    force_ms.append(system.last_force_ms())
    exchange_ms.append(system.last_exchange_ms())
  kernel_s = max_over_ranks(torch, dist, world, float(np.mean(force_ms)) * 1e-3,
                            device)
  wall = max_over_ranks(torch, dist, world, float(np.mean(walls)), device)
  interactions = float(n) * (n - 1)
  flops = interactions * 18
  # One SRKN14A step of the flow (14 force evaluations + all-gathers).
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  steps, forces = system.solve(3600.0)
  torch.cuda.synchronize()
  flow_seconds = max_over_ranks(torch, dist, world, time.perf_counter() - t0,
                                device)
  if rank == 0:
    print(json.dumps({
        'config': 'config 5: all-pairs point masses, i-block sharded',
        'n_gpus': world, 'n_bodies': n, 'arithmetic': args.arithmetic,
        'force_kernel_seconds': kernel_s,
        'exchange_seconds': float(np.mean(exchange_ms)) * 1e-3,
        'force_call_seconds_with_exchange': wall,
        'pair_interactions_per_second': interactions / kernel_s,
        'algorithmic_tflops': flops / kernel_s / 1e12,
        'srkn14a_step_seconds': flow_seconds, 'force_evaluations': forces,
        'flow_pair_interactions_per_second':
            forces * interactions / flow_seconds}))


def main():
  parser = argparse.ArgumentParser()
  parser.add_argument('--config', default='all')
  parser.add_argument('--probes', type=int, default=100000)
  parser.add_argument('--span', type=float, default=86400.0)
  parser.add_argument('--systems', type=int, default=100000)
  parser.add_argument('--steps', type=int, default=20)
  parser.add_argument('--bodies', type=int, default=2**17)
  parser.add_argument('--sink-capacity', type=int, default=0)
  parser.add_argument('--arithmetic', default='exact',
                      choices=['exact', 'contracted'])
  args = parser.parse_args()
  if args.config in ('adaptive', 'all'):
    bench_adaptive(args)
  if args.config in ('single',):
    bench_single(args)
  if args.config in ('history', 'all'):
    bench_history(args)
  if args.config in ('ensemble', 'all'):
    bench_ensemble(args)
  if args.config in ('trajectories', 'all'):
    bench_trajectories(args)
  if args.config in ('allpairs', 'all'):
    bench_allpairs(args)


if __name__ == '__main__':
  main()
