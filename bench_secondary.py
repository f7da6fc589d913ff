"""The `secondary` block of bench.py's JSON line: BASELINE.json configs 3, 4
and 5, each bounded to a few seconds, each with its own roofline from the
static counts of principia_b200/flops.py.  Kernel times are the library's own
CUDA events on the streams it launches on; whole-call times are CUDA events
around host-synchronous calls, bracketed by a barrier + device synchronize,
max over ranks.

  config 3  adaptive RKN434FM, 10^5 eccentric probes per GPU     weak
  config 4  TRAPPIST-1 ensemble, 10^5 systems per GPU            weak
  config 5  all-pairs N = 2^20 split over the ranks, NCCL        STRONG
            all-gather of the positions issued by the library
"""
import json
import os

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))


def measured_hbm_peak():
  """(GB/s, source): MEASURED_PEAKS.json, else B200_PROFILING.md's fallback."""
  try:
    with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  except (OSError, KeyError, ValueError):
    return 6650.0, 'fallback (B200_PROFILING.md)'


class Timer:
  """Device time of a region, max over ranks."""

  def __init__(self, torch, dist, world, device):
    self.torch, self.dist, self.world, self.device = torch, dist, world, device
    self.start_event = torch.cuda.Event(enable_timing=True)
    self.stop_event = torch.cuda.Event(enable_timing=True)

  def barrier(self):
    if self.world > 1:
      self.dist.barrier()
    self.torch.cuda.synchronize()

  def start(self):
    self.barrier()
    self.start_event.record()

  def stop(self):
    """Milliseconds."""
    self.stop_event.record()
    self.barrier()
    return self.max(self.start_event.elapsed_time(self.stop_event))

  def max(self, value):
    t = self.torch.tensor([value], dtype=self.torch.float64,
                          device=self.device)
    if self.world > 1:
      self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
    return float(t.item())

  def sum(self, value):
    t = self.torch.tensor([value], dtype=self.torch.float64,
                          device=self.device)
    if self.world > 1:
      self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
    return float(t.item())


def earth_degree_of_distance(ephemeris, system):
  """r -> the degree at which the Earth's field is truncated at distance r
  (Geopotential::LimitingDegree - 1; 0 beyond the outer threshold of degree
  2): the outer thresholds are three times the inner ones
  (pb_host_model.hpp MakeDamping)."""
  inner, _ = ephemeris.geopotential_damping(system.index('Earth'))
  outer = 3.0 * np.asarray(inner)

  def degree(r):
    # Partition point of r < outer[n] over the non-increasing thresholds.
    count = (r[..., None] < outer).sum(axis=-1)
    d = count - 1
    return np.where(d >= 2, d, 0)
  return degree


def adaptive_flops(ephemeris, system, state, evaluations, polynomial_degree):
  """Estimated algorithmic flops of an adaptive flow: for every probe the
  flops of one right-hand-side evaluation averaged over its osculating orbit
  around the Earth -- the Earth's field is truncated at a degree that depends
  on the distance, and the evaluations crowd where the steps are short, a
  density taken proportional to r^-3/2 per unit time, i.e. to
  (1 - e cos E)^-1/2 per unit eccentric anomaly -- times its evaluation count.
  An ESTIMATE (the kernel does not count its flops)."""
  from principia_b200 import flops
  earth = system.bodies[system.index('Earth')]
  mu = earth['mu']
  q = state[0:3] - np.asarray(earth['q'])[:, None]
  v = state[6:9] - np.asarray(earth['v'])[:, None]
  r = np.linalg.norm(q, axis=0)
  energy = 0.5 * (v * v).sum(axis=0) - mu / r
  a = -0.5 * mu / energy
  h = np.cross(q.T, v.T)
  e = np.sqrt(np.maximum(0.0, 1.0 + 2.0 * energy * (h * h).sum(axis=1) /
                         (mu * mu)))
  anomaly = (np.arange(48) + 0.5) * (2 * np.pi / 48)
  factor = 1.0 - e[:, None] * np.cos(anomaly)[None, :]
  weight = factor**-0.5
  degree = earth_degree_of_distance(ephemeris, system)(a[:, None] * factor)
  table = np.array([flops.adaptive_rhs_evaluation(d, polynomial_degree)
                    for d in range(0, 11)], dtype=np.float64)
  per_evaluation = ((table[np.minimum(degree, 10)] * weight).sum(axis=1) /
                    weight.sum(axis=1))
  return float((per_evaluation * evaluations).sum()), float(
      per_evaluation.mean())


def config3_adaptive(timer, local_rank, rank, world, peak_flops, probes, span):
  from principia_b200 import workloads
  from principia_b200.ephemeris import Ephemeris
  system = workloads.sol_system()
  ephemeris = Ephemeris(system, device=local_rank)
  workloads.prolong_ephemeris(ephemeris, span + 600.0)
  # Mean degree of the Sol polynomials at a 1 mm fitting tolerance (the
  # degrees are per body and per segment: 3 for the outer planets ... 13 for
  # the inner moons).
  polynomial_degree = 9
  earth = system.bodies[system.index('Earth')]
  state0 = workloads.eccentric_probes(probes, earth['q'], earth['v'],
                                      earth['mu'], 42 + rank)
  # Warm-up: a small batch through the same entry point.
  small = np.ascontiguousarray(state0[:, :4096])
  ephemeris.flow_with_adaptive_step(span / 8, small,
                                    np.zeros((2, small.shape[1])), 1e-3, 1e-3)
  state, times = state0.copy(), np.zeros((2, probes))
  timer.start()
  result = ephemeris.flow_with_adaptive_step(span, state, times, 1e-3, 1e-3)
  e2e_ms = timer.stop()
  kernel_ms = timer.max(ephemeris.last_flow_kernel_ms()[0])
  assert np.all(result['status'] == 0)
  evaluations = timer.sum(float(result['evaluations'].sum()))
  steps = timer.sum(float(result['accepted'].sum()))
  flop, per_evaluation = adaptive_flops(ephemeris, system, state0,
                                        result['evaluations'],
                                        polynomial_degree)
  flop = timer.sum(flop)
  achieved = flop / (kernel_ms * 1e-3) / 1e12 / world  # Per GPU.
  ephemeris.close()
  return {
      'config': 'config 3: eccentric probes (a 6.6e6..4e8 m, e <= 0.7) in the '
                'Sol field, adaptive DormandElMikkawyPrince1986RKN434FM, '
                '1 mm / 1 mm/s, per-probe step control',
      'metric': 'adaptive_rhs_evaluations_per_second',
      'value': evaluations / (kernel_ms * 1e-3), 'unit': 'probe-stages/s',
      'scaling': 'weak', 'n_gpus': world, 'probes_per_gpu': probes,
      'span_seconds': span, 'kernel_ms': kernel_ms,
      'accepted_steps': steps, 'rhs_evaluations': evaluations,
      'e2e': {'value': evaluations / (e2e_ms * 1e-3), 'ms': e2e_ms,
              'h2d_bytes': world * 14 * probes * 8,
              'd2h_bytes': world * (14 * 8 + 4 + 3 * 8) * probes},
      'roofline': {
          'bound': 'fp64', 'kernel': 'ErknFlowKernel', 'achieved': achieved,
          'peak': peak_flops / 1e12, 'unit': 'TFLOP/s',
          'frac': achieved / (peak_flops / 1e12),
          'algorithmic_flop_per_evaluation_mean': per_evaluation,
          'flop_model': 'estimate: flops.adaptive_rhs_evaluation averaged '
                        'over each probe\'s osculating orbit (Earth field '
                        'truncated by distance), times its evaluation count'}}


def config4_ensemble(timer, local_rank, rank, world, peak_flops, hbm_peak,
                     systems, steps):
  import sys
  sys.path.insert(0, os.path.join(REPO, 'tests'))
  from principia_b200 import _capi, flops
  from principia_b200.ephemeris import Ephemeris
  from principia_b200.nbody import NBodyEnsemble
  from principia_b200.solar_system import SolarSystem
  system = SolarSystem.from_fixture('trappist_jd2457000')
  ephemeris = Ephemeris(system, device=local_rank)
  order, _ = ephemeris.internal_order()
  n_bodies = len(order)
  # Perturbed copies (tests/test_gpu_nbody.py ensemble_state, restated so that
  # the bench does not import a test module).
  rng = np.random.Generator(np.random.MT19937(42 + rank))
  q = system.initial_positions()[order]
  v = system.initial_velocities()[order]
  state0 = np.zeros((12, n_bodies, systems))
  for c in range(3):
    state0[c] = q[:, c, None] * (1 + 1e-6 * rng.normal(size=(n_bodies,
                                                             systems)))
    state0[6 + c] = v[:, c, None] * (1 + 1e-6 * rng.normal(size=(n_bodies,
                                                                 systems)))
  # Which harmonics are evaluated: an oblate source acts on a target within
  # the outer threshold of its degree 2, damped beyond the inner one.
  active = damped = idle = 0
  for source in range(n_bodies):
    body = system.bodies[int(order[source])]
    if not body.get('oblate'):
      continue
    inner, _ = ephemeris.geopotential_damping(int(order[source]))
    for target in range(n_bodies):
      if target == source:
        continue
      r = np.linalg.norm(q[source] - q[target])
      if len(inner) > 2 and r < 3.0 * inner[2]:
        active += 1
        damped += 1 if r > inner[2] else 0
      else:
        idle += 1
  step = 1800.0
  t0 = system.epoch
  method, order_k = _capi.QUINLAN_1999_ORDER_8A, 8
  ensemble = NBodyEnsemble(ephemeris, method, step, state0, t0)
  timer.start()
  startup_steps = ensemble.solve(t0 + 7 * step + 1.0)
  startup_ms = timer.stop()
  ensemble.solve(t0 + 11 * step + 1.0)  # Warm-up of the multistep kernel.
  timer.start()
  taken = ensemble.solve(t0 + (11 + steps) * step + 1.0)
  ms = timer.stop()
  assert taken == steps and startup_steps == 7
  ensemble.close()
  ephemeris.close()
  system_steps = world * systems * steps
  flop = flops.ensemble_system_step(n_bodies, order_k, active, damped, idle)
  nbytes = flops.ensemble_system_step_bytes(n_bodies, order_k)
  tflops = system_steps * flop / (ms * 1e-3) / 1e12 / world
  gbs = system_steps * nbytes / (ms * 1e-3) / 1e9 / world
  return {
      'config': 'config 4: TRAPPIST-1 ensemble (8 bodies, 7 with J2), '
                'Quinlan1999Order8A at 30 min, 10^5 systems per GPU',
      'metric': 'ensemble_system_steps_per_second',
      'value': system_steps / (ms * 1e-3), 'unit': 'system-steps/s',
      'scaling': 'weak', 'n_gpus': world, 'systems_per_gpu': systems,
      'steps': steps, 'ms': ms, 'startup_ms_7_steps_srkn14a': startup_ms,
      'pair_interactions_per_second':
          system_steps * n_bodies * (n_bodies - 1) / 2 / (ms * 1e-3),
      'roofline': {
          'bound': 'latency: neither pipe is saturated (one launch per step, '
                   'the history ring streamed from HBM, 2 blocks of 8 warps '
                   'per SM)',
          'kernel': 'EnsembleMultistepStepKernel',
          'fp64': {'achieved': tflops, 'peak': peak_flops / 1e12,
                   'unit': 'TFLOP/s', 'frac': tflops / (peak_flops / 1e12),
                   'algorithmic_flop_per_system_step': flop},
          'hbm': {'achieved': gbs, 'peak': hbm_peak[0], 'unit': 'GB/s',
                  'frac': gbs / hbm_peak[0], 'peak_source': hbm_peak[1],
                  'algorithmic_bytes_per_system_step': nbytes},
          'harmonics_per_system_step': {'active': active, 'damped': damped,
                                        'beyond_threshold': idle}}}


def config5_allpairs(timer, torch, dist, local_rank, rank, world, peak_flops,
                     n_bodies, evaluations, flow_bodies):
  from principia_b200 import _capi, flops, sharding
  from principia_b200.nbody import NBodyAllPairs

  def cluster(n, seed):
    # tests/test_gpu_nbody.py random_cluster, restated.
    rng = np.random.Generator(np.random.MT19937(seed))
    au = 149597870700.0
    direction = rng.normal(size=(3, n))
    direction /= np.linalg.norm(direction, axis=0)
    q = au * rng.uniform(0, 1, n)**(1 / 3) * direction
    mu = np.exp(rng.uniform(np.log(1e10), np.log(1e14), n))
    sigma = np.sqrt(mu.sum() / au / 3)
    return mu, q, sigma * rng.normal(size=(3, n))

  communicator = None
  if world > 1:
    communicator = sharding.NcclCommunicator(dist, local_rank)
  results = {}
  for arithmetic, name in ((_capi.PB_ARITHMETIC_EXACT, 'exact'),
                           (_capi.PB_ARITHMETIC_CONTRACTED, 'contracted')):
    mu, q, v = cluster(n_bodies, 42)
    begin, end = sharding.equal_block_range(n_bodies, rank, world)
    system = NBodyAllPairs(_capi.QUINLAN_TREMAINE_1990_ORDER_12, 3600.0, mu,
                           q, v, 0.0, i_begin=begin, i_end=end,
                           device=local_rank)
    system.set_arithmetic(arithmetic)
    if communicator is not None:
      system.set_nccl(communicator)
    # Warm-up on a small system of its own (a force evaluation at this size
    # costs seconds on one GPU), sharded the same way so that the first
    # all-gather of the communicator (connection set-up) is not in the timing.
    small_n = 4096 * world
    small_begin, small_end = sharding.equal_block_range(small_n, rank, world)
    small = NBodyAllPairs(_capi.QUINLAN_TREMAINE_1990_ORDER_12, 3600.0,
                          mu[:small_n], q[:, :small_n], v[:, :small_n], 0.0,
                          i_begin=small_begin, i_end=small_end,
                          device=local_rank)
    small.set_arithmetic(arithmetic)
    if communicator is not None:
      small.set_nccl(communicator)
    small.accelerations()
    small.close()
    force_ms = exchange_ms = 0.0
    timer.start()
    for _ in range(evaluations):
      system.accelerations()
      force_ms += system.last_force_ms()
      exchange_ms += system.last_exchange_ms()
    total_ms = timer.stop()
    force_ms = timer.max(force_ms)
    exchange_ms = timer.max(exchange_ms)
    system.close()
    interactions = float(n_bodies) * (n_bodies - 1) * evaluations
    per_gpu_tflops = (interactions * flops.ALL_PAIRS_INTERACTION /
                      (force_ms * 1e-3) / 1e12 / world)
    results[name] = {
        'value': interactions / (total_ms * 1e-3),
        'ms_per_evaluation': total_ms / evaluations,
        'force_kernel_ms_per_evaluation': force_ms / evaluations,
        'exchange_ms_per_evaluation': exchange_ms / evaluations,
        'roofline': {'bound': 'fp64', 'kernel': 'AllPairsAccelerationKernel',
                     'achieved': per_gpu_tflops, 'peak': peak_flops / 1e12,
                     'unit': 'TFLOP/s',
                     'frac': per_gpu_tflops / (peak_flops / 1e12),
                     'algorithmic_flop_per_interaction':
                         flops.ALL_PAIRS_INTERACTION}}
  entry = {
      'config': 'config 5: all-pairs point masses, N = %d split over the '
                'ranks by target block, one in-place ncclAllGather of the '
                'positions per force evaluation (issued by the library)' %
                n_bodies,
      'metric': 'all_pairs_interactions_per_second',
      'value': results['exact']['value'], 'unit': 'interactions/s',
      'scaling': 'strong', 'n_gpus': world, 'n_bodies': n_bodies,
      'force_evaluations': evaluations,
      'exact': results['exact'], 'contracted': results['contracted']}

  # The flow at a size where a step fits in a second: one BlanesMoan2002SRKN14A
  # step (the starter of the order-12 multistep integrator): 14 force
  # evaluations, 14 all-gathers, strong scaling as well.
  mu, q, v = cluster(flow_bodies, 43)
  begin, end = sharding.equal_block_range(flow_bodies, rank, world)
  system = NBodyAllPairs(_capi.BLANES_MOAN_2002_SRKN_14A, 3600.0, mu, q, v,
                         0.0, i_begin=begin, i_end=end, device=local_rank)
  if communicator is not None:
    system.set_nccl(communicator)
  system.accelerations()
  timer.start()
  steps, forces = system.solve(3600.0)
  flow_ms = timer.stop()
  exchange_ms = timer.max(system.last_exchange_ms())
  force_ms = timer.max(system.last_force_ms())
  system.close()
  assert steps == 1 and forces == 14
  entry['flow'] = {
      'config': 'one SRKN14A step of N = %d: 14 force evaluations and '
                'all-gathers' % flow_bodies,
      'scaling': 'strong', 'n_bodies': flow_bodies, 'ms': flow_ms,
      'force_kernel_ms': force_ms, 'exchange_ms': exchange_ms,
      'value': forces * float(flow_bodies) * (flow_bodies - 1) /
               (flow_ms * 1e-3),
      'unit': 'interactions/s'}
  if communicator is not None:
    communicator.close()
  return entry


def run_secondary(torch, dist, world, rank, local_rank, peak_flops, args):
  device = torch.device('cuda', local_rank)
  timer = Timer(torch, dist, world, device)
  hbm_peak = measured_hbm_peak()
  entries = [
      config3_adaptive(timer, local_rank, rank, world, peak_flops,
                       args.adaptive_probes, 20000.0),
      config4_ensemble(timer, local_rank, rank, world, peak_flops, hbm_peak,
                       args.ensemble_systems, 64),
      config5_allpairs(timer, torch, dist, local_rank, rank, world,
                       peak_flops, args.allpairs_bodies, 2,
                       args.allpairs_flow_bodies),
  ]
  return entries
