"""GPU parity at the sizes BASELINE.json states (configs 3, 4, 5): the batch is
flowed whole on the device and a subset of it is compared with the CPU oracle
bit for bit -- the trajectories (resp. systems, targets) of a batch are
independent, so the oracle only has to flow the subset.

These are the paths the small parity tests cannot reach: the adaptive kernel's
persistent lanes taking a second, third, ... trajectory from the cost-ordered
queue (pb_kernels_adaptive.cuh, the re-initialisation of first_iteration,
first_stage, hint and the FSAL stage), the CUB cost sort at size, the ensemble
kernels at 3 125 blocks, the all-pairs kernel over 2^20 sources."""
import numpy as np
import pytest

from principia_b200 import _capi, workloads
from principia_b200.ephemeris import Ephemeris
from principia_b200.nbody import (ContinuousTrajectories, NBodyAllPairs,
                                  NBodyEnsemble)
from principia_b200.solar_system import SolarSystem

from conftest import assert_bitwise_equal
from oracle_lib import OracleEphemeris

pytestmark = pytest.mark.gpu


def eccentric(system, n, seed):
  earth = system.bodies[system.index('Earth')]
  return workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                    seed)


def adaptive_lanes():
  """Lanes of one ErknFlowKernel launch: min(ceil(n / 128), SMs x 2) blocks of
  128 threads (pb_api_adaptive.cu)."""
  import torch
  return torch.cuda.get_device_properties(0).multi_processor_count * 2 * 128


def check_adaptive_subset(sol_oracle, sol_gpu, state0, picked, t_final,
                          log_capacity):
  n = state0.shape[1]
  state, time = state0.copy(), np.zeros((2, n))
  result = sol_gpu.flow_with_adaptive_step(t_final, state, time, 1e-3, 1e-3,
                                           step_log_capacity=log_capacity)
  expected_state = np.ascontiguousarray(state0[:, picked])
  expected_time = np.zeros((2, len(picked)))
  expected = sol_oracle.flow_with_adaptive_step(
      t_final, expected_state, expected_time, 1e-3, 1e-3,
      step_log_capacity=log_capacity)
  assert np.all(result['status'] == 0)
  assert np.all(time[0] == t_final)
  assert list(result['status'][picked]) == list(expected['status'])
  assert list(result['accepted'][picked]) == list(expected['accepted'])
  assert list(result['rejected'][picked]) == list(expected['rejected'])
  assert list(result['evaluations'][picked]) == list(expected['evaluations'])
  assert_bitwise_equal(result['step_log'][:, picked], expected['step_log'],
                       'step log')
  assert_bitwise_equal(time[:, picked], expected_time, 'time')
  assert_bitwise_equal(state[:, picked], expected_state, 'state')
  return result


def test_adaptive_flow_at_config_3_size(sol_system, sol_oracle, sol_gpu):
  """1.2 10^5 eccentric probes: every lane of the persistent kernel takes
  three or four trajectories from the queue, in the order of the cost sort.
  200 of them against the oracle: status, accepted / rejected / evaluation
  counts, the accepted-step sequence and the final state, bit for bit."""
  n = 120000
  lanes = adaptive_lanes()
  assert n > 3 * lanes
  state0 = eccentric(sol_system, n, seed=1234)
  rng = np.random.Generator(np.random.MT19937(99))
  picked = np.unique(np.concatenate(
      [[0, 1, n - 2, n - 1], rng.integers(0, n, 196)]))
  result = check_adaptive_subset(sol_oracle, sol_gpu, state0, picked,
                                 t_final=6000.0, log_capacity=48)
  # The batch is heterogeneous (this is what the queue is for).
  accepted = result['accepted']
  assert accepted.min() >= 1 and accepted.max() > 2 * np.median(accepted)
  assert result['rejected'].sum() > 0


def test_adaptive_flow_with_one_more_trajectory_than_lanes(sol_system,
                                                           sol_oracle, sol_gpu):
  """n = lanes + 1: exactly one lane takes a second trajectory."""
  n = adaptive_lanes() + 1
  state0 = eccentric(sol_system, n, seed=4321)
  rng = np.random.Generator(np.random.MT19937(7))
  picked = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, 62)]))
  check_adaptive_subset(sol_oracle, sol_gpu, state0, picked, t_final=1500.0,
                        log_capacity=32)


def test_trappist_ensemble_at_config_4_size():
  """10^5 perturbed TRAPPIST-1 systems, Quinlan1999Order8A at 30 min, 72 steps
  (the SRKN14A startup and 65 multistep steps): 64 of the systems against the
  oracle bit for bit, once with the plain Solve and once with every appended
  state going to the device ContinuousTrajectory sink (the sink must not
  perturb the flow; its polynomials are checked in the next test)."""
  from test_gpu_nbody import ensemble_state
  system = SolarSystem.from_fixture('trappist_jd2457000')
  gpu = Ephemeris(system)
  method = _capi.QUINLAN_1999_ORDER_8A
  h = 1800.0
  oracle = OracleEphemeris(system, method=method, step=h)
  order, _ = gpu.internal_order()
  n_bodies = len(order)
  n_systems = 100000
  state0 = ensemble_state(system, order, n_systems, 1e-6, seed=2024)
  rng = np.random.Generator(np.random.MT19937(5))
  picked = np.unique(np.concatenate(
      [[0, 31, 32, n_systems - 1], rng.integers(0, n_systems, 60)]))
  t0 = system.epoch
  n_steps = 72
  t_final = t0 + n_steps * h + 1.0

  ensemble = NBodyEnsemble(gpu, method, h, state0, t0)
  assert ensemble.solve(t_final) == n_steps
  state, time = ensemble.state()
  ensemble.close()

  expected_state = np.ascontiguousarray(state0[:, :, picked])
  expected_time = np.array([t0, 0.0])
  assert oracle.nbody_ensemble_flow(method, h, expected_state, expected_time,
                                    t_final) == n_steps
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state[:, :, picked], expected_state, 'ensemble state')

  # With the device ContinuousTrajectory sink.
  n = n_bodies * n_systems
  trajectories = ContinuousTrajectories(n, h, 1e-3, max_segments=n_steps // 8 + 2)
  ensemble = NBodyEnsemble(gpu, method, h, state0, t0)
  assert ensemble.solve_to_trajectories(t_final, trajectories) == n_steps
  sunk_state, sunk_time = ensemble.state()
  assert_bitwise_equal(sunk_state[:, :, picked], expected_state,
                       'ensemble state with the sink')
  assert trajectories.segment_count() == n_steps // 8
  assert trajectories.status() == 0
  ensemble.close()
  trajectories.close()


def test_trappist_sink_at_config_4_size_matches_oracle_trajectories():
  """The polynomials the device sink fits at full size (8 10^5 trajectories in
  lockstep) against the oracle's ContinuousTrajectory fed with the states of a
  second, stepwise, device ensemble restricted to the watched systems (the
  systems are independent and the flow is bit-exact, previous test)."""
  from test_gpu_nbody import ensemble_state
  from test_gpu_continuous_trajectory import OracleTrajectory
  system = SolarSystem.from_fixture('trappist_jd2457000')
  gpu = Ephemeris(system)
  method = _capi.QUINLAN_1999_ORDER_8A
  h = 1800.0
  order, _ = gpu.internal_order()
  n_bodies = len(order)
  n_systems = 100000
  state0 = ensemble_state(system, order, n_systems, 1e-6, seed=2024)
  picked = np.array([0, 4711, 65536, n_systems - 1])
  t0 = system.epoch
  n_steps = 8 * 5 + 3
  n = n_bodies * n_systems
  trajectories = ContinuousTrajectories(n, h, 1e-3, max_segments=8)
  ensemble = NBodyEnsemble(gpu, method, h, state0, t0)
  assert ensemble.solve_to_trajectories(t0 + n_steps * h + 1.0,
                                        trajectories) == n_steps
  assert trajectories.segment_count() == 5
  ensemble.close()

  small = NBodyEnsemble(gpu, method, h,
                        np.ascontiguousarray(state0[:, :, picked]), t0)
  oracles = {(b, k): OracleTrajectory(h, 1e-3)
             for b in range(n_bodies) for k in range(len(picked))}
  for step in range(n_steps + 1):
    if step > 0:
      assert small.solve(t0 + step * h + 1.0) == 1
    state, time = small.state()
    for (b, k), trajectory in oracles.items():
      assert trajectory.append(time[0], state[0:3, b, k],
                               state[6:9, b, k]) == 0
  for (b, k), trajectory in oracles.items():
    actual = trajectories.segments(b * n_systems + int(picked[k]))
    expected = trajectory.segments()
    assert_bitwise_equal(actual['t_max'], expected['t_max'])
    assert_bitwise_equal(actual['origin'], expected['origin'])
    assert list(actual['degree']) == list(expected['degree'])
    valid = np.arange(18)[None, :, None] <= actual['degree'][:, None, None]
    assert_bitwise_equal(np.where(valid, actual['coefficients'], 0.0),
                         np.where(valid, expected['coefficients'], 0.0),
                         'body %d system %d' % (b, picked[k]))
  trajectories.close()


def test_all_pairs_at_config_5_size():
  """N = 2^20 point masses: the accelerations of a block of 8 192 targets at an
  odd offset (what one rank of a sharded system evaluates: 8.6 10^9
  interactions), 64 of them against the O(N) oracle sum in the kernel's
  summation order, bit for bit; and against a plain sequential sum in
  extended precision to 1e-13 relative (the reference's own order, scattered
  third-law updates over all pairs, cannot be followed for a subset)."""
  import oracle_lib
  from test_gpu_nbody import random_cluster
  n = 2**20
  mu, q, v = random_cluster(n, seed=2020)
  begin, end = 123457, 123457 + 8192
  system = NBodyAllPairs(_capi.QUINLAN_TREMAINE_1990_ORDER_12, 3600.0, mu, q,
                         v, 0.0, i_begin=begin, i_end=end)
  a = system.accelerations()
  assert system.last_force_ms() > 0.0
  rng = np.random.Generator(np.random.MT19937(3))
  picked = np.unique(np.concatenate([[0, 31, 32, 8191],
                                     rng.integers(0, 8192, 60)]))
  expected = oracle_lib.allpairs_target_accelerations(mu, q, begin + picked)
  assert_bitwise_equal(a[:, picked], expected, 'accelerations at N = 2^20')
  # Independent of the summation order to 1e-13.
  for k in picked[:8]:
    i = begin + int(k)
    d = (q - q[:, i:i + 1]).astype(np.longdouble)
    d2 = (d * d).sum(axis=0)
    d2[i] = 1.0
    w = mu.astype(np.longdouble) / (d2 * np.sqrt(d2))
    w[i] = 0.0
    reference = (d * w).sum(axis=1).astype(np.float64)
    assert (np.linalg.norm(a[:, k] - reference) /
            np.linalg.norm(reference)) < 1e-13
  system.close()


@pytest.mark.parametrize('n', [1, 2, 63, 64, 65, 700, 4097, 20000])
def test_all_pairs_matches_the_restated_order_bitwise(n):
  """Every target of small systems, ragged sizes included (n below one stage,
  sub-ranges that are empty, a last stage that is partial)."""
  import oracle_lib
  from test_gpu_nbody import random_cluster
  mu, q, v = random_cluster(n, seed=n)
  system = NBodyAllPairs(_capi.BLANES_MOAN_2002_SRKN_14A, 60.0, mu, q, v, 0.0)
  a = system.accelerations()
  expected = oracle_lib.allpairs_target_accelerations(mu, q, np.arange(n))
  assert_bitwise_equal(a, expected, 'n = %d' % n)
  system.close()


def test_all_pairs_coincident_bodies_take_the_ieee_path():
  """Two bodies at the same place: the distance is outside the range of the
  branch-free square root and division, the targets involved are redone with
  the IEEE operators and produce the reference's NaN; every other target is
  unaffected."""
  import oracle_lib
  from test_gpu_nbody import random_cluster
  n = 3000
  mu, q, v = random_cluster(n, seed=77)
  q[:, 1234] = q[:, 17]
  system = NBodyAllPairs(_capi.BLANES_MOAN_2002_SRKN_14A, 60.0, mu, q, v, 0.0)
  a = system.accelerations()
  expected = oracle_lib.allpairs_target_accelerations(mu, q, np.arange(n))
  assert np.all(np.isnan(expected[:, [17, 1234]]))
  assert np.isfinite(np.delete(expected, [17, 1234], axis=1)).all()
  assert_bitwise_equal(a, expected, 'coincident bodies')
  system.close()


def test_all_pairs_flow_at_two_to_the_seventeen():
  """The Quinlan-Tremaine order-12 flow at N = 2^17: the startup (SRKN14A at
  step / 16) cut short to its first steps by t_final, i.e. 1 + 2 x 14 force
  evaluations of 1.7 10^10 interactions each; a sharded block owner gives the
  same bits as the whole system on its block."""
  from test_gpu_nbody import random_cluster
  n = 2**17
  mu, q, v = random_cluster(n, seed=5)
  method = _capi.QUINLAN_TREMAINE_1990_ORDER_12
  step = 3600.0
  t_final = 2 * step / 16 + 1.0  # Two steps of the startup integrator.
  whole = NBodyAllPairs(method, step, mu, q, v, 0.0)
  steps, forces = whole.solve(t_final)
  assert steps == 0 and forces == 1 + 2 * 14
  qw, vw, tw = whole.state()
  assert tw[0] == 2 * step / 16
  whole.close()
  # The second of eight blocks, positions of the others provided by a callback
  # that replays the whole system's exchange (one GPU stands in for eight: the
  # block owner only integrates its own targets, so the other rows are frozen
  # at the initial positions -- the first force evaluation, which sees exactly
  # those, must agree bit for bit).
  begin, end = n // 8, 2 * n // 8
  part = NBodyAllPairs(method, step, mu, q, v, 0.0, i_begin=begin, i_end=end)
  a_part = part.accelerations()
  fresh = NBodyAllPairs(method, step, mu, q, v, 0.0)
  a_whole = fresh.accelerations()
  assert_bitwise_equal(a_part, a_whole[:, begin:end], 'block of a sharded system')
  part.close()
  fresh.close()
  # Energy-like sanity at size: the centre-of-mass acceleration vanishes.
  total = (a_whole * mu).sum(axis=1)
  scale = np.abs(a_whole * mu).sum(axis=1)
  assert np.all(np.abs(total) < 1e-12 * scale)


def test_all_pairs_multi_rank_bit_equality():
  """tools/multigpu_allpairs_check.py under torchrun on the visible GPUs: the
  i-block sharded flows (NCCL all-gather issued by the library, and the
  callback form over torch.distributed) reproduce the slices of the unsharded
  flow bit for bit.  Needs at least two GPUs."""
  import os
  import socket
  import subprocess
  import sys
  import torch
  count = torch.cuda.device_count()
  if count < 2:
    pytest.skip('needs at least 2 GPUs (runs under gpurun --gpus 2/4/8 and in '
                'the driver\'s multi-GPU tier)')
  world = 2 if count < 4 else 4 if count < 8 else 8
  repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
  result = subprocess.run(
      [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
       '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
       '--master-port', str(port),
       os.path.join(repo, 'tools', 'multigpu_allpairs_check.py')],
      capture_output=True, text=True, timeout=900)
  assert result.returncode == 0, result.stdout[-4000:] + result.stderr[-4000:]
  assert 'multi-GPU all-pairs check OK on %d ranks' % world in result.stdout
