"""GPU parity of the adaptive massless flow (Ephemeris::FlowWithAdaptiveStep,
DormandElMikkawyPrince1986RKN434FM) against the CPU oracle: identical
accepted-step sequence, rejection counts, evaluation counts and final state,
bit for bit (both sides use a correctly-rounded fourth root in the step-size
control; BASELINE.json asks for an identical accepted-step sequence)."""
import numpy as np
import pytest

from principia_b200 import _capi, workloads

from conftest import assert_bitwise_equal

pytestmark = pytest.mark.gpu


def probes(system, n, seed=42):
  earth = system.bodies[system.index('Earth')]
  return workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                    seed)


def test_adaptive_flow_matches_oracle(sol_system, sol_oracle, sol_gpu):
  n = 333
  state = probes(sol_system, n)
  expected_state = state.copy()
  time = np.zeros((2, n))
  expected_time = time.copy()
  t_final = 20000.0
  result = sol_gpu.flow_with_adaptive_step(t_final, state, time, 1e-3, 1e-3,
                                           step_log_capacity=64)
  expected = sol_oracle.flow_with_adaptive_step(
      t_final, expected_state, expected_time, 1e-3, 1e-3,
      step_log_capacity=64)
  assert np.all(expected['status'] == 0)
  assert list(result['status']) == list(expected['status'])
  assert list(result['accepted']) == list(expected['accepted'])
  assert list(result['rejected']) == list(expected['rejected'])
  assert list(result['evaluations']) == list(expected['evaluations'])
  assert expected['accepted'].min() >= 3 and expected['rejected'].min() >= 1
  assert_bitwise_equal(result['step_log'], expected['step_log'], 'step log')
  assert_bitwise_equal(time, expected_time, 'time')
  assert np.all(time[0] == t_final)
  assert_bitwise_equal(state, expected_state, 'state')
  # FSAL: 4 evaluations per attempt until the first acceptance, 3 afterwards.
  attempts = result['accepted'] + result['rejected']
  assert np.all(result['evaluations'] <= 3 * attempts + attempts.clip(max=50))


def test_adaptive_flow_statuses(sol_system, sol_oracle, sol_gpu):
  n = 40
  # ReachedMaximalStepCount.
  state, time = probes(sol_system, n, 7), np.zeros((2, n))
  expected_state, expected_time = state.copy(), time.copy()
  result = sol_gpu.flow_with_adaptive_step(50000.0, state, time, 1e-3, 1e-3,
                                           max_steps=5)
  expected = sol_oracle.flow_with_adaptive_step(
      50000.0, expected_state, expected_time, 1e-3, 1e-3, max_steps=5)
  assert list(result['status']) == list(expected['status'])
  assert np.all(result['status'] == _capi.PB_RESOURCE_EXHAUSTED)
  assert np.all(result['accepted'] == 5)
  assert_bitwise_equal(state, expected_state, 'state after 5 steps')
  assert_bitwise_equal(time, expected_time)
  # Beyond the ephemeris: integrates to t_max and reports DeadlineExceeded.
  state, time = probes(sol_system, n, 8), np.full((2, n), 0.0)
  time[0] = 170000.0
  expected_state, expected_time = state.copy(), time.copy()
  result = sol_gpu.flow_with_adaptive_step(180000.0, state, time, 1e-3, 1e-3)
  expected = sol_oracle.flow_with_adaptive_step(
      180000.0, expected_state, expected_time, 1e-3, 1e-3)
  assert np.all(result['status'] == _capi.PB_DEADLINE_EXCEEDED)
  assert list(result['status']) == list(expected['status'])
  assert np.all(time[0] == 172800.0)
  assert_bitwise_equal(state, expected_state, 'deadline state')
  # Already there: nothing to do.
  before = state.copy()
  result = sol_gpu.flow_with_adaptive_step(172800.0, state, time, 1e-3, 1e-3)
  assert np.all(result['status'] == 0) and np.all(result['accepted'] == 0)
  assert_bitwise_equal(state, before)
  # A probe falling into the Earth: the collision is swallowed
  # (ephemeris_body.hpp:1681-1683), the flow carries on.
  earth = sol_system.bodies[sol_system.index('Earth')]
  state = np.zeros((12, 2))
  state[0:3] = np.array(earth['q'])[:, None] + np.array([[6.4e6, 7e6],
                                                         [0.0, 0.0],
                                                         [0.0, 0.0]])
  state[6:9] = np.array(earth['v'])[:, None]
  time = np.zeros((2, 2))
  expected_state, expected_time = state.copy(), time.copy()
  result = sol_gpu.flow_with_adaptive_step(300.0, state, time, 1e-3, 1e-3,
                                           max_steps=2000)
  expected = sol_oracle.flow_with_adaptive_step(
      300.0, expected_state, expected_time, 1e-3, 1e-3, max_steps=2000)
  assert list(result['status']) == list(expected['status'])
  assert list(result['accepted']) == list(expected['accepted'])
  assert_bitwise_equal(state, expected_state, 'collision state')
  # No probes.
  sol_gpu.flow_with_adaptive_step(100.0, np.zeros((12, 0)), np.zeros((2, 0)),
                                  1e-3, 1e-3)


def test_inertially_fixed_burns_match_oracle(sol_system, sol_oracle, sol_gpu):
  """FlowWithAdaptiveStep with an IntrinsicAcceleration (ephemeris_body.hpp:
  415-444): a tabulated inertially fixed burn per probe
  (ksp_plugin/manœuvre_body.hpp:395-406), starting and ending inside the flow,
  covering it, or outside it.  Same accepted steps and states as the oracle."""
  n = 40
  earth = sol_system.bodies[sol_system.index('Earth')]
  state = workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                     17)
  rng = np.random.default_rng(4)
  burns = np.zeros((8, n))
  direction = rng.normal(size=(3, n))
  burns[0:3] = direction / np.linalg.norm(direction, axis=0)
  burns[3] = rng.uniform(1e3, 5e5, n)        # Thrust, N.
  burns[4] = rng.uniform(5e3, 5e4, n)        # Initial mass, kg.
  burns[5] = burns[3] / (9.80665 * 320.0)    # Mass flow at 320 s.
  burns[6] = rng.uniform(-500.0, 4000.0, n)  # Initial time.
  burns[7] = burns[6] + rng.uniform(10.0, 600.0, n)
  burns[6, 0], burns[7, 0] = -1e9, 1e9       # Always on ...
  burns[5, 0] = 0.0                          # ... without losing mass.
  burns[6, 1], burns[7, 1] = 1e8, 2e8        # Never on.
  t_final = 6000.0
  capacity = 3000
  expected_state, expected_time = state.copy(), np.zeros((2, n))
  time = np.zeros((2, n))
  result = sol_gpu.flow_with_adaptive_step(t_final, state, time, 1e-3, 1e-3,
                                           step_log_capacity=capacity,
                                           burns=burns)
  expected = sol_oracle.flow_with_adaptive_step(
      t_final, expected_state, expected_time, 1e-3, 1e-3,
      step_log_capacity=capacity, burns=burns)
  assert np.all(result['status'] == expected['status'])
  assert list(result['accepted']) == list(expected['accepted'])
  assert list(result['rejected']) == list(expected['rejected'])
  assert list(result['evaluations']) == list(expected['evaluations'])
  assert result['accepted'].max() < capacity
  assert_bitwise_equal(result['step_log'], expected['step_log'], 'step log')
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'state')
  # The burns change the trajectories.
  free_state, free_time = expected_state.copy(), np.zeros((2, n))
  free_state[:] = workloads.eccentric_probes(n, earth['q'], earth['v'],
                                             earth['mu'], 17)
  sol_oracle.flow_with_adaptive_step(t_final, free_state, free_time, 1e-3,
                                     1e-3)
  moved = np.linalg.norm(free_state[0:3] - state[0:3], axis=0)
  assert moved[1] == 0.0 and moved[0] > 1e3


@pytest.mark.parametrize('integrator,expected_size', [
    (_capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, 379),
    (_capi.QUINLAN_1999_ORDER_8A, 406),
])
def test_reference_earth_probe(sol_system, integrator, expected_size):
  """EphemerisTest.EarthProbe of the reference (physics/ephemeris_test.cpp:
  374-497; tests/earth_probe.py) on the GPU: a flow made of rounding errors.
  The trajectory has one of the sizes the reference pins (379: its "Ubuntu AVX"
  value) and every accepted time and the final state equal the oracle's."""
  import earth_probe
  from principia_b200.ephemeris import Ephemeris
  from oracle_lib import OracleEphemeris
  earth_only, period, v1 = earth_probe.set_up(sol_system)
  t0 = earth_probe.T0
  oracle = OracleEphemeris(earth_only, method=integrator, step=period / 100,
                           fitting_tolerance=5e-3,
                           initial_positions=np.zeros((1, 3)),
                           initial_velocities=np.array([[v1, 0.0, 0.0]]),
                           initial_time=t0)
  assert oracle.prolong(t0 + period) == 0
  gpu = Ephemeris(earth_only)
  gpu.append_segments(0, oracle.segments(0))
  state, time, burns = earth_probe.probe(sol_system, v1)
  expected_state, expected_time = state.copy(), time.copy()
  capacity = 512
  result = gpu.flow_with_adaptive_step(t0 + period, state, time, 1e-9, 2.6e-15,
                                       max_steps=1000000,
                                       step_log_capacity=capacity, burns=burns)
  expected = oracle.flow_with_adaptive_step(
      t0 + period, expected_state, expected_time, 1e-9, 2.6e-15,
      max_steps=1000000, step_log_capacity=capacity, burns=burns)
  size = int(result['accepted'][0]) + 1
  assert size == int(expected['accepted'][0]) + 1 == expected_size
  assert size in earth_probe.REFERENCE_TRAJECTORY_SIZES
  assert result['status'][0] == expected['status'][0] == 0
  assert list(result['rejected']) == list(expected['rejected'])
  assert_bitwise_equal(result['step_log'], expected['step_log'], 'step log')
  assert_bitwise_equal(state, expected_state, 'state')
  low, high = earth_probe.REFERENCE_PROBE_ULPS
  assert low <= earth_probe.ulp_distance(state[0, 0],
                                         period * state[6, 0]) <= high
  assert state[1, 0] == earth_probe.DISTANCE
  # Without prolonging, flowing to infinity stops at t_max: DeadlineExceeded.
  t_max = gpu.t_max()
  result = gpu.flow_with_adaptive_step(np.inf, state, time, 1e-9, 2.6e-15,
                                       max_steps=1000000, burns=burns)
  assert result['status'][0] == 4 and time[0, 0] == t_max


def test_generalized_flow_fine1987rkng34_matches_oracle(sol_system, sol_oracle,
                                                        sol_gpu):
  """The generalized FlowWithAdaptiveStep (ephemeris_body.hpp:446-478) with
  Fine1987RKNG34, the plugin's burn integrator, with and without (time-only)
  intrinsic accelerations: the oracle's accepted steps and states."""
  n = 48
  earth = sol_system.bodies[sol_system.index('Earth')]
  state = workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                     19)
  rng = np.random.default_rng(6)
  burns = np.zeros((8, n))
  burns[2] = 1.0
  burns[3] = rng.uniform(1e3, 5e5, n)
  burns[4] = 2e4
  burns[5] = burns[3] / (9.80665 * 300.0)
  burns[6] = rng.uniform(0.0, 2000.0, n)
  burns[7] = burns[6] + 300.0
  method = _capi.FINE_1987_RKNG_34
  for table in (None, burns):
    gpu_state, gpu_time = state.copy(), np.zeros((2, n))
    expected_state, expected_time = state.copy(), np.zeros((2, n))
    capacity = 3000
    result = sol_gpu.flow_with_adaptive_step(
        5000.0, gpu_state, gpu_time, 1e-3, 1e-3, step_log_capacity=capacity,
        burns=table, method=method)
    expected = sol_oracle.flow_with_adaptive_step(
        5000.0, expected_state, expected_time, 1e-3, 1e-3,
        step_log_capacity=capacity, burns=table, method=method)
    assert list(result['accepted']) == list(expected['accepted'])
    assert list(result['rejected']) == list(expected['rejected'])
    # 5 stages, no FSAL: 5 evaluations per attempt.
    assert list(result['evaluations']) == list(
        5 * (expected['accepted'] + expected['rejected']))
    assert result['accepted'].max() < capacity
    assert_bitwise_equal(result['step_log'], expected['step_log'], 'step log')
    assert_bitwise_equal(gpu_state, expected_state, 'state')


@pytest.mark.parametrize('integrator', [
    _capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, _capi.QUINLAN_1999_ORDER_8A])
def test_reference_elephant_at_the_pole(sol_system, integrator):
  """physics/ephemeris_test.cpp:682-759 on the GPU: 5 points, the abscissa of
  the elephant -9.8(1)e-19 m because the cosine of 90 degrees is not zero."""
  import earth_probe
  from principia_b200.ephemeris import Ephemeris
  from oracle_lib import OracleEphemeris
  t0 = earth_probe.T0
  polar_radius = 6.3568e6
  earth_only = sol_system.copy()
  for name in list(earth_only.names()):
    if name != 'Earth':
      earth_only.remove_massive_body(name)
  earth_only.limit_oblateness_to_degree('Earth', 2)
  earth_only.limit_oblateness_to_zonal('Earth')
  oracle = OracleEphemeris(earth_only, method=integrator, step=1.0 / 100,
                           fitting_tolerance=5e-3,
                           initial_positions=np.zeros((1, 3)),
                           initial_velocities=np.zeros((1, 3)),
                           initial_time=t0)
  assert oracle.prolong(t0 + 1.0) == 0
  gpu = Ephemeris(earth_only)
  gpu.append_segments(0, oracle.segments(0))
  state = np.zeros((12, 1))
  state[2, 0] = polar_radius
  time = np.zeros((2, 1))
  time[0, 0] = t0
  expected_state, expected_time = state.copy(), time.copy()
  result = gpu.flow_with_adaptive_step(
      t0 + 1.0, state, time, 1e-9, 2.6e-15, max_steps=1000000,
      method=_capi.FINE_1987_RKNG_34)
  oracle.flow_with_adaptive_step(
      t0 + 1.0, expected_state, expected_time, 1e-9, 2.6e-15,
      max_steps=1000000, method=_capi.FINE_1987_RKNG_34)
  assert int(result['accepted'][0]) + 1 == 5
  assert_bitwise_equal(state, expected_state, 'elephant')
  assert -9.9e-19 <= state[0, 0] <= -9.7e-19
  assert 5.9e-35 <= state[1, 0] <= 6.1e-35
  assert abs(state[2, 0] - polar_radius) / polar_radius < 8e-7
  acceleration, _ = gpu.compute_gravitational_acceleration_on_massless_bodies(
      time[0, 0], np.ascontiguousarray(state[0:3]))
  assert -2.1e-18 <= acceleration[0, 0] <= -1.9e-18
  assert abs(acceleration[2, 0] + 9.832) / 9.832 < 6.7e-6


def test_frenet_burns_match_oracle(sol_system, sol_oracle, sol_gpu):
  """The generalized FlowWithAdaptiveStep (ephemeris_body.hpp:446-478) with a
  velocity-dependent intrinsic acceleration: burns given in the Frenet frame
  of the body-centred non-rotating frame of the Earth (or of the Moon)
  (Manœuvre::FrenetIntrinsicAcceleration, ksp_plugin/manœuvre_body.hpp:
  310-321), integrated with Fine1987RKNG34, whose velocity stages a' now feed
  the right-hand side.  Same accepted steps and states as the oracle's
  generalized integrator (itself pinned to the reference's Legendre test), and
  a prograde burn does what a prograde burn does."""
  n = 48
  earth = sol_system.bodies[sol_system.index('Earth')]
  moon_index = sol_system.index('Moon')
  earth_index = sol_system.index('Earth')
  state = workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                     23)
  rng = np.random.default_rng(6)
  burns = np.zeros((8, n))
  direction = rng.normal(size=(3, n))
  burns[0:3] = direction / np.linalg.norm(direction, axis=0)
  burns[0:3, 0] = [1.0, 0.0, 0.0]            # Prograde.
  burns[0:3, 1] = [-1.0, 0.0, 0.0]           # Retrograde.
  burns[0:3, 2] = [0.0, 0.0, 1.0]            # Along the orbit normal.
  state[:, 2] = state[:, 0]                  # (On the orbit of probe 0.)
  burns[4] = rng.uniform(5e3, 5e4, n)        # Initial mass, kg.
  # 0.1 to 3 m/s^2 at ignition; at most 400 s at 320 s of specific impulse
  # burn 38 % of the mass (a burn that exhausts the mass is singular and the
  # step control then crawls for ever, here as in the reference).
  burns[3] = burns[4] * rng.uniform(0.1, 3.0, n)
  burns[5] = burns[3] / (9.80665 * 320.0)    # Mass flow.
  burns[6] = rng.uniform(-200.0, 3000.0, n)  # Initial time.
  burns[7] = burns[6] + rng.uniform(10.0, 400.0, n)
  burns[3:8, 0:3] = np.array([2e4, 2e4, 2e3, 100.0, 400.0])[:, None]
  burns[5, 0:3] = burns[3, 0:3] / (9.80665 * 320.0)
  centres = np.full(n, earth_index, dtype=np.int32)
  centres[5::7] = moon_index
  centres[3::5] = -1  # Inertially fixed burns in the same batch.
  method = _capi.FINE_1987_RKNG_34
  t_final = 6000.0
  time = np.zeros((2, n))
  expected_state, expected_time = state.copy(), time.copy()
  initial = state.copy()
  result = sol_gpu.flow_with_adaptive_step(
      t_final, state, time, 1e-3, 1e-3, step_log_capacity=96, burns=burns,
      method=method, frenet_centres=centres, max_steps=50000)
  expected = sol_oracle.flow_with_adaptive_step(
      t_final, expected_state, expected_time, 1e-3, 1e-3,
      step_log_capacity=96, burns=burns, method=method,
      frenet_centres=centres, max_steps=50000)
  assert np.all(expected['status'] == 0)
  assert list(result['status']) == list(expected['status'])
  assert list(result['accepted']) == list(expected['accepted'])
  assert list(result['rejected']) == list(expected['rejected'])
  assert list(result['evaluations']) == list(expected['evaluations'])
  assert_bitwise_equal(result['step_log'], expected['step_log'], 'step log')
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'state')
  # The Frenet burns differ from the same burns held inertially fixed.
  fixed_state, fixed_time = initial.copy(), np.zeros((2, n))
  sol_oracle.flow_with_adaptive_step(t_final, fixed_state, fixed_time, 1e-3,
                                     1e-3, burns=burns, method=method)
  frenet = centres >= 0
  # (A burn that ends before the flow starts, or that falls between the stage
  # times of a probe far from the Earth, is never seen.)
  fired = frenet & (burns[7] > 0.0) & (result['accepted'] > 50)
  assert fired.sum() > 12
  assert np.all(np.linalg.norm((state - fixed_state)[0:3, fired], axis=0) > 1.0)
  assert_bitwise_equal(state[:, ~frenet], fixed_state[:, ~frenet])
  # Physics: against the unpowered flow, the prograde burn (300 s at 1 to
  # 1.3 m/s^2) raises the orbital energy relative to the Earth, the retrograde
  # one lowers it, the normal one changes it much less.
  coast_state, coast_time = initial.copy(), np.zeros((2, n))
  sol_oracle.flow_with_adaptive_step(t_final, coast_state, coast_time, 1e-3,
                                     1e-3, method=method)
  q_earth = sol_oracle.evaluate_all_positions(t_final)[earth_index]
  v_earth = sol_oracle.evaluate_all_velocities(t_final)[earth_index]

  def energy(s):
    q = s[0:3] - q_earth[:, None]
    v = s[6:9] - v_earth[:, None]
    return 0.5 * (v * v).sum(axis=0) - earth['mu'] / np.linalg.norm(q, axis=0)
  gain = energy(state) - energy(coast_state)
  assert gain[0] > 0 and gain[1] < 0
  assert abs(gain[2]) < 0.05 * min(gain[0], -gain[1])
