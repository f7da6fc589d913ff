"""GPU parity of the fixed-step flow of massless bodies with the symmetric
linear multistep integrators (the plugin's history
integrator is Quinlan1999Order8A at 10 s, ksp_plugin/integrators.cpp:66-72):
bit for bit against the oracle, through the C ABI."""
import numpy as np
import pytest

from principia_b200 import _capi
from principia_b200 import workloads

from conftest import assert_bitwise_equal

pytestmark = pytest.mark.gpu


def leo(system, n, seed=42):
  earth = system.bodies[system.index('Earth')]
  return workloads.leo_probes(n, earth['q'], earth['v'], earth['mu'], seed)


@pytest.mark.parametrize('method,order', [
    (_capi.QUINLAN_1999_ORDER_8A, 8),
    (_capi.QUINLAN_TREMAINE_1990_ORDER_12, 12),
])
def test_multistep_flow_matches_oracle(sol_system, sol_oracle, sol_gpu, method,
                                       order):
  n = 700  # Ragged.
  state = leo(sol_system, n)
  expected_state = state.copy()
  time = np.zeros(2)
  expected_time = time.copy()
  # 10 s steps; the startup (order - 1 steps of 16 SRKN14A substeps) and then
  # more than one chunk of the stage table (> 256 steps); t_final is not a
  # multiple of the step.
  t_final = 3005.0
  result = sol_gpu.flow_with_fixed_step(method, 10.0, t_final, state, time,
                                        sample_every=7, max_samples=50)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, t_final,
                                             expected_state, expected_time,
                                             sample_every=7, max_samples=50)
  assert result['n_steps'] == expected['n_steps'] == 300
  assert result['status'] == expected['status'] == 0
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'state')
  assert len(result['samples']) == len(expected['samples']) == 42
  assert_bitwise_equal(result['sample_times'], expected['sample_times'])
  assert_bitwise_equal(result['samples'], expected['samples'], 'samples')


def test_instance_keeps_its_history(sol_system, sol_oracle, sol_gpu):
  """FlowWithFixedStep called repeatedly on one instance (the plugin's
  pattern) equals one Solve to the final time: the calls stop in the middle of
  the startup, right after it, and in the multistep regime."""
  method = _capi.QUINLAN_1999_ORDER_8A
  n = 96
  state0 = leo(sol_system, n, seed=5)
  instance = sol_gpu.new_instance(method, 10.0, state0, np.zeros(2))
  appended = 0
  for t in (4.0, 33.0, 70.0, 75.0, 400.0, 1000.0):
    appended += instance.flow(t)['n_steps']
  state, time = instance.state()
  expected_state = state0.copy()
  expected_time = np.zeros(2)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, 1000.0,
                                             expected_state, expected_time)
  assert appended == expected['n_steps'] == 100
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'state')
  # An SRKN instance goes through the same entry points.
  method = _capi.BLANES_MOAN_2002_SRKN_14A
  instance = sol_gpu.new_instance(method, 10.0, state0, np.zeros(2))
  instance.flow(300.0)
  result = instance.flow(600.0, sample_every=10, max_samples=3)
  state, time = instance.state()
  expected_state = state0.copy()
  expected_time = np.zeros(2)
  sol_oracle.flow_with_fixed_step(method, 10.0, 600.0, expected_state,
                                  expected_time)
  assert result['n_steps'] == 30 and len(result['samples']) == 3
  assert_bitwise_equal(state, expected_state, 'SRKN instance')
  assert_bitwise_equal(result['samples'][-1][:3], expected_state[:3])


def test_multistep_collision_status(sol_system, sol_oracle, sol_gpu):
  """Collisions are reported by the multistep steps only: the reference ignores
  the status of the startup and of FillStepFromState (starter_body.hpp:68-73,
  symmetric_linear_multistep_integrator_body.hpp:258-261).  One probe starts
  deep inside the Earth and is flung out during the startup; another one skims
  the surface (6300 km, inside 0.99 times the minimal radius) all along."""
  method = _capi.QUINLAN_1999_ORDER_8A
  earth = sol_system.bodies[sol_system.index('Earth')]
  for radius, expected_status in ((1.0e6, None), (6.30e6, _capi.PB_OUT_OF_RANGE)):
    state = leo(sol_system, 8, seed=9)
    r = state[0:3, 3] - np.asarray(earth['q'])
    state[0:3, 3] = np.asarray(earth['q']) + r * (radius / np.linalg.norm(r))
    expected_state = state.copy()
    time, expected_time = np.zeros(2), np.zeros(2)
    result = sol_gpu.flow_with_fixed_step(method, 10.0, 200.0, state, time)
    expected = sol_oracle.flow_with_fixed_step(method, 10.0, 200.0,
                                               expected_state, expected_time)
    assert result['status'] == expected['status']
    if expected_status is not None:
      assert result['status'] == expected_status
    assert result['n_steps'] == expected['n_steps'] == 20
    assert_bitwise_equal(state, expected_state, 'radius %g' % radius)


def test_reference_collision_detection_apple(sol_system):
  """physics/ephemeris_test.cpp:763-811 (CollisionDetection): an apple 10 m
  above the Earth (alone, at rest at the origin), Quinlan1999Order8A at 1 ms.
  OK after 1 s, OutOfRange when the same instance is flowed to 354.8 s; the
  product agrees with the oracle bit for bit at both times."""
  from principia_b200.ephemeris import Ephemeris
  from oracle_lib import OracleEphemeris
  system = sol_system.copy()
  for name in list(system.names()):
    if name != 'Earth':
      system.remove_massive_body(name)
  zeros = np.zeros((1, 3))
  oracle = OracleEphemeris(system, step=0.01, fitting_tolerance=5e-3,
                           initial_positions=zeros, initial_velocities=zeros,
                           initial_time=0.0)
  assert oracle.prolong(355.0) == 0
  gpu = Ephemeris(system)
  gpu.append_segments(0, oracle.segments(0))
  earth = system.bodies[0]
  state0 = np.zeros((12, 1))
  state0[2, 0] = earth['mean_radius'] + 10.0
  method = _capi.QUINLAN_1999_ORDER_8A
  instance = gpu.new_instance(method, 1e-3, state0, np.zeros(2))
  expected_state, expected_time = state0.copy(), np.zeros(2)
  result = instance.flow(1.0)
  expected = oracle.flow_with_fixed_step(method, 1e-3, 1.0, expected_state,
                                         expected_time)
  assert result['status'] == expected['status'] == 0
  assert result['n_steps'] == expected['n_steps']
  state, time = instance.state()
  assert_bitwise_equal(state, expected_state, 'after 1 s')
  # The oracle's instance lives for one call: compare with one Solve from the
  # start to 354.8 s.
  result = instance.flow(354.8)
  expected_state, expected_time = state0.copy(), np.zeros(2)
  expected = oracle.flow_with_fixed_step(method, 1e-3, 354.8, expected_state,
                                         expected_time)
  assert result['status'] == _capi.PB_OUT_OF_RANGE
  assert expected['status'] == _capi.PB_OUT_OF_RANGE
  state, time = instance.state()
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'after 354.8 s')


@pytest.mark.parametrize('integrator', [
    _capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, _capi.QUINLAN_1999_ORDER_8A])
def test_reference_earth_two_probes(sol_system, integrator):
  """EphemerisTest.EarthTwoProbes of the reference (physics/ephemeris_test.cpp:
  501-620): the Earth alone in uniform motion, two probes whose constant
  intrinsic accelerations cancel its pull, flowed with a fixed step of
  period / 1000 through NewInstance.  1001 points each; the abscissae lie
  25-70 and 2-13 ULPs from period * v, the ordinates are exactly constant; the
  product agrees with the oracle bit for bit."""
  import earth_probe
  from principia_b200.ephemeris import Ephemeris
  from oracle_lib import OracleEphemeris
  earth_only, period, v1 = earth_probe.set_up(sol_system)
  t0 = earth_probe.T0
  mu = sol_system.bodies[sol_system.index('Earth')]['mu']
  oracle = OracleEphemeris(earth_only, method=integrator, step=period / 100,
                           fitting_tolerance=5e-3,
                           initial_positions=np.zeros((1, 3)),
                           initial_velocities=np.array([[v1, 0.0, 0.0]]),
                           initial_time=t0)
  assert oracle.prolong(t0 + period) == 0
  gpu = Ephemeris(earth_only)
  gpu.append_segments(0, oracle.segments(0))
  state = np.zeros((12, 2))
  state[1] = [1e9, -3e9]
  state[6] = v1
  burns = np.zeros((8, 2))
  burns[1] = 1.0
  burns[3] = [mu / (1e9 * 1e9), -mu / (3e9 * 3e9)]
  burns[4] = 1.0
  burns[6], burns[7] = -1e30, 1e30
  expected_state = state.copy()
  expected_time = np.array([t0, 0.0])
  expected = oracle.flow_with_fixed_step(integrator, period / 1000,
                                         t0 + period, expected_state,
                                         expected_time, burns=burns)
  instance = gpu.new_instance(integrator, period / 1000, state,
                              np.array([t0, 0.0]))
  instance.set_intrinsic_accelerations(burns)
  # In two calls, like a plugin that advances its vessels frame by frame.
  steps = instance.flow(t0 + 0.3 * period)['n_steps']
  steps += instance.flow(t0 + period)['n_steps']
  assert steps == expected['n_steps'] == 1000
  final_state, final_time = instance.state()
  assert_bitwise_equal(final_time, expected_time, 'time')
  assert_bitwise_equal(final_state, expected_state, 'state')
  for k, (low, high) in enumerate(((25, 70), (2, 13))):
    distance = earth_probe.ulp_distance(final_state[0, k],
                                        1.00 * period * final_state[6, k])
    assert low <= distance <= high, (k, distance)
  assert list(final_state[1]) == [1e9, -3e9]
  # The stateless entry point takes the same table.
  state2, time2 = state.copy(), np.array([t0, 0.0])
  gpu.flow_with_fixed_step(integrator, period / 1000, t0 + period, state2,
                           time2, burns=burns)
  assert_bitwise_equal(state2, expected_state, 'stateless')


@pytest.mark.parametrize('method,stops', [
    (_capi.QUINLAN_1999_ORDER_8A, (33.0, 500.0)),   # Mid-startup, multistep.
    (_capi.QUINLAN_TREMAINE_1990_ORDER_12, (700.0,)),
    (_capi.BLANES_MOAN_2002_SRKN_14A, (300.0,)),
])
def test_instance_checkpoint_and_restore(sol_system, sol_gpu, method, stops):
  """WriteToMessage / ReadFromMessage of an instance: a restored instance (the
  original destroyed) continues bit for bit like an uninterrupted one, from the
  middle of the startup and from the multistep regime, burns included."""
  n = 50
  state0 = leo(sol_system, n, seed=31)
  rng = np.random.default_rng(2)
  burns = np.zeros((8, n))
  burns[0] = 1.0
  burns[3] = rng.uniform(1e3, 1e5, n)
  burns[4] = 2e4
  burns[5] = 1.0
  burns[6], burns[7] = 100.0, 900.0
  t_final = 1500.0
  uninterrupted = sol_gpu.new_instance(method, 10.0, state0, np.zeros(2))
  uninterrupted.set_intrinsic_accelerations(burns)
  uninterrupted.flow(t_final)
  expected_state, expected_time = uninterrupted.state()
  instance = sol_gpu.new_instance(method, 10.0, state0, np.zeros(2))
  instance.set_intrinsic_accelerations(burns)
  for stop in stops:
    instance.flow(stop)
    checkpoint = instance.checkpoint()
    del instance
    instance = type(uninterrupted).from_checkpoint(sol_gpu, checkpoint)
  instance.flow(t_final)
  state, time = instance.state()
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'state')
  from principia_b200.ephemeris import StatusError
  with pytest.raises(StatusError):
    type(uninterrupted).from_checkpoint(sol_gpu, checkpoint[:-1])


def test_multistep_flow_rejects_backward_steps(sol_system, sol_gpu):
  state = leo(sol_system, 4)
  with pytest.raises(Exception):
    sol_gpu.flow_with_fixed_step(_capi.QUINLAN_1999_ORDER_8A, -10.0, -100.0,
                                 state, np.zeros(2))


def test_full_size_multistep_flow_is_sliced_consistently(sol_system, sol_oracle,
                                                         sol_gpu):
  """At 40 000 probes the call is cut into probe slices and step sub-chunks on
  helper streams (principia_b200.cu); probes are independent, so any subset
  equals the oracle bit for bit, samples across slice boundaries included."""
  n = 40000
  method = _capi.QUINLAN_1999_ORDER_8A
  state0 = leo(sol_system, n, seed=9)
  state = state0.copy()
  time = np.zeros(2)
  t_final = 455.0  # Startup (7 steps) + 38 multistep steps: 4 sub-chunks.
  result = sol_gpu.flow_with_fixed_step(method, 10.0, t_final, state, time,
                                        sample_every=9, max_samples=8)
  assert result['n_steps'] == 45 and result['status'] == 0
  subset = np.r_[0:33, 9990:10030, 19999:20001, 29950:29990, n - 33:n]
  expected_state = np.ascontiguousarray(state0[:, subset])
  expected_time = np.zeros(2)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, t_final,
                                             expected_state, expected_time,
                                             sample_every=9, max_samples=8)
  assert expected['n_steps'] == 45
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state[:, subset], expected_state, 'subset')
  assert_bitwise_equal(result['sample_times'], expected['sample_times'])
  assert_bitwise_equal(result['samples'][:, :, subset], expected['samples'],
                       'subset samples')
