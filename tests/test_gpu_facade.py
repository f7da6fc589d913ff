"""GPU: the C++ facade (principia_b200/cpp/ephemeris.hpp) -- the reference's
Ephemeris surface over the C ABI -- against the oracle: Prolong (device
massive-body flow + host Newhall fit) produces the oracle's polynomials bit for
bit; FlowWithFixedStep / FlowWithAdaptiveStep append the oracle's points."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from principia_b200 import _capi, workloads

from conftest import assert_bitwise_equal
import oracle_lib
from oracle_lib import OracleEphemeris

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
_d, _i32, _i64, _vp = (ctypes.c_double, ctypes.c_int32, ctypes.c_int64,
                       ctypes.c_void_p)
_dp, _i32p, _i64p = _capi.c_double_p, _capi.c_int32_p, _capi.c_int64_p


@pytest.fixture(scope='module')
def shim():
  directory = os.path.join(HERE, 'cpp')
  subprocess.check_call(
      ['/usr/bin/g++', '-std=c++17', '-O2', '-ffp-contract=off', '-fPIC',
       '-shared', '-o', 'libfacadetest.so', 'facade_capi.cpp',
       '-L../../principia_b200', '-lprincipia_b200',
       '-Wl,-rpath,$ORIGIN/../../principia_b200'], cwd=directory)
  lib = ctypes.CDLL(os.path.join(directory, 'libfacadetest.so'))
  lib.ft_create.restype = _vp
  lib.ft_create.argtypes = [_i32, ctypes.POINTER(_capi.Body), _dp, _dp, _d, _d,
                            _d, _i32, _d]
  lib.ft_destroy.argtypes = [_vp]
  lib.ft_prolong.argtypes = [_vp, _d]
  lib.ft_t_max.restype = _d
  lib.ft_t_max.argtypes = [_vp]
  lib.ft_t_min.restype = _d
  lib.ft_t_min.argtypes = [_vp]
  lib.ft_empty.argtypes = [_vp]
  lib.ft_segment_count.argtypes = [_vp, _i32]
  lib.ft_get_segments.argtypes = [_vp, _i32, _dp, _dp, _i32p, _dp]
  lib.ft_flow_with_fixed_step.argtypes = [_vp, _i32, _dp, _dp, _d, _i32, _d,
                                          _d, _i32p, _dp, _dp]
  lib.ft_flow_with_fixed_step_twice.argtypes = [_vp, _i32, _dp, _dp, _d, _i32,
                                                _d, _d, _d, _i32p, _dp, _dp]
  lib.ft_flow_with_adaptive_step.argtypes = [_vp, _dp, _dp, _d, _d, _d, _d,
                                             _i64, _i64, _i64p, _dp, _dp]
  lib.ft_acceleration_on_massless_body.argtypes = [_vp, _dp, _d, _dp]
  lib.ft_jerk_and_jacobian.argtypes = [_vp, _dp, _dp, _i32, _d, _dp, _dp, _dp]
  return lib


def test_facade_prolong_and_flows(shim, sol_system):
  system = sol_system
  bodies, keep = system.c_bodies()
  q0 = np.ascontiguousarray(system.initial_positions())
  v0 = np.ascontiguousarray(system.initial_velocities())
  handle = shim.ft_create(len(system.bodies), bodies, oracle_lib.dptr(q0),
                          oracle_lib.dptr(v0), 0.0, 1e-3, 2.0**-24,
                          _capi.QUINLAN_TREMAINE_1990_ORDER_12, 600.0)
  assert handle
  oracle = OracleEphemeris(system)
  t = 0.4 * 86400.0
  assert shim.ft_prolong(handle, t) == 0
  assert oracle.prolong(t) == 0
  assert shim.ft_t_max(handle) == oracle.t_max() >= t
  for body in range(len(system.bodies)):
    expected = oracle.segments(body)
    count = shim.ft_segment_count(handle, body)
    assert count == len(expected['t_max'])
    t_max, origin = np.zeros(count), np.zeros(count)
    degree = np.zeros(count, dtype=np.int32)
    coefficients = np.zeros((count, 18, 3))
    shim.ft_get_segments(handle, body, oracle_lib.dptr(t_max),
                         oracle_lib.dptr(origin),
                         degree.ctypes.data_as(_i32p),
                         oracle_lib.dptr(coefficients))
    assert_bitwise_equal(t_max, expected['t_max'])
    assert_bitwise_equal(origin, expected['origin'])
    assert list(degree) == list(expected['degree']), system.names()[body]
    assert_bitwise_equal(coefficients, expected['coefficients'],
                         system.names()[body])

  # FlowWithFixedStep through NewInstance, beyond the current t_max (the facade
  # prolongs), every step appended.
  n = 37
  earth = system.bodies[system.index('Earth')]
  state = workloads.leo_probes(n, earth['q'], earth['v'], earth['mu'], 5)
  q = np.ascontiguousarray(state[0:3].T)
  v = np.ascontiguousarray(state[6:9].T)
  points = np.zeros(n, dtype=np.int32)
  last = np.zeros((n, 6))
  last_time = _d()
  t_flow = 40000.0
  status = shim.ft_flow_with_fixed_step(
      handle, n, oracle_lib.dptr(q), oracle_lib.dptr(v), 0.0,
      _capi.BLANES_MOAN_2002_SRKN_14A, 10.0, t_flow,
      points.ctypes.data_as(_i32p), oracle_lib.dptr(last), last_time)
  assert status == 0
  assert shim.ft_t_max(handle) >= t_flow
  oracle.prolong(t_flow)
  expected_state = state.copy()
  expected_time = np.zeros(2)
  oracle.flow_with_fixed_step(_capi.BLANES_MOAN_2002_SRKN_14A, 10.0, t_flow,
                              expected_state, expected_time)
  assert np.all(points == 4001)
  assert last_time.value == expected_time[0] == t_flow
  assert_bitwise_equal(last[:, 0:3].T, expected_state[0:3], 'positions')
  assert_bitwise_equal(last[:, 3:6].T, expected_state[6:9], 'velocities')

  # The history integrator (Quinlan1999Order8A at 10 s) on one instance flowed
  # twice: the instance keeps its previous steps between the calls.
  points[:] = 0
  status = shim.ft_flow_with_fixed_step_twice(
      handle, n, oracle_lib.dptr(q), oracle_lib.dptr(v), 0.0,
      _capi.QUINLAN_1999_ORDER_8A, 10.0, 45.0, 2000.0,
      points.ctypes.data_as(_i32p), oracle_lib.dptr(last), last_time)
  assert status == 0
  expected_state = state.copy()
  expected_time = np.zeros(2)
  oracle.flow_with_fixed_step(_capi.QUINLAN_1999_ORDER_8A, 10.0, 2000.0,
                              expected_state, expected_time)
  assert np.all(points == 201)
  assert last_time.value == expected_time[0] == 2000.0
  assert_bitwise_equal(last[:, 0:3].T, expected_state[0:3], 'history q')
  assert_bitwise_equal(last[:, 3:6].T, expected_state[6:9], 'history v')

  # FlowWithAdaptiveStep: every accepted step appended.
  probe = workloads.eccentric_probes(4, earth['q'], earth['v'], earth['mu'], 9)
  k = 2
  q1 = np.ascontiguousarray(probe[0:3, k])
  v1 = np.ascontiguousarray(probe[6:9, k])
  capacity = 5000
  n_points = _i64()
  times = np.zeros(capacity)
  dofs = np.zeros((capacity, 6))
  status = shim.ft_flow_with_adaptive_step(
      handle, oracle_lib.dptr(q1), oracle_lib.dptr(v1), 0.0, 30000.0, 1e-3,
      1e-3, 2**62, capacity, n_points, oracle_lib.dptr(times),
      oracle_lib.dptr(dofs))
  assert status == 0
  expected_state = np.ascontiguousarray(probe[:, k:k + 1])
  expected_time = np.zeros((2, 1))
  expected = oracle.flow_with_adaptive_step(30000.0, expected_state,
                                            expected_time, 1e-3, 1e-3,
                                            step_log_capacity=capacity)
  count = int(expected['accepted'][0])
  assert n_points.value == count > 5
  assert_bitwise_equal(times[:count], expected['step_log'][:count, 0])
  assert times[count - 1] == 30000.0
  assert_bitwise_equal(dofs[count - 1, 0:3], expected_state[0:3, 0])
  assert_bitwise_equal(dofs[count - 1, 3:6], expected_state[6:9, 0])

  # Point query.
  a = np.zeros(3)
  assert shim.ft_acceleration_on_massless_body(handle, oracle_lib.dptr(q1),
                                               100.0, oracle_lib.dptr(a)) == 0
  expected_a, _ = oracle.acceleration_on_massless_bodies(100.0,
                                                         q1.reshape(3, 1))
  assert_bitwise_equal(a, expected_a.ravel())

  # Jerk and Jacobian queries.
  v1 = np.array([1.0e3, -2.0e3, 0.5e3])
  earth = system.index('Earth')
  massless_jerk, massive_jerk, jacobian = np.zeros(3), np.zeros(3), np.zeros(9)
  assert shim.ft_jerk_and_jacobian(
      handle, oracle_lib.dptr(q1), oracle_lib.dptr(v1), earth, 100.0,
      oracle_lib.dptr(massless_jerk), oracle_lib.dptr(massive_jerk),
      oracle_lib.dptr(jacobian)) == 0
  assert_bitwise_equal(
      massless_jerk,
      oracle.jerk_on_massless_bodies(100.0, q1.reshape(3, 1),
                                     v1.reshape(3, 1)).ravel())
  assert_bitwise_equal(massive_jerk,
                       oracle.jerk_on_massive_bodies(100.0)[earth])
  assert_bitwise_equal(jacobian.reshape(3, 3),
                       oracle.jacobian_on_massive_bodies(100.0)[earth])
  shim.ft_destroy(handle)


@pytest.mark.parametrize('integrator', [
    _capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, _capi.QUINLAN_1999_ORDER_8A])
def test_facade_prolongs_the_reference_earth_moon_system(shim, sol_system,
                                                         integrator):
  """The Earth-Moon system of the reference's EphemerisTest
  (physics/ephemeris_test.cpp:134-174, epoch 1950) prolonged over one period
  through the facade: the polynomials are the oracle's, bit for bit -- and the
  oracle's lie inside the reference's windows (tests/test_oracle_kats.py)."""
  import earth_probe
  system, positions, velocities, _, period = earth_probe.set_up_earth_moon(
      sol_system)
  t0 = earth_probe.T0
  bodies, keep = system.c_bodies()
  q0 = np.ascontiguousarray(positions)
  v0 = np.ascontiguousarray(velocities)
  handle = shim.ft_create(len(system.bodies), bodies, oracle_lib.dptr(q0),
                          oracle_lib.dptr(v0), t0, 5e-3, 2.0**-24, integrator,
                          period / 100)
  assert handle
  oracle = OracleEphemeris(system, method=integrator, step=period / 100,
                           fitting_tolerance=5e-3, initial_positions=positions,
                           initial_velocities=velocities, initial_time=t0)
  assert shim.ft_prolong(handle, t0 + period) == 0
  assert oracle.prolong(t0 + period) == 0
  assert shim.ft_t_max(handle) == oracle.t_max()
  for body in range(2):
    expected = oracle.segments(body)
    count = shim.ft_segment_count(handle, body)
    assert count == len(expected['t_max']) == 13
    t_max, origin = np.zeros(count), np.zeros(count)
    degree = np.zeros(count, dtype=np.int32)
    coefficients = np.zeros((count, 18, 3))
    shim.ft_get_segments(handle, body, oracle_lib.dptr(t_max),
                         oracle_lib.dptr(origin),
                         degree.ctypes.data_as(_i32p),
                         oracle_lib.dptr(coefficients))
    assert_bitwise_equal(t_max, expected['t_max'])
    assert_bitwise_equal(origin, expected['origin'])
    assert list(degree) == list(expected['degree'])
    valid = np.arange(18)[None, :, None] <= degree[:, None, None]
    assert_bitwise_equal(np.where(valid, coefficients, 0.0),
                         np.where(valid, expected['coefficients'], 0.0))
  shim.ft_destroy(handle)


@pytest.mark.parametrize('integrator', [
    _capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, _capi.QUINLAN_1999_ORDER_8A])
def test_facade_special_cases_of_the_reference(shim, sol_system, integrator):
  """EphemerisTest.ProlongSpecialCases and .FlowWithAdaptiveStepSpecialCase
  (physics/ephemeris_test.cpp:180-266) through the facade."""
  import earth_probe
  system, positions, velocities, _, period = earth_probe.set_up_earth_moon(
      sol_system)
  t0 = earth_probe.T0
  bodies, keep = system.c_bodies()
  q0 = np.ascontiguousarray(positions)
  v0 = np.ascontiguousarray(velocities)
  handle = shim.ft_create(len(system.bodies), bodies, oracle_lib.dptr(q0),
                          oracle_lib.dptr(v0), t0, 5e-3, 2.0**-24, integrator,
                          period / 100)
  assert handle
  # ProlongSpecialCases.
  assert shim.ft_t_max(handle) == -np.inf  # InfinitePast.
  assert shim.ft_t_min(handle) == np.inf   # InfiniteFuture.
  assert shim.ft_empty(handle) == 1
  assert shim.ft_prolong(handle, t0 + period) == 0
  assert shim.ft_empty(handle) == 0
  assert shim.ft_t_min(handle) == t0
  t_max = shim.ft_t_max(handle)
  assert t0 + period <= t_max
  assert shim.ft_prolong(handle, t0 + period / 2) == 0
  assert shim.ft_t_max(handle) == t_max
  last_t = ((t0 + period) + t_max) / 2
  shim.ft_prolong(handle, last_t)
  assert shim.ft_t_max(handle) == t_max
  # FlowWithAdaptiveStepSpecialCase: a probe 10^9 m from the Earth flowed over a
  # period, then "flowed" to the time it is already at.
  q = positions[0] + np.array([0.0, 1e9, 0.0])
  v = np.array([1e3, 1e3, 1e3])
  capacity = 32768
  count = ctypes.c_int64()
  times = np.zeros(capacity)
  dofs = np.zeros((capacity, 6))
  assert shim.ft_flow_with_adaptive_step(
      handle, oracle_lib.dptr(q), oracle_lib.dptr(v), t0, t0 + period, 1e-9,
      2.6e-15, 2**62, capacity, ctypes.byref(count), oracle_lib.dptr(times),
      oracle_lib.dptr(dofs)) == 0
  assert 0 < count.value <= capacity
  back_time = times[count.value - 1]
  assert back_time == t0 + period
  back_q = np.ascontiguousarray(dofs[count.value - 1, 0:3])
  back_v = np.ascontiguousarray(dofs[count.value - 1, 3:6])
  again = ctypes.c_int64(-1)
  assert shim.ft_flow_with_adaptive_step(
      handle, oracle_lib.dptr(back_q), oracle_lib.dptr(back_v), back_time,
      back_time, 1e-9, 2.6e-15, 2**62, capacity, ctypes.byref(again),
      oracle_lib.dptr(times), oracle_lib.dptr(dofs)) == 0
  assert again.value == 0
  shim.ft_destroy(handle)
