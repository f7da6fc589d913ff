"""GPU parity tests of the massless-body path, through the C ABI, against the
CPU oracle on the same seeded inputs: bit-exact (IEEE binary64 without
contraction on both sides)."""
import ctypes
import math

import numpy as np
import pytest

from principia_b200 import _capi, workloads
from principia_b200.solar_system import SolarSystem

from conftest import assert_bitwise_equal
import oracle_lib
from oracle_lib import OracleEphemeris

pytestmark = pytest.mark.gpu


def earth_state(system, oracle, t=0.0):
  i = system.index('Earth')
  if t == 0.0:
    return (np.array(system.bodies[i]['q']), np.array(system.bodies[i]['v']),
            system.bodies[i]['mu'])
  raise NotImplementedError


def test_sm_clock_measured_on_the_device():
  """pb_measure_sm_clock and its two halves (what bench.py reports as
  clocks.sm_mhz): a plausible clock, no larger than the maximum NVML states
  plus the granularity of the global timer."""
  from principia_b200.ephemeris import (measure_sm_clock,
                                        sm_clock_probe_read,
                                        sm_clock_probe_start)
  mhz = measure_sm_clock(0, 500.0)
  assert 200.0 < mhz < 3000.0
  sm_clock_probe_start(0, 100.0, 500.0)
  again = sm_clock_probe_read(0)
  assert 200.0 < again < 3000.0
  try:
    import pynvml
    pynvml.nvmlInit()
    handle = pynvml.nvmlDeviceGetHandleByIndex(0)
    maximum = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)
    assert mhz <= maximum * 1.01 and again <= maximum * 1.01
  except ImportError:
    pass


def test_elementary_functions_on_device():
  lib = _capi.library()
  orc = oracle_lib.library()
  rng = np.random.default_rng(11)
  x = np.concatenate([rng.uniform(-10, 10, 100000),
                      rng.uniform(-1e6, 1e6, 100000),
                      rng.uniform(-2e9, 2e9, 50000),
                      10.0**rng.uniform(-30, 30, 50000)])
  n = len(x)
  s, c, r = np.zeros(n), np.zeros(n), np.zeros(n)
  f = lib.pb_internal_elementary_functions
  f.argtypes = [ctypes.c_int32, ctypes.c_int64] + [_capi.c_double_p] * 4
  assert f(0, n, oracle_lib.dptr(x), oracle_lib.dptr(s), oracle_lib.dptr(c),
           oracle_lib.dptr(r)) == 0
  step = 37  # The oracle (binary128) is slow: check a strided subset...
  for i in range(0, n, step):
    if abs(x[i]) < 2.0**31:  # The supported range of CrSinCos.
      assert s[i] == orc.orc_cr_sin(x[i]), x[i]
      assert c[i] == orc.orc_cr_cos(x[i]), x[i]
    if x[i] > 0:
      assert r[i] == orc.orc_root(4, x[i]), x[i]
  # ... and all of it against the host build of the same header, which
  # tests/test_host_arithmetic.py pins to the oracle on 10^5 points.
  assert np.all(np.abs(s) <= 1) and np.all(np.abs(c) <= 1)
  positive = x > 0
  assert np.max(np.abs(r[positive]**4 / x[positive] - 1)) < 1e-15


def test_structure_matches_oracle(sol_system, sol_oracle, sol_gpu):
  order, n_oblate = sol_gpu.internal_order()
  expected_order, expected_n_oblate = sol_oracle.internal_order()
  assert list(order) == list(expected_order)
  assert n_oblate == expected_n_oblate == 11
  for i, b in enumerate(sol_system.bodies):
    if b['oblate']:
      inner, sectoral = sol_gpu.geopotential_damping(i)
      expected_inner, expected_sectoral = sol_oracle.damping(i)
      assert_bitwise_equal(inner, expected_inner, b['name'])
      assert sectoral == expected_sectoral
  assert sol_gpu.t_min() == sol_oracle.t_min() == 0.0
  assert sol_gpu.t_max() == sol_oracle.t_max() == 172800.0


def test_golden_segments_are_the_oracles(sol_system, sol_oracle):
  golden = workloads.load_sol_ephemeris_segments()
  for i in range(len(sol_system.bodies)):
    seg = sol_oracle.segments(i)
    assert seg['t_min'] == golden[i]['t_min']
    assert_bitwise_equal(seg['t_max'], golden[i]['t_max'])
    assert_bitwise_equal(seg['origin'], golden[i]['origin'])
    assert list(seg['degree']) == list(golden[i]['degree'])
    assert_bitwise_equal(seg['coefficients'], golden[i]['coefficients'])


def test_evaluate_all_positions(sol_oracle, sol_gpu):
  rng = np.random.default_rng(1)
  times = np.concatenate([rng.uniform(0, 172800.0, 40),
                          [0.0, 4800.0, 4800.000001, 172800.0]])
  for t in times:
    assert_bitwise_equal(sol_gpu.evaluate_all_positions(t),
                         sol_oracle.evaluate_all_positions(t), 't=%r' % t)
  with pytest.raises(Exception):
    sol_gpu.evaluate_all_positions(172800.5)


def random_points_near(rng, centres, n, low=6.6e6, high=2e9):
  distance = np.exp(rng.uniform(math.log(low), math.log(high), n))
  direction = rng.normal(size=(3, n))
  direction /= np.linalg.norm(direction, axis=0)
  centre = centres[rng.integers(0, len(centres), n)].T
  return centre + distance * direction


def test_acceleration_on_massless_bodies(sol_system, sol_oracle, sol_gpu):
  rng = np.random.default_rng(42)
  names = sol_system.names()
  for t in (0.0, 12345.678, 86400.0, 172800.0):
    positions = sol_oracle.evaluate_all_positions(t)
    centres = positions[[names.index(k) for k in
                         ('Earth', 'Earth', 'Earth', 'Moon', 'Mars', 'Venus',
                          'Jupiter', 'Sun', 'Io', 'Vesta')]]
    q = random_points_near(rng, centres, 4099)
    a, status = sol_gpu.compute_gravitational_acceleration_on_massless_bodies(
        t, q)
    expected, expected_status = sol_oracle.acceleration_on_massless_bodies(t, q)
    assert_bitwise_equal(a, expected, 't=%r' % t)
    # Some of the points fall inside the Sun or Jupiter: both sides must
    # report the collision.
    assert status == expected_status == _capi.PB_OUT_OF_RANGE
  # Close to the Moon: its degree-30 field (generic path).
  positions = sol_oracle.evaluate_all_positions(1000.0)
  q = random_points_near(rng, positions[[names.index('Moon')]], 513, 1.75e6,
                         1e7)
  a, status = sol_gpu.compute_gravitational_acceleration_on_massless_bodies(
      1000.0, q)
  expected, expected_status = sol_oracle.acceleration_on_massless_bodies(
      1000.0, q)
  assert_bitwise_equal(a, expected, 'Moon')
  # Collision: one probe inside the Earth (physics/ephemeris_test.cpp:763-811).
  q[:, 7] = positions[names.index('Earth')] + [1e6, 2e6, 0.0]
  a, status = sol_gpu.compute_gravitational_acceleration_on_massless_bodies(
      1000.0, q)
  expected, expected_status = sol_oracle.acceleration_on_massless_bodies(
      1000.0, q)
  assert status == expected_status == _capi.PB_OUT_OF_RANGE
  assert_bitwise_equal(a, expected, 'collision')
  # Empty batch.
  a, status = sol_gpu.compute_gravitational_acceleration_on_massless_bodies(
      0.0, np.zeros((3, 0)))
  assert a.shape == (3, 0) and status == 0


def test_potential(sol_system, sol_oracle, sol_gpu):
  rng = np.random.default_rng(9)
  names = sol_system.names()
  positions = sol_oracle.evaluate_all_positions(5000.0)
  centres = positions[[names.index(k) for k in ('Earth', 'Moon', 'Jupiter')]]
  q = random_points_near(rng, centres, 1000)
  assert_bitwise_equal(sol_gpu.compute_gravitational_potential(5000.0, q),
                       sol_oracle.potential(5000.0, q))


def test_acceleration_on_massive_bodies(sol_system, sol_oracle, sol_gpu):
  for t in (0.0, 40000.0):
    assert_bitwise_equal(
        sol_gpu.compute_gravitational_acceleration_on_massive_bodies(t),
        sol_oracle.acceleration_on_massive_bodies(t), 't=%r' % t)
  rng = np.random.default_rng(2)
  positions = sol_oracle.evaluate_all_positions(100.0)
  positions += rng.normal(size=positions.shape) * 1e3
  assert_bitwise_equal(
      sol_gpu.compute_gravitational_acceleration_on_massive_bodies(
          100.0, positions),
      sol_oracle.acceleration_on_massive_bodies(100.0, positions))


def leo(system, n, seed=42):
  q, v, mu = earth_state(system, None)
  return workloads.leo_probes(n, q, v, mu, seed)


@pytest.mark.parametrize('method,evaluations', [
    (_capi.BLANES_MOAN_2002_SRKN_14A, 14),
    (_capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, 6)])
def test_flow_with_fixed_step_matches_oracle(sol_system, sol_oracle, sol_gpu,
                                             method, evaluations):
  n = 1500  # Ragged: not a multiple of the block size.
  state = leo(sol_system, n)
  expected_state = state.copy()
  time = np.array([0.0, 0.0])
  expected_time = time.copy()
  t_final = 605.0  # Not a multiple of the step: stops at 600 s.
  result = sol_gpu.flow_with_fixed_step(method, 10.0, t_final, state, time,
                                        sample_every=7, max_samples=5)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, t_final,
                                             expected_state, expected_time,
                                             sample_every=7, max_samples=5)
  assert result['n_steps'] == expected['n_steps'] == 60
  assert result['status'] == expected['status'] == 0
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'state')
  assert len(result['samples']) == len(expected['samples']) == 5
  assert_bitwise_equal(result['sample_times'], expected['sample_times'])
  assert_bitwise_equal(result['samples'], expected['samples'], 'samples')
  # Restart from the updated state (Instance::Solve is restartable), across a
  # chunk boundary of the stage table (> 256 steps).
  result = sol_gpu.flow_with_fixed_step(method, 10.0, 3500.0, state, time)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, 3500.0,
                                             expected_state, expected_time)
  assert result['n_steps'] == expected['n_steps'] == 290
  assert_bitwise_equal(time, expected_time, 'time 2')
  assert_bitwise_equal(state, expected_state, 'state 2')


def test_flow_with_fixed_step_backward_and_empty(sol_system, sol_oracle,
                                                 sol_gpu):
  method = _capi.BLANES_MOAN_2002_SRKN_14A
  state = leo(sol_system, 64, seed=3)
  time = np.array([0.0, 0.0])
  sol_gpu.flow_with_fixed_step(method, 10.0, 300.0, state, time)
  expected_state, expected_time = state.copy(), time.copy()
  result = sol_gpu.flow_with_fixed_step(method, -10.0, 0.0, state, time)
  expected = sol_oracle.flow_with_fixed_step(method, -10.0, 0.0,
                                             expected_state, expected_time)
  assert result['n_steps'] == expected['n_steps'] == 30
  assert_bitwise_equal(state, expected_state, 'backward')
  assert_bitwise_equal(time, expected_time)
  # No probes: only the time advances.
  empty = np.zeros((12, 0))
  time = np.array([0.0, 0.0])
  result = sol_gpu.flow_with_fixed_step(method, 10.0, 100.0, empty, time)
  assert result['n_steps'] == 10 and time[0] == 100.0
  # t_final closer than one step: nothing happens.
  state = leo(sol_system, 8)
  before = state.copy()
  time = np.array([0.0, 0.0])
  result = sol_gpu.flow_with_fixed_step(method, 10.0, 5.0, state, time)
  assert result['n_steps'] == 0 and time[0] == 0.0
  assert_bitwise_equal(state, before)
  # Outside the ephemeris: refused (the facade prolongs first).
  with pytest.raises(Exception):
    sol_gpu.flow_with_fixed_step(method, 10.0, 2e5, state, time)


def test_collision_status_and_continuation(sol_system, sol_oracle, sol_gpu):
  """physics/ephemeris_test.cpp:763-811: an apple dropped above the Earth; the
  fixed-step loop reports OutOfRange but keeps integrating."""
  q, v, mu = earth_state(sol_system, None)
  state = np.zeros((12, 3))
  state[0:3] = q[:, None] + np.array([[6371008.4 + 10.0, 7e6, 0.0],
                                      [0.0, 0.0, 7e6], [0.0, 0.0, 0.0]])
  state[6:9] = v[:, None]
  expected_state = state.copy()
  time, expected_time = np.zeros(2), np.zeros(2)
  method = _capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL
  result = sol_gpu.flow_with_fixed_step(method, 1.0, 400.0, state, time)
  expected = sol_oracle.flow_with_fixed_step(method, 1.0, 400.0,
                                             expected_state, expected_time)
  assert result['status'] == expected['status'] == _capi.PB_OUT_OF_RANGE
  assert result['n_steps'] == expected['n_steps'] == 400
  assert_bitwise_equal(state, expected_state, 'collision state')


def test_fixed_step_flow_with_intrinsic_accelerations(sol_system, sol_oracle,
                                                      sol_gpu):
  """NewInstance with intrinsic accelerations (ephemeris_body.hpp:346-412): a
  tabulated inertially fixed burn per probe, SRKN14A, against the oracle."""
  method = _capi.BLANES_MOAN_2002_SRKN_14A
  n = 300
  state = leo(sol_system, n, seed=23)
  rng = np.random.default_rng(8)
  burns = np.zeros((8, n))
  direction = rng.normal(size=(3, n))
  burns[0:3] = direction / np.linalg.norm(direction, axis=0)
  burns[3] = rng.uniform(1e3, 5e5, n)
  burns[4] = rng.uniform(5e3, 5e4, n)
  burns[5] = burns[3] / (9.80665 * 320.0)
  burns[6] = rng.uniform(-100.0, 2000.0, n)
  burns[7] = burns[6] + rng.uniform(10.0, 600.0, n)
  expected_state = state.copy()
  time, expected_time = np.zeros(2), np.zeros(2)
  result = sol_gpu.flow_with_fixed_step(method, 10.0, 3000.0, state, time,
                                        sample_every=50, max_samples=6,
                                        burns=burns)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, 3000.0,
                                             expected_state, expected_time,
                                             sample_every=50, max_samples=6,
                                             burns=burns)
  assert result['n_steps'] == expected['n_steps'] == 300
  assert_bitwise_equal(state, expected_state, 'state')
  assert_bitwise_equal(result['samples'], expected['samples'], 'samples')
  free = leo(sol_system, n, seed=23)
  sol_gpu.flow_with_fixed_step(method, 10.0, 3000.0, free, np.zeros(2))
  assert np.linalg.norm(free[0:3] - state[0:3], axis=0).min() > 1.0


def test_ten_thousand_fixed_steps_match_oracle(sol_system, sol_oracle,
                                               sol_gpu):
  """BASELINE.json's tolerance is a relative position error <= 1e-12 after
  10^4 fixed steps; the flow is bit-identical to the oracle's after 10^4 steps
  of SRKN14A at 10 s (27.8 h, 1.4e5 right-hand sides per probe, 40 launches)."""
  method = _capi.BLANES_MOAN_2002_SRKN_14A
  n = 48
  state = leo(sol_system, n, seed=11)
  expected_state = state.copy()
  time, expected_time = np.zeros(2), np.zeros(2)
  result = sol_gpu.flow_with_fixed_step(method, 10.0, 1.0e5, state, time)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, 1.0e5,
                                             expected_state, expected_time)
  assert result['n_steps'] == expected['n_steps'] == 10000
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'state after 10^4 steps')


def test_full_size_flow_properties(sol_system, sol_oracle, sol_gpu):
  """BASELINE.json config 2 at full size (10^5 probes): size-independent
  properties -- probes are independent, so any subset equals the oracle bit for
  bit; the time-reversible SRKN14A retraces its steps; the specific energy
  relative to the Earth is conserved to the level of the perturbations."""
  n = 100000
  method = _capi.BLANES_MOAN_2002_SRKN_14A
  state0 = leo(sol_system, n)
  state = state0.copy()
  time = np.zeros(2)
  # At this size the call is sliced over helper streams and sub-chunks of
  # steps; samples cross slice boundaries.
  result = sol_gpu.flow_with_fixed_step(method, 10.0, 1000.0, state, time,
                                        sample_every=33, max_samples=3)
  assert result['n_steps'] == 100 and result['status'] == 0
  subset = np.r_[0:40, 24990:25010, 49980:50020, n - 40:n]
  expected_state = np.ascontiguousarray(state0[:, subset])
  expected_time = np.zeros(2)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, 1000.0,
                                             expected_state, expected_time,
                                             sample_every=33, max_samples=3)
  assert_bitwise_equal(state[:, subset], expected_state, 'subset')
  assert_bitwise_equal(result['sample_times'], expected['sample_times'])
  assert_bitwise_equal(result['samples'][:, :, subset], expected['samples'],
                       'subset samples')
  # Energy relative to the Earth.
  earth = sol_system.index('Earth')
  mu = sol_system.bodies[earth]['mu']

  def energy(s, t):
    positions = sol_oracle.evaluate_all_positions(t)
    pe = positions[earth]
    # Earth's velocity by a centred difference of its trajectory.
    dt = 1.0
    ve = (sol_oracle.evaluate_all_positions(min(t + dt, 172800.0))[earth] -
          sol_oracle.evaluate_all_positions(max(t - dt, 0.0))[earth]) / (
              min(t + dt, 172800.0) - max(t - dt, 0.0))
    r = s[0:3] - pe[:, None]
    w = s[6:9] - ve[:, None]
    return 0.5 * np.sum(w * w, axis=0) - mu / np.linalg.norm(r, axis=0)

  e0 = energy(state0, 0.0)
  e1 = energy(state, 1000.0)
  assert np.max(np.abs(e1 / e0 - 1)) < 6e-3  # J2-level exchange, no drift.
  # Reversibility: integrate back; positions return to the start within
  # round-off accumulated over 200 steps.
  sol_gpu.flow_with_fixed_step(method, -10.0, 0.0, state, time)
  assert time[0] == 0.0
  displacement = np.linalg.norm(state[0:3] - state0[0:3], axis=0)
  assert np.max(displacement) < 1e-2


def mixed_batch(system, n):
  """LEO probes with, scattered among them, probes in low lunar orbit (the
  Moon's degree-30 field: neither the hot nor the warm body of the batch, whose
  first probe is in LEO) and probes far from everything (no harmonics at
  all)."""
  state = leo(system, n)
  moon = system.bodies[system.index('Moon')]
  earth = system.bodies[system.index('Earth')]
  lunar = workloads.leo_probes(n, moon['q'], moon['v'], moon['mu'], seed=7)
  # Scale the LEO-sized orbits (6.6e6 .. 8e6 m) down to 1.8e6 .. 2.2e6 m.
  k = 1.8e6 / 6.6e6
  lunar_q = np.asarray(moon['q'])[:, None] + k * (
      lunar[0:3] - np.asarray(moon['q'])[:, None])
  lunar_v = np.asarray(moon['v'])[:, None] + (
      lunar[6:9] - np.asarray(moon['v'])[:, None]) / math.sqrt(k)
  far = workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                   seed=11)
  special = np.zeros(n, dtype=int)
  special[5::997] = 1   # Lunar.
  special[11::1499] = 2  # Eccentric, up to 4e8 m from the Earth.
  state[0:3, special == 1] = lunar_q[:, special == 1]
  state[6:9, special == 1] = lunar_v[:, special == 1]
  state[:, special == 2] = far[:, special == 2]
  return state, special


@pytest.mark.parametrize('method', [
    _capi.BLANES_MOAN_2002_SRKN_14A,
    _capi.QUINLAN_1999_ORDER_8A,
])
def test_mixed_batch_takes_every_path(sol_system, sol_oracle, sol_gpu, method):
  """A batch whose probes need different evaluations of the right-hand side:
  harmonics of the hot and warm bodies (LEO), of a body that is neither (low
  lunar orbit: the body-by-body evaluation and the thread-local-memory
  harmonics), none at all.  The kernels pick the path per probe (per block for
  the warp-specialised one); the bits do not depend on it."""
  n = 80000  # Enough for the widest kernel shape (512 probes per SM).
  state0, special = mixed_batch(sol_system, n)
  state = state0.copy()
  time = np.zeros(2)
  result = sol_gpu.flow_with_fixed_step(method, 10.0, 200.0, state, time)
  assert result['n_steps'] == 20 and result['status'] == 0
  assert sol_gpu.last_flow_path()[1]  # Some probe took the generic path.
  subset = np.unique(np.concatenate(
      [np.flatnonzero(special == 1)[:40], np.flatnonzero(special == 2)[:20],
       np.r_[0:30, 40000:40030, n - 30:n]]))
  expected_state = np.ascontiguousarray(state0[:, subset])
  expected_time = np.zeros(2)
  expected = sol_oracle.flow_with_fixed_step(method, 10.0, 200.0,
                                             expected_state, expected_time)
  assert expected['n_steps'] == 20
  assert_bitwise_equal(state[:, subset], expected_state, 'mixed subset')


def test_branch_free_sqrt_and_division_are_ieee():
  """FastSqrt / FastDivide (pb_math.cuh) against the IEEE operators on the
  device: 2e9 operand sets, no mismatch allowed."""
  lib = _capi.library()
  f = lib.pb_internal_fast_math_check
  f.argtypes = [ctypes.c_int32, ctypes.c_int64,
                ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
  total = 0
  for seed in (1, 2, 3):
    checked, bad = ctypes.c_int64(), ctypes.c_int64()
    assert f(0, seed, ctypes.byref(checked), ctypes.byref(bad)) == 0
    assert bad.value == 0
    total += checked.value
  assert total > 2e9


def _world_system(cos, sin, degree, zonal):
  """The 17 m^3/s^2 body of GeopotentialTest (physics/geopotential_test.cpp
  :84-91,162-168) alone at the origin."""
  import test_oracle_kats as kats
  size = (degree + 1) * (degree + 2) // 2
  c, s = [0.0] * size, [0.0] * size
  for (n, m), value in cos.items():
    c[n * (n + 1) // 2 + m] = value / kats._normalization(n, m)
  for (n, m), value in sin.items():
    s[n * (n + 1) // 2 + m] = value / kats._normalization(n, m)
  body = dict(name='World', mu=17.0, rotating=True, min_radius=1.0,
              mean_radius=1.0, reference_angle=3.0, reference_instant=4.0,
              angular_frequency=-1.5, right_ascension=0.0,
              declination=90 * (math.pi / 180), oblate=True, degree=degree,
              zonal=zonal, reference_radius=1.0, cos=c, sin=s,
              q=[0.0, 0.0, 0.0], v=[0.0, 0.0, 0.0])
  return SolarSystem([body], 0.0)


@pytest.mark.parametrize('tolerance', [0.0, 2.0**-24])
@pytest.mark.parametrize('shape', ['j2', 'c22s22', 'j3', 'degree6'])
def test_geopotential_test_bodies_on_the_device(shape, tolerance):
  """The bodies of GeopotentialTest.J2 / C22S22 / J3 (geopotential_test.cpp
  :175-320) and a dense degree-6 one, as one-body ephemerides at rest: the
  accelerations and potentials on the device are the oracle's, bit for bit, at
  the test's special points (on the axis, in the equatorial plane) and at
  random ones, and the exact symmetries the reference pins hold on the GPU."""
  from principia_b200.ephemeris import Ephemeris
  if shape == 'j2':
    system = _world_system({(2, 0): -6.0}, {}, 2, True)
  elif shape == 'c22s22':
    system = _world_system({(2, 0): -6.0, (2, 2): 10.0}, {(2, 2): -13.0}, 2,
                           False)
  elif shape == 'j3':
    system = _world_system({(2, 0): -6.0, (2, 2): 1e-20, (3, 0): 5.0},
                           {(2, 2): 1e-20}, 3, False)
  else:
    rng = np.random.Generator(np.random.MT19937(6))
    cos = {(n, m): rng.normal() for n in range(2, 7) for m in range(n + 1)}
    sin = {(n, m): rng.normal() for n in range(2, 7) for m in range(1, n + 1)}
    system = _world_system(cos, sin, 6, False)
  gpu = Ephemeris(system, geopotential_tolerance=tolerance)
  oracle = OracleEphemeris(system, geopotential_tolerance=tolerance)
  segment = dict(t_min=-100.0, t_max=np.array([100.0]), origin=np.array([0.0]),
                 degree=np.array([3], dtype=np.int32),
                 coefficients=np.zeros((1, 18, 3)))
  gpu.append_segments(0, segment)
  oracle.append_segments(0, segment)
  rng = np.random.Generator(np.random.MT19937(8))
  special = np.array([[0.0, 0.0, 10.0], [30.0, 40.0, 0.0], [1e2, 0.0, 1e2],
                      [0.0, 0.0, -3.0], [5.0, 0.0, 0.0], [0.0, 7.0, 0.0]]).T
  direction = rng.normal(size=(3, 3000))
  direction /= np.linalg.norm(direction, axis=0)
  q = np.concatenate(
      [special, direction * 10.0**rng.uniform(0.2, 4.0, 3000)], axis=1)
  for t in (0.0, 4.0, -37.5, 99.0):
    a, status = gpu.compute_gravitational_acceleration_on_massless_bodies(t, q)
    expected, expected_status = oracle.acceleration_on_massless_bodies(t, q)
    assert status == expected_status
    assert_bitwise_equal(a, expected, '%s t=%r' % (shape, t))
    assert_bitwise_equal(gpu.compute_gravitational_potential(t, q),
                         oracle.potential(t, q), '%s potential' % shape)
  if tolerance == 0.0 and shape in ('j2', 'j3'):
    # In the equatorial plane the acceleration of a zonal field points to the
    # axis: a_x / a_y == 0.75 exactly (AlmostEquals(0.75, 0)), point mass
    # included.
    a, _ = gpu.compute_gravitational_acceleration_on_massless_bodies(0.0, q)
    assert a[0, 1] / a[1, 1] == 0.75


def test_sin_cos_hard_cases_on_the_device():
  """numerics/sin_cos_test.cpp:363-413 (HardRounding, HardReduction within the
  supported range): the device's correctly rounded sine and cosine on the
  reference's hard-to-round arguments, AlmostEquals(·, 0)."""
  sines = {
      1.777288458404935767021016: 0.9787561457198967196367773,
      5.528810471911395296729097: -0.6848332450871304488693046,
      2.670333644894535396474566: 0.4540084183741445456039384,
      -2.496680544289484160458414: -0.6011282027544306294509797,
      3.348980952786005715893225: -0.2059048676737040700634683,
      3.523452575387961971387085: -0.3726470704519433130297035,
      1.881458091523454001503524: 0.9521314843257784876761001,
      -1.763163156774038675678185: -0.9815544881044536151825223,
      -2.58105062034143273308473: -0.5316453603071467637339815,
      1.657419885978818285821035: 0.99625052493662308306103561,
      5.094301519947547873812255: -0.9279535374988051033005616,
      5.262137362438826571064965: -0.8526560125576488347044409,
      float.fromhex('0x5ad7bcep-8'): 0.99999999999999999818,
  }
  cosines = {
      3.912491942337291916942377: -0.7172843528140595004137653,
      1.486604973422413600303571: 0.0840919279825555407437241,
      -6.265702600230396157598989: 0.99984718137127853720984932,
      -3.885819786017697730073905: -0.7356116652133562472394118,
      -5.026994177012682030181168: 0.3094410694753661206223057,
      0.2388111698570396512764091: 0.9716198764286143041422587,
      float.fromhex('0x5ad7bcep-8'): 1.9060601347136373148e-9,
  }
  lib = _capi.library()
  x = np.array(list(sines) + list(cosines))
  n = len(x)
  s, c, r = np.zeros(n), np.zeros(n), np.zeros(n)
  f = lib.pb_internal_elementary_functions
  f.argtypes = [ctypes.c_int32, ctypes.c_int64] + [_capi.c_double_p] * 4
  assert f(0, n, oracle_lib.dptr(x), oracle_lib.dptr(s), oracle_lib.dptr(c),
           oracle_lib.dptr(r)) == 0
  for i, (argument, expected) in enumerate(sines.items()):
    assert s[i] == expected, argument
  for i, (argument, expected) in enumerate(cosines.items()):
    assert c[len(sines) + i] == expected, argument
