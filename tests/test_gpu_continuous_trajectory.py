"""GPU parity of the device-side ContinuousTrajectory production (SURVEY.md §8f
rank 1: Append + ComputeBestNewhallApproximation + Newhall fit + evaluation)
against the CPU oracle's ContinuousTrajectory: bit-exact polynomials, degrees
and evaluations."""
import ctypes

import numpy as np
import pytest

from principia_b200 import _capi
from principia_b200.ephemeris import Ephemeris, StatusError
from principia_b200.nbody import ContinuousTrajectories, NBodyEnsemble
from principia_b200.solar_system import SolarSystem

from conftest import assert_bitwise_equal
import oracle_lib

pytestmark = pytest.mark.gpu


class OracleTrajectory:

  def __init__(self, step, tolerance):
    self.lib = oracle_lib.library()
    self.handle = self.lib.orc_continuous_trajectory_create(step, tolerance)

  def __del__(self):
    self.lib.orc_continuous_trajectory_destroy(self.handle)

  def append(self, time, q, v):
    q = np.ascontiguousarray(q, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    return self.lib.orc_continuous_trajectory_append(
        self.handle, time, oracle_lib.dptr(q), oracle_lib.dptr(v))

  def segments(self):
    count = self.lib.orc_continuous_trajectory_segment_count(self.handle)
    t_max, origin = np.zeros(count), np.zeros(count)
    degree = np.zeros(count, dtype=np.int32)
    coefficients = np.zeros((count, 18, 3))
    for s in range(count):
      d = ctypes.c_int32()
      self.lib.orc_continuous_trajectory_get_segment(
          self.handle, s, oracle_lib.dptr(t_max[s:]),
          oracle_lib.dptr(origin[s:]), ctypes.byref(d),
          oracle_lib.dptr(coefficients[s]))
      degree[s] = d.value
    return dict(t_max=t_max, origin=origin, degree=degree,
                coefficients=coefficients)

  def evaluate(self, t):
    q, v = np.zeros(3), np.zeros(3)
    self.lib.orc_continuous_trajectory_evaluate(self.handle, t,
                                                oracle_lib.dptr(q),
                                                oracle_lib.dptr(v))
    return q, v


def kepler_states(n, times, seed):
  """Planar-ish ellipses of varied period, eccentricity and orientation,
  sampled at `times`: (len(times), 3, n) positions and velocities.  The periods
  span 20 to 20000 steps so that the fits settle at every degree from 3 to 17,
  and some never reach the tolerance (the unstable branch)."""
  rng = np.random.Generator(np.random.MT19937(seed))
  step = times[1] - times[0]
  period = step * 10.0**rng.uniform(1.3, 4.3, n)
  e = rng.uniform(0.0, 0.6, n)
  a = 10.0**rng.uniform(1, 11, n)
  phase = rng.uniform(0, 2 * np.pi, n)
  tilt = rng.uniform(0, np.pi, n)
  node = rng.uniform(0, 2 * np.pi, n)
  mean_motion = 2 * np.pi / period
  q = np.zeros((len(times), 3, n))
  v = np.zeros((len(times), 3, n))
  for k, t in enumerate(times):
    mean_anomaly = phase + mean_motion * t
    eccentric = mean_anomaly.copy()
    for _ in range(50):
      eccentric = mean_anomaly + e * np.sin(eccentric)
    x = a * (np.cos(eccentric) - e)
    y = a * np.sqrt(1 - e**2) * np.sin(eccentric)
    rate = mean_motion / (1 - e * np.cos(eccentric))
    vx = -a * np.sin(eccentric) * rate
    vy = a * np.sqrt(1 - e**2) * np.cos(eccentric) * rate
    for (px, py, out) in ((x, y, q), (vx, vy, v)):
      x1 = px * np.cos(node) - py * np.cos(tilt) * np.sin(node)
      y1 = px * np.sin(node) + py * np.cos(tilt) * np.cos(node)
      z1 = py * np.sin(tilt)
      out[k, 0], out[k, 1], out[k, 2] = x1, y1, z1
  return q, v


def compare(gpu, oracles, indices, evaluation_times):
  positions = {}
  for t in evaluation_times:
    positions[t] = gpu.evaluate(t)
  for i in indices:
    actual = gpu.segments(i)
    expected = oracles[i].segments()
    assert_bitwise_equal(actual['t_max'], expected['t_max'], 't_max %d' % i)
    assert_bitwise_equal(actual['origin'], expected['origin'], 'origin %d' % i)
    assert list(actual['degree']) == list(expected['degree']), i
    assert_bitwise_equal(actual['coefficients'], expected['coefficients'],
                         'coefficients %d' % i)
    for t in evaluation_times:
      q, v = oracles[i].evaluate(t)
      assert_bitwise_equal(positions[t][0][:, i], q, 'q %d at %r' % (i, t))
      assert_bitwise_equal(positions[t][1][:, i], v, 'v %d at %r' % (i, t))


def test_newhall_fit_of_keplerian_orbits_matches_oracle():
  n, step, tolerance = 600, 600.0, 1e-3
  n_points = 8 * 110 + 4  # More than max_degree_age = 100 segments.
  times = 1000.0 + step * np.arange(n_points)
  q, v = kepler_states(n, times, seed=42)
  # One trajectory turns into noise: its adjusted tolerance grows a million
  # times at once, the "apocalypse" of continuous_trajectory_body.hpp:851-857.
  noise = np.random.Generator(np.random.MT19937(1)).normal(
      size=(len(times) - 403, 3))
  q[403:, :, n - 1] += 1e9 * noise
  gpu = ContinuousTrajectories(n, step, tolerance, max_segments=112)
  oracles = [OracleTrajectory(step, tolerance) for _ in range(n)]
  assert gpu.t_min() == np.inf and gpu.t_max() == -np.inf
  apocalypse = False
  for k, t in enumerate(times):
    expected_status = 0
    for i in range(n):
      expected_status |= oracles[i].append(t, q[k, :, i], v[k, :, i])
    try:
      gpu.append(t, q[k], v[k])
      status = 0
    except StatusError as error:
      status = error.code
    apocalypse = apocalypse or expected_status != 0
    # The device reports the apocalypse of any trajectory, from then on.
    assert (status != 0) == apocalypse, (k, status, expected_status)
  assert apocalypse
  assert gpu.segment_count() == 110
  assert gpu.t_min() == times[0]
  assert gpu.t_max() == times[8 * 110]
  compare(gpu, oracles, range(n),
          [times[0], times[3] + 1.0, times[8], times[8] + 1e-3,
           times[8 * 57] - 17.0, times[8 * 110]])
  degrees = np.array([gpu.segments(i)['degree'] for i in range(0, n, 7)])
  # The sample exercises the whole range of degrees.
  assert degrees.min() == 3 and degrees.max() == 17
  with pytest.raises(StatusError):
    gpu.evaluate(times[8 * 110] + 1.0)
  with pytest.raises(StatusError):
    gpu.evaluate(times[0] - 1.0)


def test_max_segments_is_enforced():
  gpu = ContinuousTrajectories(4, 1.0, 1e-3, max_segments=1)
  q = np.zeros((3, 4))
  for k in range(16):
    q[0] = k
    gpu.append(float(k), q, np.ones((3, 4)))
  assert gpu.segment_count() == 1
  with pytest.raises(StatusError) as info:
    gpu.append(16.0, q, q)
  assert info.value.code == _capi.PB_RESOURCE_EXHAUSTED


def test_trappist_ensemble_produces_the_oracles_trajectories():
  """The massive-body flow feeding the fit without leaving the device
  (pb_nbody_ensemble_solve_to_trajectories) against a second ensemble stepped
  one state at a time whose states go through the oracle's
  ContinuousTrajectory."""
  from test_gpu_nbody import ensemble_state
  system = SolarSystem.from_fixture('trappist_jd2457000')
  gpu = Ephemeris(system)
  order, _ = gpu.internal_order()
  n_systems, n_bodies, h = 48, len(order), 1800.0
  method = _capi.QUINLAN_1999_ORDER_8A
  state0 = ensemble_state(system, order, n_systems, 1e-6)
  t0 = system.epoch
  n_steps = 8 * 12 + 3
  n = n_bodies * n_systems
  trajectories = ContinuousTrajectories(n, h, 1e-3, max_segments=16)
  ensemble = NBodyEnsemble(gpu, method, h, state0, t0)
  assert ensemble.solve_to_trajectories(t0 + n_steps * h + 1.0,
                                        trajectories) == n_steps
  assert trajectories.segment_count() == 12
  # Restartable.
  assert ensemble.solve_to_trajectories(t0 + (n_steps + 5) * h + 1.0,
                                        trajectories) == 5
  assert trajectories.segment_count() == 13

  stepwise = NBodyEnsemble(gpu, method, h, state0, t0)
  picked = [0, 1, n_systems + 5, 3 * n_systems + 17, n - 1]
  oracles = {i: OracleTrajectory(h, 1e-3) for i in picked}
  for k in range(n_steps + 5 + 1):
    if k > 0:
      assert stepwise.solve(t0 + k * h + 1.0) == 1
    state, time = stepwise.state()
    flat = state.reshape(12, n)
    for i in picked:
      assert oracles[i].append(time[0], flat[0:3, i], flat[6:9, i]) == 0
  compare(trajectories, oracles, picked,
          [t0, t0 + 12345.6, t0 + 8 * h, t0 + 104 * h])


def test_full_size_fit_reproduces_the_appended_points():
  """Size-independent property at the size of BASELINE.json config 4 (8 bodies
  of 10^5 systems = 8 10^5 trajectories): where the fit settled below the
  maximal degree, the polynomials reproduce the appended positions and
  velocities to within the tolerance (times a margin for the unfitted modes),
  at the nodes and between them."""
  n, step, tolerance = 800000, 1800.0, 1e-3
  rng = np.random.Generator(np.random.MT19937(5))
  period = step * 10.0**rng.uniform(1.5, 4.0, n)
  radius = 10.0**rng.uniform(6, 10, n)
  phase = rng.uniform(0, 2 * np.pi, n)
  omega = 2 * np.pi / period

  def state(t):
    angle = phase + omega * t
    q = np.stack([radius * np.cos(angle), radius * np.sin(angle),
                  0.1 * radius * np.cos(angle)])
    v = np.stack([-radius * omega * np.sin(angle),
                  radius * omega * np.cos(angle),
                  -0.1 * radius * omega * np.sin(angle)])
    return q, v

  gpu = ContinuousTrajectories(n, step, tolerance, max_segments=3)
  apocalypse = False
  for k in range(17):
    try:
      gpu.append(k * step, *state(k * step))
    except StatusError:
      apocalypse = True
  assert gpu.segment_count() == 2
  settled = np.ones(n, dtype=bool)
  degrees = np.zeros((2, n), dtype=np.int32)
  # The degrees, through the evaluation of a few whole trajectories.
  for i in range(0, n, n // 64):
    degrees[:, i] = gpu.segments(i)['degree']
  for t in (0.0, 4 * step, 8 * step, 11.5 * step, 16 * step):
    q, v = gpu.evaluate(t)
    expected_q, expected_v = state(t)
    q_error = np.abs(q - expected_q).max(axis=0)
    v_error = np.abs(v - expected_v).max(axis=0)
    # Slow motions are fitted to the tolerance; the fast ones (period below
    # ~60 steps) exhaust the degrees and carry their error estimate instead.
    slow = period > 200 * step
    assert q_error[slow].max() < 10 * tolerance, (t, q_error[slow].max())
    assert v_error[slow].max() < 10 * tolerance / step, (t,
                                                         v_error[slow].max())
    assert np.all(np.isfinite(q)) and np.all(np.isfinite(v))
  sampled = degrees[:, ::n // 64]
  assert sampled.min() >= 3 and sampled.max() <= 17


def test_sol_ephemeris_produced_on_the_device(sol_system):
  """Ephemeris::Prolong without the host: the Sol system (one system, the
  cooperative single-system kernel) integrated with QuinlanTremaine1990Order12
  at 10 min over three days, every state fitted on the device; the polynomials
  of every body are the oracle's ContinuousTrajectory's, bit for bit."""
  from oracle_lib import OracleEphemeris
  system = sol_system
  gpu = Ephemeris(system)
  order, _ = gpu.internal_order()
  n, h = len(order), 600.0
  state = np.zeros((12, n, 1))
  state[0:3, :, 0] = system.initial_positions()[order].T
  state[6:9, :, 0] = system.initial_velocities()[order].T
  t_final = 3 * 86400.0
  trajectories = ContinuousTrajectories(n, h, 1e-3, 64)
  ensemble = NBodyEnsemble(gpu, _capi.QUINLAN_TREMAINE_1990_ORDER_12, h, state,
                           0.0)
  t = t_final
  while trajectories.t_max() < t_final:  # Ephemeris::Prolong.
    ensemble.solve_to_trajectories(t, trajectories)
    t += h
  oracle = OracleEphemeris(system)
  assert oracle.prolong(t_final) == 0
  assert trajectories.t_max() == oracle.t_max()
  for internal, caller in enumerate(order):
    actual = trajectories.segments(internal)
    expected = oracle.segments(int(caller))
    assert actual['t_min'] == expected['t_min'] == 0.0
    assert_bitwise_equal(actual['t_max'], expected['t_max'])
    assert_bitwise_equal(actual['origin'], expected['origin'])
    assert list(actual['degree']) == list(expected['degree'])
    valid = np.arange(18)[None, :, None] <= actual['degree'][:, None, None]
    assert_bitwise_equal(np.where(valid, actual['coefficients'], 0.0),
                         np.where(valid, expected['coefficients'], 0.0),
                         system.bodies[int(caller)]['name'])
  q, v = trajectories.evaluate(123456.7)
  assert_bitwise_equal(q, oracle.evaluate_all_positions(123456.7)[order].T)
  assert_bitwise_equal(v, oracle.evaluate_all_velocities(123456.7)[order].T)
