"""Concurrency, cancellation and lifetime of the C ABI's handles.

The reference guarantees that flows on distinct objects run concurrently and
that the const queries of an Ephemeris are thread-safe
(physics/ephemeris.hpp:64-66), and its integrators return after every step if
the thread was asked to stop (RETURN_IF_STOPPED,
integrators/symplectic_runge_kutta_nyström_integrator_body.hpp:137,
embedded_explicit_runge_kutta_nyström_integrator_body.hpp:243)."""
import ctypes
import threading
import time

import numpy as np
import pytest

from principia_b200 import _capi, workloads
from principia_b200.ephemeris import Ephemeris

from conftest import assert_bitwise_equal

pytestmark = pytest.mark.gpu

SRKN = _capi.BLANES_MOAN_2002_SRKN_14A


def leo(system, n, seed):
  earth = system.bodies[system.index('Earth')]
  return workloads.leo_probes(n, earth['q'], earth['v'], earth['mu'], seed)


def eccentric(system, n, seed):
  earth = system.bodies[system.index('Earth')]
  return workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                    seed)


def run_threads(targets):
  errors = []

  def guarded(f):
    def run():
      try:
        f()
      except BaseException as error:  # noqa: B902
        errors.append(error)
    return run
  threads = [threading.Thread(target=guarded(f)) for f in targets]
  for t in threads:
    t.start()
  for t in threads:
    t.join()
  if errors:
    raise errors[0]


def test_concurrent_flows_and_queries_on_one_ephemeris(sol_system, sol_gpu,
                                                        sol_oracle):
  """Four threads on ONE ephemeris at once: two fixed-step flows of different
  batches, an adaptive flow, and a stream of acceleration / potential /
  position queries; every result is bit-identical to the one the same call
  gives alone."""
  n = 20000
  batches = [leo(sol_system, n, 1), leo(sol_system, n, 2)]
  adaptive0 = eccentric(sol_system, 3000, 3)
  points = leo(sol_system, 5000, 4)[0:3]

  def fixed(state):
    t = np.zeros(2)
    result = sol_gpu.flow_with_fixed_step(SRKN, 10.0, 3000.0, state, t)
    assert result['n_steps'] == 300 and not result['cancelled']
    return state

  def adaptive(state):
    times = np.zeros((2, state.shape[1]))
    result = sol_gpu.flow_with_adaptive_step(8000.0, state, times, 1e-3, 1e-3)
    assert np.all(result['status'] == 0)
    return state

  def queries():
    a, _ = sol_gpu.compute_gravitational_acceleration_on_massless_bodies(
        1234.5, points)
    u = sol_gpu.compute_gravitational_potential(1234.5, points)
    p = sol_gpu.evaluate_all_positions(777.0)
    return a, u, p

  alone = [fixed(batches[0].copy()), fixed(batches[1].copy()),
           adaptive(adaptive0.copy()), queries()]
  for repeat in range(3):
    together = [None] * 4
    query_results = []

    def many_queries():
      for _ in range(20):
        query_results.append(queries())
      together[3] = query_results[-1]
    run_threads([
        lambda: together.__setitem__(0, fixed(batches[0].copy())),
        lambda: together.__setitem__(1, fixed(batches[1].copy())),
        lambda: together.__setitem__(2, adaptive(adaptive0.copy())),
        many_queries])
    for k in range(3):
      assert_bitwise_equal(together[k], alone[k], 'flow %d' % k)
    for a, u, p in query_results:
      assert_bitwise_equal(a, alone[3][0])
      assert_bitwise_equal(u, alone[3][1])
      assert_bitwise_equal(p, alone[3][2])
  # And the oracle agrees with what ran concurrently.
  expected = batches[0][:, :64].copy()
  sol_oracle.flow_with_fixed_step(SRKN, 10.0, 3000.0, expected, np.zeros(2))
  assert_bitwise_equal(together[0][:, :64], expected)


def test_two_ephemerides_flow_concurrently(sol_system):
  """Distinct ephemerides with different hot bodies share the constant bank by
  turns (the second flow waits for the first); both results are those of the
  calls run alone."""
  a = Ephemeris(sol_system)
  b = Ephemeris(sol_system)
  workloads.upload_sol_ephemeris(a)
  workloads.upload_sol_ephemeris(b)
  near_earth = leo(sol_system, 30000, 5)
  moon = sol_system.bodies[sol_system.index('Moon')]
  earth = sol_system.bodies[sol_system.index('Earth')]
  # Low lunar orbits: the Moon's degree-30 field is not a hot-body candidate
  # (degree > 10), so `b` flows with no hot body: other constants than `a`'s.
  near_moon = workloads.leo_probes(2000, moon['q'], moon['v'], moon['mu'], 6)
  near_moon[0:3] = (np.array(moon['q'])[:, None] +
                    (near_moon[0:3] - np.array(moon['q'])[:, None]) * 0.3)

  def flow(ephemeris, state, steps):
    result = ephemeris.flow_with_fixed_step(SRKN, 10.0, 10.0 * steps, state,
                                            np.zeros(2))
    assert result['n_steps'] == steps
    return state

  alone = [flow(a, near_earth.copy(), 200), flow(b, near_moon.copy(), 40)]
  assert a.last_flow_path()[0] == sol_system.index('Earth')
  assert b.last_flow_path()[0] != sol_system.index('Earth')
  together = [None, None]
  run_threads([
      lambda: together.__setitem__(0, flow(a, near_earth.copy(), 200)),
      lambda: together.__setitem__(1, flow(b, near_moon.copy(), 40))])
  assert_bitwise_equal(together[0], alone[0])
  assert_bitwise_equal(together[1], alone[1])
  a.close()
  b.close()


def test_prolong_while_flowing(sol_system):
  """pb_ephemeris_append_segments (Ephemeris::Prolong) from one thread while
  another flows: the append waits for the flow in progress, the next flow sees
  the longer ephemeris; only the appended segments are uploaded."""
  segments = workloads.load_sol_ephemeris_segments()
  half = len(segments[0]['t_max']) // 2

  def part(s, lo, hi):
    return dict(t_min=s['t_min'], t_max=s['t_max'][lo:hi],
                origin=s['origin'][lo:hi], degree=s['degree'][lo:hi],
                coefficients=s['coefficients'][lo:hi])

  gpu = Ephemeris(sol_system)
  for body, s in enumerate(segments):
    gpu.append_segments(body, part(s, 0, half))
  t_half = gpu.t_max()
  state = leo(sol_system, 20000, 9)
  reference = Ephemeris(sol_system)
  workloads.upload_sol_ephemeris(reference)

  def first_flow():
    s = state.copy()
    gpu.flow_with_fixed_step(SRKN, 10.0, 4000.0, s, np.zeros(2))
    return s

  def prolong():
    for body, s in enumerate(segments):
      gpu.append_segments(body, part(s, half, None))
  results = [None]
  run_threads([lambda: results.__setitem__(0, first_flow()), prolong])
  assert gpu.t_max() == reference.t_max() > t_half
  expected = state.copy()
  reference.flow_with_fixed_step(SRKN, 10.0, 4000.0, expected, np.zeros(2))
  assert_bitwise_equal(results[0], expected)
  # Beyond the first half: needs the appended segments.
  t0 = np.array([t_half - 500.0, 0.0])
  late, late_expected = state.copy(), state.copy()
  gpu.flow_with_fixed_step(SRKN, 10.0, t_half + 2000.0, late, t0.copy())
  reference.flow_with_fixed_step(SRKN, 10.0, t_half + 2000.0, late_expected,
                                 t0.copy())
  assert_bitwise_equal(late, late_expected)
  gpu.close()
  reference.close()


def test_fixed_step_flow_is_cancelled_at_a_chunk_boundary(sol_system, sol_gpu,
                                                          sol_oracle):
  """RETURN_IF_STOPPED: a stop request from another thread makes a long flow
  return PB_CANCELLED; the state and the time are those of a whole number of
  steps (a multiple of the 256-step chunk), equal to the oracle's after as
  many steps, and the flow can be resumed."""
  n = 30000
  state = leo(sol_system, n, 11)
  initial = state.copy()
  t = np.zeros(2)
  outcome = {}
  sol_gpu.flow_with_fixed_step(SRKN, 10.0, 100.0, state.copy(), np.zeros(2))

  def flow():
    outcome.update(sol_gpu.flow_with_fixed_step(SRKN, 10.0, 150000.0, state, t))

  thread = threading.Thread(target=flow)
  thread.start()
  time.sleep(0.5)
  sol_gpu.request_stop()
  thread.join()
  try:
    assert outcome['cancelled']
    steps = outcome['n_steps']
    assert 0 < steps < 15000 and steps % 256 == 0
    assert t[0] == 10.0 * steps
    expected = initial[:, :32].copy()
    sol_oracle.flow_with_fixed_step(SRKN, 10.0, 10.0 * steps, expected,
                                    np.zeros(2))
    assert_bitwise_equal(state[:, :32], expected, 'state at cancellation')
    # While the request stands, flows stop before their first step.
    again = sol_gpu.flow_with_fixed_step(SRKN, 10.0, t[0] + 100.0,
                                         state.copy(), t.copy())
    assert again['cancelled'] and again['n_steps'] == 0
  finally:
    sol_gpu.clear_stop()
  # Resumed.
  result = sol_gpu.flow_with_fixed_step(SRKN, 10.0, t[0] + 100.0, state, t)
  assert not result['cancelled'] and result['n_steps'] == 10
  sol_oracle.flow_with_fixed_step(SRKN, 10.0, 10.0 * (steps + 10), expected,
                                  np.array([10.0 * steps, 0.0]))
  assert_bitwise_equal(state[:, :32], expected, 'resumed state')


def test_adaptive_flow_is_cancelled_after_a_step(sol_system, sol_gpu,
                                                 sol_oracle):
  """The adaptive kernel polls the stop word after every accepted step: the
  trajectories that were cut short report PB_CANCELLED with the state of their
  last accepted step -- the state the oracle reaches with max_steps set to the
  number of steps they took."""
  n = 60000
  state = eccentric(sol_system, n, 13)
  initial = state.copy()
  times = np.zeros((2, n))
  outcome = {}
  sol_gpu.flow_with_adaptive_step(100.0, state[:, :256].copy(),
                                  np.zeros((2, 256)), 1e-3, 1e-3)

  def flow():
    outcome.update(sol_gpu.flow_with_adaptive_step(
        170000.0, state, times, 1e-3, 1e-3, step_log_capacity=0))

  thread = threading.Thread(target=flow)
  thread.start()
  time.sleep(0.4)
  sol_gpu.request_stop()
  thread.join()
  sol_gpu.clear_stop()
  assert outcome['cancelled']
  status = outcome['status']
  assert set(np.unique(status)) <= {_capi.PB_OK, _capi.PB_CANCELLED}
  cut = np.flatnonzero(status == _capi.PB_CANCELLED)
  assert len(cut) > 0
  started = cut[outcome['accepted'][cut] > 0]
  untouched = cut[outcome['accepted'][cut] == 0]
  assert_bitwise_equal(state[:, untouched], initial[:, untouched])
  assert np.all(times[0, untouched] == 0.0)
  for i in started[:12]:
    expected = np.ascontiguousarray(initial[:, i:i + 1])
    expected_time = np.zeros((2, 1))
    result = sol_oracle.flow_with_adaptive_step(
        170000.0, expected, expected_time, 1e-3, 1e-3,
        max_steps=int(outcome['accepted'][i]))
    assert result['accepted'][0] == outcome['accepted'][i]
    assert_bitwise_equal(state[:, i:i + 1], expected, 'probe %d' % i)
    assert_bitwise_equal(times[:, i:i + 1], expected_time)


def test_ephemeris_outlives_its_instances(sol_system):
  """pb_ephemeris_destroy refuses while an instance is alive (a destroyed
  ephemeris under a live instance was a use-after-free)."""
  lib = _capi.library()
  gpu = Ephemeris(sol_system)
  workloads.upload_sol_ephemeris(gpu)
  state = leo(sol_system, 100, 17)
  instance = gpu.new_instance(_capi.QUINLAN_1999_ORDER_8A, 10.0, state,
                              np.zeros(2))
  handle = gpu.handle
  lib.pb_ephemeris_destroy(handle)
  assert b'still alive' in lib.pb_last_error()
  # Still usable.
  assert instance.flow(200.0)['n_steps'] == 20
  lib.pb_fixed_step_instance_destroy(instance.handle)
  instance.handle = None
  gpu.close()


def test_argument_validation_of_the_flows(sol_system, sol_gpu):
  n = 8
  state = leo(sol_system, n, 19)
  with pytest.raises(Exception) as info:
    sol_gpu.flow_with_adaptive_step(100.0, state.copy(), np.zeros((2, n)),
                                    1e-3, 1e-3, max_steps=0)
  assert info.value.code == _capi.PB_INVALID_ARGUMENT
  with pytest.raises(Exception) as info:
    sol_gpu.flow_with_adaptive_step(100.0, state.copy(), np.zeros((2, n)),
                                    -1.0, 1e-3)
  assert info.value.code == _capi.PB_INVALID_ARGUMENT
  # A trajectory that starts before t_min: the reference CHECKs; here its
  # status says so and the others flow.
  times = np.zeros((2, n))
  times[0, 3] = -1000.0
  s = state.copy()
  result = sol_gpu.flow_with_adaptive_step(100.0, s, times, 1e-3, 1e-3)
  assert result['status'][3] == _capi.PB_INVALID_ARGUMENT
  assert np.all(np.delete(result['status'], 3) == 0)
  assert_bitwise_equal(s[:, 3], state[:, 3])
  with pytest.raises(Exception) as info:
    sol_gpu.flow_with_fixed_step(SRKN, 10.0, float('nan'), state.copy(),
                                 np.zeros(2))
  assert info.value.code == _capi.PB_INVALID_ARGUMENT


def test_hot_body_is_chosen_by_all_the_probes(sol_system):
  """The probes of a batch vote for the hot and warm bodies: a first probe
  that is nowhere near the others does not send the batch down the generic
  path."""
  gpu = Ephemeris(sol_system)
  workloads.upload_sol_ephemeris(gpu)
  earth = sol_system.bodies[sol_system.index('Earth')]
  state = leo(sol_system, 4096, 23)
  far = workloads.eccentric_probes(64, earth['q'], earth['v'], earth['mu'], 5)
  distance = np.linalg.norm(far[0:3] - np.array(earth['q'])[:, None], axis=0)
  state[:, 0] = far[:, int(np.argmax(distance))]
  assert distance.max() > 1.0e8  # Beyond every harmonic of the Earth but J2.
  gpu.flow_with_fixed_step(SRKN, 10.0, 100.0, state, np.zeros(2))
  assert gpu.last_flow_path() == (sol_system.index('Earth'), False)
  gpu.close()


def test_hot_body_choice_is_reported_and_revised(sol_system):
  """The hot body is chosen by the probes of the first flow and kept;
  a later batch around another body takes the generic harmonics, which the
  flow reports, and the next flow chooses afresh.  Same bits either way."""
  gpu = Ephemeris(sol_system)
  workloads.upload_sol_ephemeris(gpu)
  earth_index = sol_system.index('Earth')
  mars = sol_system.bodies[sol_system.index('Mars')]
  near_earth = leo(sol_system, 4096, 21)
  near_mars = workloads.leo_probes(4096, mars['q'], mars['v'], mars['mu'], 22)
  near_mars[0:3] = (np.array(mars['q'])[:, None] +
                    (near_mars[0:3] - np.array(mars['q'])[:, None]) * 0.55)

  def flow(state):
    s = state.copy()
    gpu.flow_with_fixed_step(SRKN, 10.0, 100.0, s, np.zeros(2))
    return s

  flow(near_earth)
  assert gpu.last_flow_path() == (earth_index, False)
  slow = flow(near_mars)  # Earth is still hot: Mars' field is generic.
  hot, generic = gpu.last_flow_path()
  assert hot == earth_index and generic
  fast = flow(near_mars)
  hot, generic = gpu.last_flow_path()
  assert hot == sol_system.index('Mars') and not generic
  assert_bitwise_equal(fast, slow)
  gpu.close()
