"""GPU parity of the trajectory sink:
DiscreteTrajectorySegment::Append with downsampling on the device against the
oracle, bit for bit, on the reference's own test trajectories and on the states
appended by the flows."""
import numpy as np
import pytest

from principia_b200 import _capi
from principia_b200 import workloads
from principia_b200.trajectory import Downsampler

import oracle_lib
import trajectory_factories
from conftest import assert_bitwise_equal
from test_oracle_kats import oracle_downsample

pytestmark = pytest.mark.gpu


def test_reference_circle_and_straight_line():
  """physics/discrete_trajectory_segment_test.cpp:252-324: 1001 points on a
  circle are downsampled to 49, on a straight line to 2 (tolerance 1 mm); both
  trajectories go through one batch."""
  circle = trajectory_factories.circular_timeline(3.0, 2.0, 10e-3, 0.0, 10.0)
  line = trajectory_factories.linear_timeline([1.0, 2.0, 3.0], 10e-3, 0.0,
                                              10.0)
  assert_bitwise_equal(circle[0], line[0])  # The same times.
  times = circle[0]
  samples = np.zeros((len(times), 6, 2))
  for k, (_, q, v) in enumerate((circle, line)):
    samples[:, 0:3, k] = q
    samples[:, 3:6, k] = v
  downsampler = Downsampler(2, 1e-3, 64)
  # In three calls: the state of the segments persists on the device.
  for begin, end in ((0, 1), (1, 400), (400, len(times))):
    downsampler.append(times[begin:end], samples[begin:end])
  counts, kept_times, points, errors = downsampler.timelines()
  assert list(counts) == [49, 2]
  for k, (_, q, v) in enumerate((circle, line)):
    expected = oracle_downsample(times, q, v, 1e-3)
    c = counts[k]
    assert_bitwise_equal(kept_times[:c, k], expected['times'])
    assert_bitwise_equal(points[:c, 0:3, k], expected['q'])
    assert_bitwise_equal(points[:c, 3:6, k], expected['v'])
    assert_bitwise_equal(errors[:c, k], expected['errors'])
  # Evaluation: at a kept point, between kept points, at both ends, outside
  # (discrete_trajectory_segment_test.cpp:283-300 evaluates every millisecond).
  evaluation_times = np.concatenate(
      [np.arange(0.0, 10.0, 0.1237), [0.0, kept_times[3, 0], times[-1]]])
  for k, (_, q, v) in enumerate((circle, line)):
    expected = oracle_downsample(times, q, v, 1e-3, evaluation_times)
    for j, t in enumerate(evaluation_times):
      eq, ev, in_range = downsampler.evaluate(float(t))
      assert in_range.all()
      assert_bitwise_equal(eq[:, k], expected['evaluated_q'][j], 'q at %r' % t)
      assert_bitwise_equal(ev[:, k], expected['evaluated_v'][j], 'v at %r' % t)
  eq, ev, in_range = downsampler.evaluate(10.5)
  assert not in_range.any() and np.isnan(eq).all() and np.isnan(ev).all()
  eq, ev, in_range = downsampler.evaluate(-1e-9)
  assert not in_range.any()


def test_flow_samples_are_downsampled_like_the_oracle(sol_system, sol_gpu):
  """The plugin's history: Quinlan1999Order8A at 10 s, downsampled with the
  default tolerance of 10 m (ksp_plugin/integrators.cpp:23-29)."""
  n = 300
  earth = sol_system.bodies[sol_system.index('Earth')]
  state = workloads.leo_probes(n, earth['q'], earth['v'], earth['mu'], 77)
  time = np.zeros(2)
  first = np.concatenate([state[0:3], state[6:9]])[None]
  result = sol_gpu.flow_with_fixed_step(_capi.QUINLAN_1999_ORDER_8A, 10.0,
                                        6000.0, state, time, sample_every=1,
                                        max_samples=600)
  assert len(result['samples']) == 600
  times = np.concatenate([[0.0], result['sample_times']])
  samples = np.concatenate([first, result['samples']])
  downsampler = Downsampler(n, 10.0, 601)
  downsampler.append(times, samples)
  counts, kept_times, points, errors = downsampler.timelines()
  assert 10 < counts.min() and counts.max() < 300
  for k in range(0, n, 7):
    expected = oracle_downsample(
        times, np.ascontiguousarray(samples[:, 0:3, k]),
        np.ascontiguousarray(samples[:, 3:6, k]), 10.0)
    c = counts[k]
    assert c == len(expected['times'])
    assert_bitwise_equal(kept_times[:c, k], expected['times'])
    assert_bitwise_equal(points[:c, 0:3, k], expected['q'])
    assert_bitwise_equal(points[:c, 3:6, k], expected['v'])
    assert_bitwise_equal(errors[:c, k], expected['errors'])


@pytest.mark.parametrize('method', [_capi.QUINLAN_1999_ORDER_8A,
                                    _capi.BLANES_MOAN_2002_SRKN_14A])
def test_instance_flows_into_the_device_sink(sol_system, sol_gpu, method):
  """FlowWithFixedStep with the sink on the device (nothing travels to the
  host) gives the timelines of flow-then-append."""
  n = 200
  earth = sol_system.bodies[sol_system.index('Earth')]
  state = workloads.leo_probes(n, earth['q'], earth['v'], earth['mu'], 78)
  first = np.concatenate([state[0:3], state[6:9]])[None]
  t_final = 5005.0
  # Reference: all the samples through the host.
  instance = sol_gpu.new_instance(method, 10.0, state, np.zeros(2))
  result = instance.flow(t_final, sample_every=1, max_samples=1000)
  assert result['n_steps'] == 500
  expected = Downsampler(n, 10.0, 501)
  expected.append(np.concatenate([[0.0], result['sample_times']]),
                  np.concatenate([first, result['samples']]))
  # Device sink, in two calls (the second one crosses chunk boundaries).
  instance = sol_gpu.new_instance(method, 10.0, state, np.zeros(2))
  sink = Downsampler(n, 10.0, 501)
  sink.append([0.0], first)
  steps = instance.flow_to_downsampler(45.0, sink)['n_steps']
  steps += instance.flow_to_downsampler(t_final, sink)['n_steps']
  assert steps == 500
  counts, times, points, errors = sink.timelines()
  wanted_counts, wanted_times, wanted_points, wanted_errors = (
      expected.timelines())
  assert list(counts) == list(wanted_counts)
  # Slots beyond the size of a timeline are unspecified.
  valid = np.arange(501)[:, None] < counts[None, :]
  assert_bitwise_equal(np.where(valid, times, 0.0),
                       np.where(valid, wanted_times, 0.0), 'times')
  assert_bitwise_equal(np.where(valid, errors, 0.0),
                       np.where(valid, wanted_errors, 0.0), 'errors')
  assert_bitwise_equal(np.where(valid[:, None, :], points, 0.0),
                       np.where(valid[:, None, :], wanted_points, 0.0),
                       'points')
  final_state, final_time = instance.state()
  assert final_time[0] == 5000.0


def test_adaptive_flow_appends_to_the_sink_in_the_kernel(sol_system,
                                                         sol_oracle, sol_gpu):
  """FlowWithAdaptiveStep into a DiscreteTrajectory with the plugin's default
  downsampling: every accepted state goes through Append inside the kernel.
  Expected: the oracle's accepted states (its step log) through the oracle's
  DiscreteTrajectorySegment."""
  n = 64
  earth = sol_system.bodies[sol_system.index('Earth')]
  state = workloads.eccentric_probes(n, earth['q'], earth['v'], earth['mu'],
                                     13)
  first = np.concatenate([state[0:3], state[6:9]])[None]
  t_final = 20000.0
  capacity = 4000
  sink = Downsampler(n, 10.0, capacity)
  sink.append([0.0], first)
  gpu_state, gpu_time = state.copy(), np.zeros((2, n))
  result = sol_gpu.flow_with_adaptive_step(t_final, gpu_state, gpu_time, 1e-3,
                                           1e-3, sink=sink)
  assert np.all(result['status'] == 0)
  counts, times, points, errors = sink.timelines()
  # The oracle, one probe at a time, with the states of its accepted steps.
  orc = oracle_lib.library()
  checked = 0
  for k in range(0, n, 5):
    accepted = int(result['accepted'][k])
    expected_state = np.ascontiguousarray(state[:, k:k + 1])
    expected_time = np.zeros((2, 1))
    flow = sol_oracle.flow_with_adaptive_step(
        t_final, expected_state, expected_time, 1e-3, 1e-3,
        step_log_capacity=accepted, state_log=True)
    assert int(flow['accepted'][0]) == accepted
    log_times = np.concatenate([[0.0], flow['step_log'][:accepted, 0]])
    log_q = np.concatenate([first[0, 0:3, k][None],
                            flow['state_log'][:accepted, 0:3, 0]])
    log_v = np.concatenate([first[0, 3:6, k][None],
                            flow['state_log'][:accepted, 3:6, 0]])
    expected = oracle_downsample(log_times, log_q, log_v, 10.0)
    c = counts[k]
    assert c == len(expected['times']) < accepted + 1
    assert_bitwise_equal(times[:c, k], expected['times'])
    assert_bitwise_equal(points[:c, 0:3, k], expected['q'])
    assert_bitwise_equal(points[:c, 3:6, k], expected['v'])
    assert_bitwise_equal(errors[:c, k], expected['errors'])
    checked += 1
  assert checked == 13


def test_capacity_overflow_is_reported():
  circle = trajectory_factories.circular_timeline(3.0, 2.0, 10e-3, 0.0, 10.0)
  times, q, v = circle
  samples = np.zeros((len(times), 6, 1))
  samples[:, 0:3, 0] = q
  samples[:, 3:6, 0] = v
  downsampler = Downsampler(1, 1e-3, 20)
  downsampler.append(times, samples)
  from principia_b200.ephemeris import StatusError
  with pytest.raises(StatusError):
    downsampler.timelines()
  # No downsampling: every point is kept.
  downsampler = Downsampler(1, -1.0, 1001)
  downsampler.append(times, samples)
  counts, kept_times, _, _ = downsampler.timelines()
  assert counts[0] == 1001
  assert_bitwise_equal(kept_times[:, 0], times)


def test_reference_evaluate_on_the_device():
  """physics/discrete_trajectory_segment_test.cpp:224-250 (Evaluate) through
  pb_downsampler_evaluate: 1001 points on a circle, no downsampling, evaluated
  every millisecond: errors 4.2(1) nm and 10.4(1) nm/s, and the oracle's bits."""
  from test_oracle_kats import is_near
  times, q, v = trajectory_factories.circular_timeline(3.0, 2.0, 10e-3, 0.0,
                                                       10.0)
  samples = np.zeros((len(times), 6, 1))
  samples[:, 0:3, 0] = q
  samples[:, 3:6, 0] = v
  downsampler = Downsampler(1, -1.0, 1001)
  downsampler.append(times, samples)
  counts, _, _, _ = downsampler.timelines()
  assert list(counts) == [1001]
  evaluation_times = []
  t = times[0]
  while t <= times[-1]:
    evaluation_times.append(t)
    t += 1e-3
  expected = oracle_downsample(times, q, v, -1.0, np.array(evaluation_times))
  position_error = velocity_error = 0.0
  for j, t in enumerate(evaluation_times):
    eq, ev, in_range = downsampler.evaluate(t)
    assert in_range.all()
    if j % 7 == 0:
      assert_bitwise_equal(eq[:, 0], expected['evaluated_q'][j])
      assert_bitwise_equal(ev[:, 0], expected['evaluated_v'][j])
    position_error = max(position_error,
                         abs(np.linalg.norm(eq[:, 0]) - 2.0))
    velocity_error = max(velocity_error,
                         abs(np.linalg.norm(ev[:, 0]) - 6.0))
  assert is_near(position_error, 4.2e-9)
  assert is_near(velocity_error, 10.4e-9, digits=3)
