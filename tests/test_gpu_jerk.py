"""GPU parity tests of the jerk and Jacobian queries,
through the C ABI, against the CPU oracle: bit-exact."""
import numpy as np
import pytest

from principia_b200 import workloads

from conftest import assert_bitwise_equal

pytestmark = pytest.mark.gpu

TIMES = [0.0, 1234.5, 86400.0, 170000.0, 172800.0]


def test_velocities_of_massive_bodies(sol_oracle, sol_gpu):
  for t in TIMES:
    assert_bitwise_equal(sol_gpu.evaluate_all_velocities(t),
                         sol_oracle.evaluate_all_velocities(t), 't=%r' % t)


def test_jerk_on_massless_bodies(sol_system, sol_oracle, sol_gpu):
  n = 20000
  earth = sol_system.index('Earth')
  body = sol_system.bodies[earth]
  state = workloads.leo_probes(n, body['q'], body['v'], body['mu'], seed=42)
  q = np.ascontiguousarray(state[0:3])
  v = np.ascontiguousarray(state[6:9])
  for t in TIMES[:3]:
    assert_bitwise_equal(
        sol_gpu.compute_gravitational_jerk_on_massless_bodies(t, q, v),
        sol_oracle.jerk_on_massless_bodies(t, q, v), 't=%r' % t)
  # Probes anywhere in the system, not only around the Earth (the shape of
  # ephemeris_test.cpp:915-989: within 10^9 m and 10^3 m/s of the Earth).
  rng = np.random.Generator(np.random.MT19937(42))
  q = (sol_oracle.evaluate_all_positions(1234.5)[earth][:, None] +
       rng.uniform(-1e9, 1e9, (3, 5000)))
  v = (sol_oracle.evaluate_all_velocities(1234.5)[earth][:, None] +
       rng.uniform(-1e3, 1e3, (3, 5000)))
  assert_bitwise_equal(
      sol_gpu.compute_gravitational_jerk_on_massless_bodies(1234.5, q, v),
      sol_oracle.jerk_on_massless_bodies(1234.5, q, v))


def test_jerk_on_massless_bodies_empty_and_ragged(sol_oracle, sol_gpu):
  empty = np.zeros((3, 0))
  assert sol_gpu.compute_gravitational_jerk_on_massless_bodies(
      0.0, empty, empty).shape == (3, 0)
  rng = np.random.Generator(np.random.MT19937(7))
  for n in (1, 31, 129):
    q = rng.uniform(-1e11, 1e11, (3, n))
    v = rng.uniform(-3e4, 3e4, (3, n))
    assert_bitwise_equal(
        sol_gpu.compute_gravitational_jerk_on_massless_bodies(100.0, q, v),
        sol_oracle.jerk_on_massless_bodies(100.0, q, v), 'n=%d' % n)


def test_jerk_and_jacobian_on_massive_bodies(sol_oracle, sol_gpu):
  for t in TIMES:
    assert_bitwise_equal(
        sol_gpu.compute_gravitational_jerk_on_massive_bodies(t),
        sol_oracle.jerk_on_massive_bodies(t), 'jerk t=%r' % t)
    jacobians = sol_gpu.compute_jacobian_on_massive_bodies(t)
    assert_bitwise_equal(jacobians, sol_oracle.jacobian_on_massive_bodies(t),
                         'jacobian t=%r' % t)
    assert np.array_equal(jacobians, jacobians.transpose(0, 2, 1))


def test_jerk_on_massive_bodies_with_given_degrees_of_freedom(
    sol_system, sol_oracle, sol_gpu):
  # ComputeGravitationalJerkOnMassiveBodies(bodies, degrees of freedom).
  rng = np.random.Generator(np.random.MT19937(3))
  q = sol_system.initial_positions() + rng.uniform(-1e6, 1e6,
                                                   (len(sol_system.bodies), 3))
  v = sol_system.initial_velocities() + rng.uniform(
      -10, 10, (len(sol_system.bodies), 3))
  assert_bitwise_equal(
      sol_gpu.compute_gravitational_jerk_on_massive_bodies(0.0, q, v),
      sol_oracle.jerk_on_massive_bodies(0.0, q, v))
  assert_bitwise_equal(sol_gpu.compute_jacobian_on_massive_bodies(0.0, q),
                       sol_oracle.jacobian_on_massive_bodies(0.0, q))


def test_times_outside_the_ephemeris_are_refused(sol_gpu):
  from principia_b200.ephemeris import StatusError
  q = np.zeros((3, 1))
  with pytest.raises(StatusError):
    sol_gpu.compute_gravitational_jerk_on_massless_bodies(1e9, q, q)
  with pytest.raises(StatusError):
    sol_gpu.evaluate_all_velocities(-1.0)
