"""PB_ARITHMETIC_CONTRACTED: the opt-in arithmetic of the fixed-step flows and
of the all-pairs force (fused multiply-adds; r^-3 from a refined reciprocal
square root in the all-pairs kernel).  Not bit-identical to the reference --
its build forbids contraction (Makefile:123-169 there) -- but within the
tolerance BASELINE.json states for this path: relative position error <= 1e-12
after 10^4 fixed steps, energy drift matching the reference's.  The default
arithmetic stays exact (every other GPU test compares bits)."""
import numpy as np
import pytest

from principia_b200 import _capi, workloads
from principia_b200.nbody import NBodyAllPairs

from conftest import assert_bitwise_equal

pytestmark = pytest.mark.gpu

SRKN = _capi.BLANES_MOAN_2002_SRKN_14A
CONTRACTED = _capi.PB_ARITHMETIC_CONTRACTED


def leo(system, n, seed):
  earth = system.bodies[system.index('Earth')]
  return workloads.leo_probes(n, earth['q'], earth['v'], earth['mu'], seed)


def earth_relative_energy(oracle, system, state, t):
  earth = system.index('Earth')
  q_earth = oracle.evaluate_all_positions(t)[earth]
  v_earth = oracle.evaluate_all_velocities(t)[earth]
  q = state[0:3] - q_earth[:, None]
  v = state[6:9] - v_earth[:, None]
  mu = system.bodies[earth]['mu']
  return 0.5 * (v * v).sum(axis=0) - mu / np.linalg.norm(q, axis=0)


def test_contracted_srkn_flow_is_within_tolerance(sol_system, sol_oracle,
                                                  sol_gpu):
  """10^3 LEO probes, 10^4 BlanesMoan2002SRKN14A steps of 10 s (17 orbits):
  the contracted flow against the CPU oracle (the reference's arithmetic)."""
  n, steps, h = 1000, 10000, 10.0
  state0 = leo(sol_system, n, 31)
  contracted = state0.copy()
  t = np.zeros(2)
  result = sol_gpu.flow_with_fixed_step(SRKN, h, steps * h, contracted, t,
                                        arithmetic=CONTRACTED)
  assert result['n_steps'] == steps and result['status'] == 0
  exact = state0.copy()
  sol_gpu.flow_with_fixed_step(SRKN, h, steps * h, exact, np.zeros(2))
  expected = state0.copy()
  expected_time = np.zeros(2)
  sol_oracle.flow_with_fixed_step(SRKN, h, steps * h, expected, expected_time)
  assert_bitwise_equal(exact, expected, 'the default arithmetic stays exact')
  assert_bitwise_equal(t, expected_time)
  # The tolerance of BASELINE.json: relative error of the position (the state
  # is barycentric: |q| is 1 au).
  dq = np.linalg.norm(contracted[0:3] - expected[0:3], axis=0)
  relative = dq / np.linalg.norm(expected[0:3], axis=0)
  assert 0 < relative.max() <= 1e-12, relative.max()
  # A sharper statement: relative to the orbital radius (7 10^6 m) the error
  # after 17 orbits is that of two differently rounded evaluations of the same
  # force (a few 10^-10, along the track).
  earth = sol_oracle.evaluate_all_positions(steps * h)[
      sol_system.index('Earth')]
  radius = np.linalg.norm(expected[0:3] - earth[:, None], axis=0)
  assert (dq / radius).max() < 2e-8, (dq / radius).max()
  dv = np.linalg.norm(contracted[6:9] - expected[6:9], axis=0)
  assert (dv / np.linalg.norm(expected[6:9], axis=0)).max() <= 1e-9
  # Energy drift (relative to the Earth): the contracted flow drifts like the
  # reference's arithmetic does.
  e0 = earth_relative_energy(sol_oracle, sol_system, state0, 0.0)
  drift_reference = earth_relative_energy(sol_oracle, sol_system, expected,
                                          steps * h) - e0
  drift_contracted = earth_relative_energy(sol_oracle, sol_system, contracted,
                                           steps * h) - e0
  assert np.abs(drift_contracted - drift_reference).max() <= (
      1e-9 * np.abs(e0).max())
  # (The two-body energy itself oscillates by the Earth's oblateness: 10^-3.)
  assert np.abs(drift_contracted / e0).max() < 1e-2


@pytest.mark.parametrize('method', [_capi.QUINLAN_1999_ORDER_8A,
                                    _capi.QUINLAN_TREMAINE_1990_ORDER_12])
def test_contracted_multistep_instance_is_within_tolerance(sol_system,
                                                           sol_oracle, sol_gpu,
                                                           method):
  """The history integrator with contraction: startup (SRKN14A at step / 16),
  FillStepFromState and the multistep steps all run the contracted kernels."""
  n, steps, h = 500, 10000, 10.0
  state0 = leo(sol_system, n, 37)
  instance = sol_gpu.new_instance(method, h, state0, np.zeros(2))
  instance.set_arithmetic(CONTRACTED)
  assert instance.flow(steps * h)['n_steps'] == steps
  contracted, time = instance.state()
  expected = state0.copy()
  expected_time = np.zeros(2)
  sol_oracle.flow_with_fixed_step(method, h, steps * h, expected,
                                  expected_time)
  assert_bitwise_equal(time, expected_time)
  dq = np.linalg.norm(contracted[0:3] - expected[0:3], axis=0)
  relative = dq / np.linalg.norm(expected[0:3], axis=0)
  assert 0 < relative.max() <= 1e-12, relative.max()


def test_contracted_all_pairs_force_and_flow():
  """The all-pairs force with contraction: accelerations within 1e-13 of the
  exact ones, positions after a QuinlanTremaine1990Order12 flow within 1e-12."""
  from test_gpu_nbody import random_cluster
  n = 3000
  mu, q, v = random_cluster(n, seed=41)
  method = _capi.QUINLAN_TREMAINE_1990_ORDER_12
  exact = NBodyAllPairs(method, 3600.0, mu, q, v, 0.0)
  contracted = NBodyAllPairs(method, 3600.0, mu, q, v, 0.0)
  contracted.set_arithmetic(CONTRACTED)
  a_exact = exact.accelerations()
  a_contracted = contracted.accelerations()
  scale = np.linalg.norm(a_exact, axis=0)
  error = np.linalg.norm(a_contracted - a_exact, axis=0) / scale
  assert 0 < error.max() < 1e-13, error.max()
  t_final = 14 * 3600.0
  assert exact.solve(t_final)[0] == contracted.solve(t_final)[0] == 14
  q_exact, v_exact, _ = exact.state()
  q_contracted, v_contracted, _ = contracted.state()
  au = 149597870700.0
  assert np.abs(q_contracted - q_exact).max() / au < 1e-12
  exact.close()
  contracted.close()
