"""GPU parity of the massive-body flows against the CPU oracle."""
import numpy as np
import pytest

from principia_b200 import _capi
from principia_b200.ephemeris import Ephemeris
from principia_b200.nbody import NBodyAllPairs, NBodyEnsemble
from principia_b200.solar_system import SolarSystem

from conftest import assert_bitwise_equal
from oracle_lib import OracleEphemeris

pytestmark = pytest.mark.gpu


def ensemble_state(system, order, n_systems, relative_sigma, seed=42):
  """(12, n_bodies, n_systems) state in internal order: the system's initial
  state, each Cartesian coordinate perturbed by N(0, sigma) relative."""
  rng = np.random.Generator(np.random.MT19937(seed))
  q = system.initial_positions()[order]
  v = system.initial_velocities()[order]
  n_bodies = len(order)
  state = np.zeros((12, n_bodies, n_systems))
  for c in range(3):
    state[c] = q[:, c, None] * (1 + relative_sigma *
                                rng.normal(size=(n_bodies, n_systems)))
    state[6 + c] = v[:, c, None] * (1 + relative_sigma *
                                    rng.normal(size=(n_bodies, n_systems)))
  state[..., 0] = 0.0
  for c in range(3):
    state[c, :, 0] = q[:, c]
    state[6 + c, :, 0] = v[:, c]
  return state


@pytest.mark.parametrize('method,t_days', [
    (_capi.QUINLAN_1999_ORDER_8A, 4.0),
    (_capi.BLANES_MOAN_2002_SRKN_14A, 0.5)])
def test_trappist_ensemble_matches_oracle(method, t_days):
  """BASELINE.json config 4 shape: perturbed copies of TRAPPIST-1 (8 bodies,
  7 with J2), 30 min steps."""
  system = SolarSystem.from_fixture('trappist_jd2457000')
  gpu = Ephemeris(system)
  oracle = OracleEphemeris(system, method=method, step=1800.0)
  order, n_oblate = gpu.internal_order()
  assert n_oblate == 7
  n_systems = 100
  state0 = ensemble_state(system, order, n_systems, 1e-6)
  t0 = system.epoch
  t_final = t0 + t_days * 86400.0
  ensemble = NBodyEnsemble(gpu, method, 1800.0, state0, t0)
  steps = ensemble.solve(t_final)
  state, time = ensemble.state()
  expected_state = state0.copy()
  expected_time = np.array([t0, 0.0])
  expected_steps = oracle.nbody_ensemble_flow(method, 1800.0, expected_state,
                                              expected_time, t_final)
  assert steps == expected_steps == int(t_days * 48)
  assert_bitwise_equal(time, expected_time, 'time')
  assert_bitwise_equal(state, expected_state, 'ensemble state')
  # Restartable: a second Solve continues the same instance.
  steps = ensemble.solve(t_final + 5 * 1800.0 + 10.0)
  oracle_state2 = state0.copy()
  oracle_time2 = np.array([t0, 0.0])
  oracle.nbody_ensemble_flow(method, 1800.0, oracle_state2, oracle_time2,
                             t_final + 5 * 1800.0 + 10.0)
  assert steps == 5
  state, time = ensemble.state()
  assert_bitwise_equal(state, oracle_state2, 'continued state')


def test_sol_system_massive_flow_matches_oracle():
  """The Sol ephemeris' own integration (Ephemeris::Prolong's instance:
  34 bodies, 11 oblate, 5 with tesseral fields, Quinlan-Tremaine order 12 at
  10 min), three perturbed copies: generic kernels, surface frames, every
  geopotential interaction between massive bodies."""
  system = SolarSystem.from_fixture('sol_jd2451545')
  gpu = Ephemeris(system)
  method = _capi.QUINLAN_TREMAINE_1990_ORDER_12
  oracle = OracleEphemeris(system, method=method, step=600.0)
  order, _ = gpu.internal_order()
  state0 = ensemble_state(system, order, 3, 1e-9)
  t_final = 0.0 + 20 * 600.0
  ensemble = NBodyEnsemble(gpu, method, 600.0, state0, 0.0)
  steps = ensemble.solve(t_final)
  state, time = ensemble.state()
  expected_state, expected_time = state0.copy(), np.zeros(2)
  expected_steps = oracle.nbody_ensemble_flow(method, 600.0, expected_state,
                                              expected_time, t_final)
  assert steps == expected_steps == 20
  assert_bitwise_equal(state, expected_state, 'Sol state')
  # The unperturbed member reproduces the oracle's Prolong state.
  oracle.prolong(t_final)


def point_mass_system(n):
  bodies = [dict(name='b%05d' % i, mu=0.0, rotating=False, oblate=False,
                 q=[0.0, 0.0, 0.0], v=[0.0, 0.0, 0.0]) for i in range(n)]
  return SolarSystem(bodies, 0.0)


def random_cluster(n, seed=42):
  """BASELINE.json config 5 shape: positions uniform in a ball of 1 au,
  mu ~ logU(1e10, 1e14), virialised random velocities."""
  rng = np.random.Generator(np.random.MT19937(seed))
  au = 149597870700.0
  direction = rng.normal(size=(3, n))
  direction /= np.linalg.norm(direction, axis=0)
  q = au * rng.uniform(0, 1, n)**(1 / 3) * direction
  mu = np.exp(rng.uniform(np.log(1e10), np.log(1e14), n))
  sigma = np.sqrt(mu.sum() / au / 3)
  v = sigma * rng.normal(size=(3, n))
  return mu, q, v


@pytest.mark.parametrize('method', [_capi.QUINLAN_TREMAINE_1990_ORDER_12,
                                    _capi.BLANES_MOAN_2002_SRKN_14A])
def test_all_pairs_matches_oracle_within_tolerance(method):
  """The tiled all-pairs kernel sums each target's interactions over sources
  in index order, the reference scatters third-law updates pair by pair, so
  the comparison is to a tolerance: accelerations relative 1e-13, positions
  after the flow relative 1e-12 (BASELINE.json's tolerance)."""
  n = 700
  mu, q, v = random_cluster(n)
  system = point_mass_system(n)
  for b, m in zip(system.bodies, mu):
    b['mu'] = float(m)
  oracle = OracleEphemeris(system, method=method, step=3600.0,
                           initial_positions=q.T.copy(),
                           initial_velocities=v.T.copy())
  gpu = NBodyAllPairs(method, 3600.0, mu, q, v, 0.0)
  a = gpu.accelerations()
  expected_a = oracle.acceleration_between_all_massive_bodies(
      0.0, q.T.copy()).T
  scale = np.linalg.norm(expected_a, axis=0)
  assert np.max(np.linalg.norm(a - expected_a, axis=0) / scale) < 1e-13
  steps_wanted = 16 if method == _capi.QUINLAN_TREMAINE_1990_ORDER_12 else 3
  t_final = steps_wanted * 3600.0
  steps, forces = gpu.solve(t_final)
  state = np.zeros((12, n, 1))
  state[0:3, :, 0] = q
  state[6:9, :, 0] = v
  time = np.zeros(2)
  expected_steps = oracle.nbody_ensemble_flow(method, 3600.0, state, time,
                                              t_final)
  assert steps == expected_steps == steps_wanted
  if method == _capi.QUINLAN_TREMAINE_1990_ORDER_12:
    assert forces == 12 + 11 * 16 * 14 + (16 - 11)
  else:
    assert forces == 3 * 14
  q_gpu, v_gpu, time_gpu = gpu.state()
  assert_bitwise_equal(time_gpu, time)
  au = 149597870700.0
  assert np.max(np.abs(q_gpu - state[0:3, :, 0])) / au < 1e-12
  speed = np.linalg.norm(state[6:9, :, 0], axis=0).max()
  assert np.max(np.abs(v_gpu - state[6:9, :, 0])) / speed < 1e-10


def test_all_pairs_sharded_blocks_agree_with_the_whole():
  """Two target blocks of the same system produce the two halves of the
  accelerations bit for bit (the per-target summation order does not depend
  on the partition)."""
  n = 1024
  mu, q, v = random_cluster(n, seed=3)
  whole = NBodyAllPairs(_capi.BLANES_MOAN_2002_SRKN_14A, 60.0, mu, q, v, 0.0)
  a = whole.accelerations()
  for begin, end in ((0, 512), (512, 1024), (100, 333)):
    part = NBodyAllPairs(_capi.BLANES_MOAN_2002_SRKN_14A, 60.0, mu, q, v, 0.0,
                         i_begin=begin, i_end=end)
    assert_bitwise_equal(part.accelerations(), a[:, begin:end])
