"""Synthetic workloads of BASELINE.json's configs, seeded like
the reference's tests (std::mt19937_64(42) there; numpy's MT19937(42) here);
Ephemeris::Prolong on the device for the Sol system; and the loader of the
oracle-made Sol ephemeris segments that the parity tests upload
(tests/golden/, test input only).
"""
import os

import numpy as np

from .solar_system import GOLDEN, SolarSystem


def load_sol_ephemeris_segments():
  """tests/golden/sol_ephemeris_2d.npz (tools/make_ephemeris_fixture.py)."""
  data = np.load(os.path.join(GOLDEN, 'sol_ephemeris_2d.npz'))
  n = len([k for k in data.files if k.startswith('t_max_')])
  return [dict(t_min=float(data['t_min_%d' % i]), t_max=data['t_max_%d' % i],
               origin=data['origin_%d' % i], degree=data['degree_%d' % i],
               coefficients=data['coefficients_%d' % i]) for i in range(n)]


def upload_sol_ephemeris(ephemeris):
  """TEST INPUT: the segments the CPU oracle produced (the parity tests upload
  them so that a failure points at the flow, not at Prolong).  The product's
  own way is prolong_ephemeris below; the two are bit-equal
  (tests/test_gpu_continuous_trajectory.py)."""
  for body, segments in enumerate(load_sol_ephemeris_segments()):
    ephemeris.append_segments(body, segments)


def prolong_ephemeris(ephemeris, t_final, step=600.0, fitting_tolerance=1e-3,
                      method=None, t_initial=0.0):
  """Ephemeris::Prolong(t_final) from the initial state of the ephemeris'
  system, everything on the device (physics/ephemeris_body.hpp:312-344,
  1071-1116): the massive bodies are integrated with the ephemeris' own
  fixed-step integrator (QuinlanTremaine1990Order12 at 10 min for the Sol
  system, astronomy/..., ksp_plugin/integrators.cpp), every state goes to a
  device ContinuousTrajectory per body (Newhall fit every 8 points), and the
  resulting polynomials become the segments of `ephemeris`.  Returns the
  number of steps."""
  from . import _capi
  from .nbody import ContinuousTrajectories, NBodyEnsemble
  system = ephemeris.system
  if method is None:
    method = _capi.QUINLAN_TREMAINE_1990_ORDER_12
  order, _ = ephemeris.internal_order()
  n = len(order)
  state = np.zeros((12, n, 1))
  state[0:3, :, 0] = system.initial_positions()[order].T
  state[6:9, :, 0] = system.initial_velocities()[order].T
  segments = int((t_final - t_initial) / (8 * step)) + 3
  trajectories = ContinuousTrajectories(n, step, fitting_tolerance, segments,
                                        device=ephemeris.device)
  ensemble = NBodyEnsemble(ephemeris, method, step, state, t_initial)
  steps = 0
  t = t_final
  while trajectories.t_max() < t_final:  # A whole segment beyond t_final.
    steps += ensemble.solve_to_trajectories(t, trajectories)
    t += step
  for internal, caller in enumerate(order):
    ephemeris.append_segments(int(caller), trajectories.segments(internal))
  ensemble.close()
  trajectories.close()
  return steps


def rotation(inclination, node, argument):
  ci, si = np.cos(inclination), np.sin(inclination)
  co, so = np.cos(node), np.sin(node)
  cu, su = np.cos(argument), np.sin(argument)
  # Unit vectors along the radius and along the in-plane normal to it.
  radial = np.array([co * cu - so * su * ci, so * cu + co * su * ci, su * si])
  along = np.array([-co * su - so * cu * ci, -so * su + co * cu * ci, cu * si])
  return radial, along


def leo_probes(n, earth_position, earth_velocity, earth_mu, seed=42):
  """Config 2: circular LEO, a ~ U(6.6e6, 8.0e6) m, i ~ U(0, pi),
  node, argument of latitude ~ U(0, 2 pi).  Returns the (12, n) ODE state
  (errors zero)."""
  rng = np.random.Generator(np.random.MT19937(seed))
  a = rng.uniform(6.6e6, 8.0e6, n)
  inclination = rng.uniform(0, np.pi, n)
  node = rng.uniform(0, 2 * np.pi, n)
  u = rng.uniform(0, 2 * np.pi, n)
  radial, along = rotation(inclination, node, u)
  speed = np.sqrt(earth_mu / a)
  state = np.zeros((12, n))
  state[0:3] = np.asarray(earth_position)[:, None] + a * radial
  state[6:9] = np.asarray(earth_velocity)[:, None] + speed * along
  return state


def eccentric_probes(n, earth_position, earth_velocity, earth_mu, seed=42):
  """Config 3: a ~ logU(6.6e6, 4e8) m, e ~ U(0, 0.7) (perigee kept above
  6.6e6 m), random orientation, random true anomaly."""
  rng = np.random.Generator(np.random.MT19937(seed))
  a = np.exp(rng.uniform(np.log(6.6e6), np.log(4e8), n))
  e = rng.uniform(0, 0.7, n)
  e = np.minimum(e, 1 - 6.6e6 / a)
  inclination = np.arccos(rng.uniform(-1, 1, n))
  node = rng.uniform(0, 2 * np.pi, n)
  argument = rng.uniform(0, 2 * np.pi, n)
  true_anomaly = rng.uniform(0, 2 * np.pi, n)
  p = a * (1 - e * e)
  r = p / (1 + e * np.cos(true_anomaly))
  radial, along = rotation(inclination, node, argument + true_anomaly)
  vr = np.sqrt(earth_mu / p) * e * np.sin(true_anomaly)
  vt = np.sqrt(earth_mu / p) * (1 + e * np.cos(true_anomaly))
  state = np.zeros((12, n))
  state[0:3] = np.asarray(earth_position)[:, None] + r * radial
  state[6:9] = (np.asarray(earth_velocity)[:, None] + vr * radial +
                vt * along)
  return state


def sol_system():
  return SolarSystem.from_fixture('sol_jd2451545')
