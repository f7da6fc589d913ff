#!/usr/bin/env python3
"""The reference's MercuryPerihelionTest.Year1960 (astronomy/
mercury_perihelion_test.cpp:49-60,160-205) on the product: the 1950 Sol system
integrated over ten years on the GPU (QuinlanTremaine1990Order12 at 10 min, the
single-system massive-body flow), every state fitted on the device
(pb_nbody_ensemble_solve_to_trajectories, fitting tolerance 5 mm), Mercury and
the Sun evaluated at 1960-01-01 from the device polynomials; compared bit for
bit with the CPU oracle and with the figures the reference pins.

  python tools/mercury_on_gpu.py
"""
import json
import math
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

from principia_b200 import _capi  # noqa: E402
from principia_b200.ephemeris import Ephemeris  # noqa: E402
from principia_b200.nbody import ContinuousTrajectories, NBodyEnsemble  # noqa: E402
from principia_b200.solar_system import SolarSystem  # noqa: E402
from oracle_lib import OracleEphemeris  # noqa: E402


def main():
  system = SolarSystem.from_fixture('sol_jd2433282')
  t0 = system.epoch
  t_1960 = (2436934.5 - 2451545.0) * 86400.0
  step = 600.0
  gpu = Ephemeris(system)
  order, _ = gpu.internal_order()
  n = len(order)
  state = np.zeros((12, n, 1))
  state[0:3, :, 0] = system.initial_positions()[order].T
  state[6:9, :, 0] = system.initial_velocities()[order].T
  steps = int(math.ceil((t_1960 - t0) / step)) + 16
  trajectories = ContinuousTrajectories(n, step, 5e-3, steps // 8 + 2)
  ensemble = NBodyEnsemble(gpu, _capi.QUINLAN_TREMAINE_1990_ORDER_12, step,
                           state, t0)
  start = time.perf_counter()
  taken = 0
  t_final = t_1960
  while trajectories.t_max() < t_1960:  # Ephemeris::Prolong.
    taken += ensemble.solve_to_trajectories(t_final, trajectories)
    t_final += step
  gpu_seconds = time.perf_counter() - start
  q, v = trajectories.evaluate(t_1960)

  oracle = OracleEphemeris(system, fitting_tolerance=5e-3)
  start = time.perf_counter()
  assert oracle.prolong(t_1960) == 0
  cpu_seconds = time.perf_counter() - start
  expected_q = oracle.evaluate_all_positions(t_1960)[order].T
  expected_v = oracle.evaluate_all_velocities(t_1960)[order].T
  bitwise = (np.array_equal(q.view(np.int64), expected_q.view(np.int64)) and
             np.array_equal(v.view(np.int64), expected_v.view(np.int64)))

  internal = {caller: i for i, caller in enumerate(order)}
  sun, mercury = internal[system.index('Sun')], internal[system.index('Mercury')]
  r = q[:, mercury] - q[:, sun]
  w = v[:, mercury] - v[:, sun]
  mu = (system.bodies[system.index('Sun')]['mu'] +
        system.bodies[system.index('Mercury')]['mu'])
  h = np.cross(r, w)
  h_hat = h / np.linalg.norm(h)
  e_vector = np.cross(w, h) / mu - r / np.linalg.norm(r)
  e = np.linalg.norm(e_vector)
  node = np.cross([0.0, 0.0, 1.0], h)
  omega = math.atan2(np.dot(np.cross(node, e_vector), h_hat),
                     np.dot(node, e_vector))
  nu = math.atan2(np.dot(np.cross(e_vector, r), h_hat), np.dot(e_vector, r))
  big_e = 2 * math.atan(math.sqrt((1 - e) / (1 + e)) * math.tan(nu / 2))
  mean_anomaly = (big_e - e * math.sin(big_e)) % (2 * math.pi)
  degree = math.pi / 180
  arc_second = degree / 3600
  print(json.dumps({
      'workload': 'MercuryPerihelionTest.Year1960 on the GPU: 34 bodies, '
                  'QuinlanTremaine1990Order12 at 10 min, 1950 -> 1960, states '
                  'fitted on the device',
      'steps': taken, 'segments': trajectories.segment_count(),
      'gpu_seconds': gpu_seconds, 'gpu_steps_per_second': taken / gpu_seconds,
      'oracle_one_core_seconds': cpu_seconds,
      'degrees_of_freedom_at_1960_bitwise_equal_to_oracle': bool(bitwise),
      'argument_of_periapsis_deficit_arcsec':
          (6.748880915164143e+01 * degree - omega) / arc_second,
      'reference_pins_arcsec': '4.206(1)',
      'mean_anomaly_excess_arcsec':
          (mean_anomaly - 1.437427180144607e+02 * degree) / arc_second,
      'reference_pins_mean_anomaly_arcsec': '17.03(1)'}))
  assert bitwise


if __name__ == '__main__':
  main()
