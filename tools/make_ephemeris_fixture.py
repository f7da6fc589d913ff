#!/usr/bin/env python3
"""Generates tests/golden/sol_ephemeris_2d.npz: the ContinuousTrajectory
segments of the 34-body Sol ephemeris (sol_numerics_blueprint: Quinlan-Tremaine
order 12 at 10 min, fitting tolerance 1 mm, geopotential tolerance 2^-24) from
JD2451545.0 over two days, as produced by the CPU oracle's Prolong + Newhall
fit.  These segments are INPUT DATA of the massless hot path (what the
reference's ContinuousTrajectory::Append would hand over); the
parity tests and bench.py upload them with pb_ephemeris_append_segments.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

from principia_b200.solar_system import SolarSystem  # noqa: E402
from oracle_lib import OracleEphemeris  # noqa: E402

DAYS = 2


def main():
  system = SolarSystem.from_fixture('sol_jd2451545')
  oracle = OracleEphemeris(system)
  assert oracle.prolong(DAYS * 86400.0) == 0
  out = {}
  for i, name in enumerate(system.names()):
    seg = oracle.segments(i)
    out['t_min_%d' % i] = np.float64(seg['t_min'])
    out['t_max_%d' % i] = seg['t_max']
    out['origin_%d' % i] = seg['origin']
    out['degree_%d' % i] = seg['degree']
    out['coefficients_%d' % i] = seg['coefficients']
  path = os.path.join(REPO, 'tests', 'golden', 'sol_ephemeris_2d.npz')
  np.savez_compressed(path, **out)
  print(path, os.path.getsize(path), 'bytes; t_max', oracle.t_max())


if __name__ == '__main__':
  main()
