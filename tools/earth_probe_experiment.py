#!/usr/bin/env python3
"""The reference's EarthProbe test (physics/ephemeris_test.cpp:374-497) run
through the CPU oracle: the Earth alone in uniform motion, a probe 10^9 m away
whose constant intrinsic acceleration cancels the Earth's pull, adaptive flow
over one period with tolerances of 1 nm and 2.6 fm/s.  The reference pins the
size of the trajectory (358 ... 421 points depending on the platform) and the
distance of the Earth's polynomial positions from period * v_earth (241 ... 635
ULPs).  NOT REPRODUCED by the oracle (4 points; <= 20 ULPs): see DESIGN.md §2.
"""
import math, sys
import numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from principia_b200 import _capi
from principia_b200.solar_system import SolarSystem
from oracle_lib import OracleEphemeris
system = SolarSystem.from_fixture('sol_jd2451545')
mu_e = system.bodies[system.index('Earth')]['mu']
mu_m = system.bodies[system.index('Moon')]['mu']
G = 6.67430e-11
m_e, m_m = mu_e / G, mu_m / G
a = math.sqrt(0.0*0.0 + (0.0-4e8)*(0.0-4e8) + 0.0*0.0)
period = 2 * math.pi * math.sqrt((a * a * a) / (mu_e + mu_m))
wy = (0.0 - 0.0) * m_e
wy += (4e8 - 0.0) * m_m
com_y = 0.0 + wy / (m_e + m_m)
norm1 = math.sqrt(0.0 + (0.0 - com_y) * (0.0 - com_y) + 0.0)
v1x = -2 * math.pi * norm1 / period
print('period', period, 'v1x', v1x)
earth_only = system.copy()
for name in list(earth_only.names()):
  if name != 'Earth':
    earth_only.remove_massive_body(name)
earth_only.remove_oblateness('Earth')
distance = 1e9
for method, label in ((_capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, '/0'), (_capi.QUINLAN_1999_ORDER_8A, '/1')):
  oracle = OracleEphemeris(earth_only, method=method, step=period / 100, fitting_tolerance=5e-3,
                           initial_positions=np.zeros((1, 3)), initial_velocities=np.array([[v1x, 0.0, 0.0]]),
                           initial_time=0.0)
  assert oracle.prolong(period) == 0
  state = np.zeros((12, 1)); state[1, 0] = distance; state[6, 0] = v1x
  time = np.zeros((2, 1))
  burns = np.zeros((8, 1)); burns[1] = 1.0; burns[3] = mu_e / (distance * distance); burns[4] = 1.0; burns[6] = -1e30; burns[7] = 1e30
  # -1e30: (t - t0) * 0 = 0 exactly.
  result = oracle.flow_with_adaptive_step(period, state, time, 1e-9, 2.6e-15, max_steps=1000000, burns=burns)
  print(label, 'points', int(result['accepted'][0]) + 1, 'status', result['status'][0], 'x', state[0,0], 'period*v', 1.00*period*state[6,0], 'y', state[1,0])

def ulps(a, b):
  ia = np.float64(a).view(np.int64); ib = np.float64(b).view(np.int64)
  return abs(int(ia) - int(ib))
for method, label in ((_capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, '/0'), (_capi.QUINLAN_1999_ORDER_8A, '/1')):
  oracle = OracleEphemeris(earth_only, method=method, step=period / 100, fitting_tolerance=5e-3,
                           initial_positions=np.zeros((1, 3)), initial_velocities=np.array([[v1x, 0.0, 0.0]]),
                           initial_time=0.0)
  assert oracle.prolong(period) == 0
  seg = oracle.segments(0)
  print(label, 'segments', len(seg['t_max']), 'degrees', sorted(set(seg['degree'])), 't_max', oracle.t_max() / period)
  for i in (25, 50, 75, 100):
    x = oracle.evaluate_all_positions(0.0 + i * period / 100)[0][0]
    print('  i', i, 'x', x, 'ulps vs i/100*period*v1', ulps(x, i / 100 * period * v1x), 'y', oracle.evaluate_all_positions(i * period / 100)[0][1])

print('--- v_earth from the polynomial derivative at t0 + period')
for method, label in ((_capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL, '/0'), (_capi.QUINLAN_1999_ORDER_8A, '/1')):
  oracle = OracleEphemeris(earth_only, method=method, step=period / 100, fitting_tolerance=5e-3,
                           initial_positions=np.zeros((1, 3)), initial_velocities=np.array([[v1x, 0.0, 0.0]]),
                           initial_time=0.0)
  assert oracle.prolong(period) == 0
  seg = oracle.segments(0)
  t = 0.0 + period
  k = int(np.searchsorted(seg['t_max'], t))  # first with t <= t_max
  c = seg['coefficients'][k][:, 0]; d = int(seg['degree'][k]); x = t - seg['origin'][k]
  # Horner-ish derivative (degree 3): c1 + 2 c2 x + 3 c3 x^2 with fma like Estrin EvaluateDerivative.
  v_earth = math.fma(math.fma(3 * c[3], x, 2 * c[2]), x, c[1]) if hasattr(math, 'fma') else (3*c[3]*x + 2*c[2])*x + c[1]
  print(label, 'degree', d, 'c', list(c[:4]), 'x', x, 'v_earth', repr(v_earth), 'v1', repr(v1x), 'rel', (v_earth - v1x) / v1x)
  for i, want in ((25, '539-541'), (50, '241-242'), (75, '403-405'), (100, '633-635')):
    xe = oracle.evaluate_all_positions(0.0 + i * period / 100)[0][0]
    print('  i', i, 'ulps vs i/100*period*v_earth', ulps(xe, (i / 100) * period * v_earth), 'reference', want)
