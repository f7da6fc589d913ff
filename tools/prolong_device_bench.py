import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from principia_b200 import _capi
from principia_b200.ephemeris import Ephemeris
from principia_b200.nbody import NBodyEnsemble, ContinuousTrajectories
from principia_b200.solar_system import SolarSystem
system = SolarSystem.from_fixture('sol_jd2451545')
gpu = Ephemeris(system)
order, _ = gpu.internal_order()
n = len(order)
state = np.zeros((12, n, 1))
state[0:3, :, 0] = system.initial_positions()[order].T
state[6:9, :, 0] = system.initial_velocities()[order].T
h = 600.0
for sink in (False, True):
  ens = NBodyEnsemble(gpu, _capi.QUINLAN_TREMAINE_1990_ORDER_12, h, state, 0.0)
  traj = ContinuousTrajectories(n, h, 1e-3, 400)
  if sink:
    ens.solve_to_trajectories(2 * 86400.0, traj)
  else:
    ens.solve(2 * 86400.0)
  t0 = time.perf_counter()
  steps = ens.solve_to_trajectories(22 * 86400.0, traj) if sink else ens.solve(22 * 86400.0)
  dt = time.perf_counter() - t0
  print('sink', sink, 'steps', steps, 'seconds', dt, 'steps/s', steps / dt)
