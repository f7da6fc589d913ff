import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from principia_b200 import _capi, workloads
from principia_b200.ephemeris import Ephemeris
system = workloads.sol_system()
gpu = Ephemeris(system)
workloads.upload_sol_ephemeris(gpu)
earth = system.bodies[system.index('Earth')]
state = workloads.leo_probes(37, earth['q'], earth['v'], earth['mu'], 5)
for t_mid in (45.0, 85.0, 200.0):
  inst = gpu.new_instance(_capi.QUINLAN_1999_ORDER_8A, 10.0, state, np.zeros(2))
  r = inst.flow(t_mid, sample_every=1, max_samples=500)
  print('t_mid', t_mid, 'steps', r['n_steps'], 'status', r['status'], 'nan', np.isnan(inst.state()[0]).any())
  r = inst.flow(2000.0, sample_every=1, max_samples=500)
  print('  then', r['n_steps'], r['status'], np.isnan(inst.state()[0]).any(), gpu.last_flow_path())
