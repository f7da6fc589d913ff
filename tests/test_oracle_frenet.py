"""CPU: the oracle's generalized adaptive flow with Frenet-frame burns
(Manœuvre::FrenetIntrinsicAcceleration in a body-centred non-rotating frame,
oracle/ephemeris.hpp FrenetDirection).  No known answer of the reference pins
the frame algebra (its tests use mock frames), so this checks the physics the
Frenet trihedron implies: a tangential burn changes the orbital energy by
about the work of the thrust along the velocity, a binormal burn tilts the
orbital plane without changing the energy, an inward normal burn does neither
to first order."""
import numpy as np

from principia_b200 import _capi, workloads


def test_frenet_burn_physics(sol_system, sol_oracle):
  earth_index = sol_system.index('Earth')
  earth = sol_system.bodies[earth_index]
  state = workloads.leo_probes(4, earth['q'], earth['v'], earth['mu'], 3)
  state[:] = state[:, :1]  # The same orbit four times.
  burns = np.zeros((8, 4))
  burns[0:3, 0] = [1.0, 0.0, 0.0]   # Tangent.
  burns[0:3, 1] = [0.0, 1.0, 0.0]   # Normal.
  burns[0:3, 2] = [0.0, 0.0, 1.0]   # Binormal.
  burns[0:3, 3] = [1.0, 0.0, 0.0]   # (Switched off below.)
  burns[3] = 1e4                    # 1 m/s^2 on 10 t.
  burns[4] = 1e4
  burns[5] = 0.0
  burns[6] = 50.0
  burns[7] = 150.0                  # 100 s: 100 m/s.
  burns[3, 3] = 0.0
  centres = np.full(4, earth_index, dtype=np.int32)
  times = np.zeros((2, 4))
  initial = state.copy()
  result = sol_oracle.flow_with_adaptive_step(
      400.0, state, times, 1e-3, 1e-3, burns=burns,
      method=_capi.FINE_1987_RKNG_34, frenet_centres=centres)
  assert np.all(result['status'] == 0)
  q_earth = sol_oracle.evaluate_all_positions(400.0)[earth_index]
  v_earth = sol_oracle.evaluate_all_velocities(400.0)[earth_index]
  q = state[0:3] - q_earth[:, None]
  v = state[6:9] - v_earth[:, None]
  energy = 0.5 * (v * v).sum(axis=0) - earth['mu'] / np.linalg.norm(q, axis=0)
  momentum = np.cross(q.T, v.T)
  speed = np.linalg.norm(v[:, 3])
  # Tangential: the energy grows by about v dv (here 7.5 km/s x 100 m/s).
  assert abs((energy[0] - energy[3]) / (speed * 100.0) - 1.0) < 0.02
  # Normal and binormal: the thrust is orthogonal to the velocity.
  assert abs(energy[1] - energy[3]) < 0.02 * speed * 100.0
  assert abs(energy[2] - energy[3]) < 0.02 * speed * 100.0
  # Binormal: the orbital plane tilts by about dv / v.
  def angle(a, b):
    return np.arccos(np.dot(a, b) / np.linalg.norm(a) / np.linalg.norm(b))
  tilt = angle(momentum[2], momentum[3])
  assert abs(tilt / (100.0 / speed) - 1.0) < 0.05
  assert angle(momentum[0], momentum[3]) < 5e-3 * tilt
  assert angle(momentum[1], momentum[3]) < 5e-3 * tilt
