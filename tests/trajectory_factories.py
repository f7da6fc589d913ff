"""testing_utilities/discrete_trajectory_factories_body.hpp restated: the
timelines the reference's downsampling tests append."""
import numpy as np

import oracle_lib


def circular_timeline(omega, r, dt, t1, t2):
  """NewCircularTrajectoryTimeline (:90-111): t0 = 0, correctly-rounded sin and
  cos, `for (t = t1; t < t2; t += dt)`."""
  orc = oracle_lib.library()
  v = omega * r
  times, q, vel = [], [], []
  t = t1
  while t < t2:
    s = orc.orc_cr_sin(omega * (t - 0.0))
    c = orc.orc_cr_cos(omega * (t - 0.0))
    times.append(t)
    q.append([r * c, r * s, 0.0])
    vel.append([-v * s, v * c, 0.0])
    t += dt
  return np.array(times), np.array(q), np.array(vel)


def linear_timeline(velocity, dt, t1, t2):
  """NewLinearTrajectoryTimeline (:35-68): position = origin + v * (t - t1)."""
  velocity = np.asarray(velocity, dtype=np.float64)
  times, q, vel = [], [], []
  t = t1
  while t < t2:
    times.append(t)
    q.append(0.0 + velocity * (t - t1))
    vel.append(velocity.copy())
    t += dt
  return np.array(times), np.array(q), np.array(vel)
