"""The set-up of the reference's EphemerisTest.EarthProbe
(physics/ephemeris_test.cpp:134-174,374-497): the Earth alone (non-oblate) in
uniform motion at the velocity it has in the Earth-Moon barycentric frame, and a
probe 10^9 m away with the same velocity whose constant intrinsic acceleration
cancels the Earth's pull.  "The solution is a line, so the rounding errors
dominate": every number the reference pins here is made of rounding errors."""
import math

import numpy as np

from principia_b200 import _capi

# solar_system_.epoch() of the fixture, JD 2433282.5 (:122-126), in seconds
# from J2000.  At this distance from J2000 an ulp of time is 2.4e-7 s, 2.9 um of
# Earth motion: this is what the reference's ULP distances are made of.
T0 = (2433282.5 - 2451545.0) * 86400.0
GRAVITATIONAL_CONSTANT = 6.67430e-11  # quantities/constants.hpp:30-31.
DISTANCE = 1e9

# AnyOf(...) of the reference for the size of the probe's trajectory, :465-472.
REFERENCE_TRAJECTORY_SIZES = (358, 366, 373, 379, 396, 406, 420, 421)
# AlmostEquals(i / 100 * period * v_earth, min, max) for the Earth, :437-448.
REFERENCE_EARTH_ULPS = {25: (539, 541), 50: (241, 242), 75: (403, 405),
                        100: (633, 635)}
# AlmostEquals(period * v_probe, 219, 266) for the probe, :473-474.
REFERENCE_PROBE_ULPS = (219, 266)

INTEGRATORS = (_capi.MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL,
               _capi.QUINLAN_1999_ORDER_8A)


def ulp_distance(a, b):
  ia = int(np.float64(a).view(np.int64))
  ib = int(np.float64(b).view(np.int64))
  return abs(ia - ib)


# EphemerisTest.Moon, :330-372: AlmostEquals(i / 100 * period * v, min, max).
REFERENCE_MOON_ULPS = {25: (343, 345), 50: (151, 153), 75: (512, 516),
                       100: (401, 403)}
# EphemerisTest.EarthMoon, :262-318: bounds in metres on the coordinate that
# vanishes on a circular orbit, relative to the centre of mass.
REFERENCE_EARTH_MOON_BOUNDS = {
    'Earth': {25: ('y', 6e-4), 50: ('x', 7e-3), 75: ('y', 2e-2),
              100: ('x', 3e-2)},
    'Moon': {25: ('y', 5e-2), 50: ('x', 6e-1), 75: ('y', 2.0),
             100: ('x', 2.0)}}


def set_up_earth_moon(sol_system):
  """SetUpEarthMoonSystem, :134-174: returns (system of the non-oblate Earth
  and Moon, initial positions, initial velocities, centre of mass (y), period).
  """
  mu_earth = sol_system.bodies[sol_system.index('Earth')]['mu']
  mu_moon = sol_system.bodies[sol_system.index('Moon')]['mu']
  m_earth = mu_earth / GRAVITATIONAL_CONSTANT
  m_moon = mu_moon / GRAVITATIONAL_CONSTANT
  semi_major_axis = math.sqrt(0.0 * 0.0 + (0.0 - 4e8) * (0.0 - 4e8) +
                              0.0 * 0.0)
  period = 2 * math.pi * math.sqrt(
      (semi_major_axis * semi_major_axis * semi_major_axis) /
      (mu_earth + mu_moon))
  weighted_sum_y = (0.0 - 0.0) * m_earth
  weighted_sum_y += (4e8 - 0.0) * m_moon
  centre_of_mass_y = 0.0 + weighted_sum_y / (m_earth + m_moon)
  norm1 = math.sqrt(0.0 + (0.0 - centre_of_mass_y) *
                    (0.0 - centre_of_mass_y) + 0.0)
  norm2 = math.sqrt(0.0 + (4e8 - centre_of_mass_y) *
                    (4e8 - centre_of_mass_y) + 0.0)
  v1 = -2 * math.pi * norm1 / period
  v2 = 2 * math.pi * norm2 / period
  system = sol_system.copy()
  for name in list(system.names()):
    if name not in ('Earth', 'Moon'):
      system.remove_massive_body(name)
  system.remove_oblateness('Earth')
  system.remove_oblateness('Moon')
  positions = np.zeros((2, 3))
  velocities = np.zeros((2, 3))
  positions[system.index('Moon'), 1] = 4e8
  velocities[system.index('Earth'), 0] = v1
  velocities[system.index('Moon'), 0] = v2
  return system, positions, velocities, centre_of_mass_y, period


def set_up(sol_system):
  """Returns (earth_only_system, period, earth_velocity_x)."""
  mu_earth = sol_system.bodies[sol_system.index('Earth')]['mu']
  mu_moon = sol_system.bodies[sol_system.index('Moon')]['mu']
  # MassiveBody::mass(), physics/massive_body_body.hpp:35.
  m_earth = mu_earth / GRAVITATIONAL_CONSTANT
  m_moon = mu_moon / GRAVITATIONAL_CONSTANT
  # SetUpEarthMoonSystem, :152-168.
  semi_major_axis = math.sqrt(0.0 * 0.0 + (0.0 - 4e8) * (0.0 - 4e8) +
                              0.0 * 0.0)
  period = 2 * math.pi * math.sqrt(
      (semi_major_axis * semi_major_axis * semi_major_axis) /
      (mu_earth + mu_moon))
  # Barycentre({q1, q2}, {m1, m2}), geometry/barycentre_calculator_body.hpp.
  weighted_sum_y = (0.0 - 0.0) * m_earth
  weighted_sum_y += (4e8 - 0.0) * m_moon
  centre_of_mass_y = 0.0 + weighted_sum_y / (m_earth + m_moon)
  norm = math.sqrt(0.0 + (0.0 - centre_of_mass_y) * (0.0 - centre_of_mass_y) +
                   0.0)
  v1 = -2 * math.pi * norm / period
  earth_only = sol_system.copy()
  for name in list(earth_only.names()):
    if name != 'Earth':
      earth_only.remove_massive_body(name)
  earth_only.remove_oblateness('Earth')
  return earth_only, period, v1


def probe(sol_system, v1):
  """The initial ODE::State (12, 1) and time (2, 1) of the probe, and the
  tabulated form of its intrinsic acceleration: the constant vector
  (0, mu / distance^2, 0), i.e. direction (0, 1, 0) * thrust / mass with
  mass = 1 and no mass flow ((t - t_initial) * 0 = 0 exactly)."""
  mu_earth = sol_system.bodies[sol_system.index('Earth')]['mu']
  state = np.zeros((12, 1))
  state[1, 0] = DISTANCE
  state[6, 0] = v1
  time = np.zeros((2, 1))
  time[0, 0] = T0
  burns = np.zeros((8, 1))
  burns[1] = 1.0
  burns[3] = mu_earth / (DISTANCE * DISTANCE)
  burns[4] = 1.0
  burns[6] = -1e30
  burns[7] = 1e30
  return state, time, burns
