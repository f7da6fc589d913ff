"""Host-side description of a system of massive bodies: the analogue of the
reference's physics::SolarSystem (physics/solar_system_body.hpp:100-168), fed
from the JSON files under principia_b200/data/ (tools/make_system_fixtures.py
documents how those are derived from the reference's configuration files).
"""
import copy
import ctypes
import json
import os

import numpy as np

from . import _capi

# The gravity models and initial states (product inputs).
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
# Test vectors (the oracle's ephemeris segments, reference known answers).
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      'tests', 'golden')


class SolarSystem:
  """Bodies in name order (std::map order of the reference) with their
  initial degrees of freedom at `epoch`."""

  def __init__(self, bodies, epoch):
    self.bodies = bodies
    self.epoch = epoch

  @classmethod
  def from_fixture(cls, name):
    with open(os.path.join(DATA, name + '.json')) as f:
      data = json.load(f)
    return cls(data['bodies'], data['epoch'])

  def names(self):
    return [b['name'] for b in self.bodies]

  def index(self, name):
    return self.names().index(name)

  def copy(self):
    return SolarSystem(copy.deepcopy(self.bodies), self.epoch)

  # physics/solar_system_body.hpp:470-500.
  def limit_oblateness_to_degree(self, name, max_degree):
    b = self.bodies[self.index(name)]
    if not b['oblate']:
      return
    if max_degree < 2:
      b['oblate'] = False
      return
    degree = min(b['degree'], max_degree)
    size = (degree + 1) * (degree + 2) // 2
    b['degree'] = degree
    b['cos'] = b['cos'][:size]
    b['sin'] = b['sin'][:size]
    b['zonal'] = all(
        b['cos'][n * (n + 1) // 2 + m] == 0 and b['sin'][n * (n + 1) // 2 + m] == 0
        for n in range(degree + 1) for m in range(1, n + 1))

  # physics/solar_system_body.hpp:503-525.
  def limit_oblateness_to_zonal(self, name):
    b = self.bodies[self.index(name)]
    if not b['oblate']:
      return
    for n in range(b['degree'] + 1):
      for m in range(1, n + 1):
        b['cos'][n * (n + 1) // 2 + m] = 0.0
        b['sin'][n * (n + 1) // 2 + m] = 0.0
    b['zonal'] = True

  def remove_massive_body(self, name):
    del self.bodies[self.index(name)]

  def remove_oblateness(self, name):
    self.bodies[self.index(name)]['oblate'] = False

  def initial_positions(self):
    return np.array([b['q'] for b in self.bodies], dtype=np.float64)

  def initial_velocities(self):
    return np.array([b['v'] for b in self.bodies], dtype=np.float64)

  def c_bodies(self):
    """Returns (ctypes array of pb_body_t, keep-alive list)."""
    array = (_capi.Body * len(self.bodies))()
    keep = []
    for i, b in enumerate(self.bodies):
      c = array[i]
      c.gravitational_parameter = b['mu']
      c.is_rotating = int(b.get('rotating', False))
      if c.is_rotating:
        c.min_radius = b['min_radius']
        c.mean_radius = b['mean_radius']
        c.reference_angle = b['reference_angle']
        c.reference_instant = b['reference_instant']
        c.angular_frequency = b['angular_frequency']
        c.right_ascension_of_pole = b['right_ascension']
        c.declination_of_pole = b['declination']
      c.is_oblate = int(b.get('oblate', False))
      if c.is_oblate:
        c.geopotential_degree = b['degree']
        c.is_zonal = int(b['zonal'])
        c.reference_radius = b['reference_radius']
        cos = np.array(b['cos'], dtype=np.float64)
        sin = np.array(b['sin'], dtype=np.float64)
        keep += [cos, sin]
        c.cos = cos.ctypes.data_as(_capi.c_double_p)
        c.sin = sin.ctypes.data_as(_capi.c_double_p)
    return array, keep
