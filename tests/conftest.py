import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: test needs a CUDA device (B200)')


@pytest.fixture(scope='session')
def sol_system():
  from principia_b200.solar_system import SolarSystem
  return SolarSystem.from_fixture('sol_jd2451545')


@pytest.fixture(scope='session')
def sol_oracle(sol_system):
  """The CPU oracle's Sol ephemeris, prolonged over the two days the golden
  segment fixture covers."""
  from oracle_lib import OracleEphemeris
  oracle = OracleEphemeris(sol_system)
  assert oracle.prolong(2 * 86400.0) == 0
  return oracle


@pytest.fixture(scope='session')
def sol_gpu(sol_system):
  """The product's Sol ephemeris on cuda:0 with the golden segments."""
  from principia_b200.ephemeris import Ephemeris
  from principia_b200 import workloads
  ephemeris = Ephemeris(sol_system)
  workloads.upload_sol_ephemeris(ephemeris)
  return ephemeris


def bits(a):
  return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def assert_bitwise_equal(actual, expected, what=''):
  actual = np.asarray(actual, dtype=np.float64)
  expected = np.asarray(expected, dtype=np.float64)
  assert actual.shape == expected.shape, what
  same = (bits(actual) == bits(expected)) | (np.isnan(actual) &
                                             np.isnan(expected))
  if not same.all():
    bad = np.argwhere(~same)
    first = tuple(bad[0])
    raise AssertionError(
        '%s: %d of %d values differ; first at %s: %r vs %r' %
        (what, len(bad), actual.size, first, actual[first], expected[first]))
