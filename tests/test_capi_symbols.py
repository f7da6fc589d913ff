"""CPU: the C-ABI library loads without a GPU and exports every symbol that
include/principia_b200.h declares; entry points fail loudly (no CPU fallback)
when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from principia_b200 import _capi
from principia_b200.solar_system import SolarSystem

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
  text = open(os.path.join(REPO, 'include', 'principia_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(pb_[a-z0-9_]+)\s*\(', text)) -
                {'pb_exchange_positions_t', 'pb_status_t'})


def test_library_exports_every_declared_symbol():
  lib = _capi.library()
  names = declared_symbols()
  assert len(names) >= 28
  for name in names:
    assert hasattr(lib, name), name
  assert sorted(n for n, _, _ in _capi.SYMBOLS) == names


def test_fails_loudly_without_a_device():
  import torch
  if torch.cuda.is_available():
    pytest.skip('a CUDA device is present')
  lib = _capi.library()
  system = SolarSystem.from_fixture('trappist_jd2457000')
  bodies, keep = system.c_bodies()
  handle = ctypes.c_void_p()
  status = lib.pb_ephemeris_create(len(system.bodies), bodies, 2.0**-24, 0,
                                   ctypes.byref(handle))
  assert status == _capi.PB_FAILED_PRECONDITION
  assert b'no CPU fallback' in lib.pb_last_error()
  assert not handle.value


def test_argument_validation_needs_no_device():
  """Contract violations are reported as PB_INVALID_ARGUMENT with a message
  (where the reference would CHECK), before any CUDA call."""
  lib = _capi.library()
  null = ctypes.c_void_p()
  d3 = (ctypes.c_double * 3)()
  handle = ctypes.c_void_p()
  assert lib.pb_continuous_trajectories_create(
      0, 0, 600.0, 1e-3, 4, ctypes.byref(handle)) == _capi.PB_INVALID_ARGUMENT
  assert lib.pb_continuous_trajectories_create(
      0, 8, -1.0, 1e-3, 4, ctypes.byref(handle)) == _capi.PB_INVALID_ARGUMENT
  assert lib.pb_continuous_trajectories_create(
      0, 8, 600.0, 1e-3, 0, ctypes.byref(handle)) == _capi.PB_INVALID_ARGUMENT
  assert not handle.value
  assert lib.pb_continuous_trajectories_append(
      null, 0.0, d3, d3) == _capi.PB_INVALID_ARGUMENT
  assert lib.pb_continuous_trajectories_status(null) == _capi.PB_INVALID_ARGUMENT
  assert lib.pb_compute_gravitational_jerk_on_massless_bodies(
      null, 0.0, 1, d3, d3, d3) == _capi.PB_INVALID_ARGUMENT
  assert lib.pb_compute_gravitational_jerk_and_jacobian_on_massive_bodies(
      null, 0.0, None, None, d3, None) == _capi.PB_INVALID_ARGUMENT
  assert lib.pb_ephemeris_evaluate_all_velocities(
      null, 0.0, d3) == _capi.PB_INVALID_ARGUMENT
  assert lib.pb_downsampler_evaluate(null, 0.0, d3, d3,
                                     None) == _capi.PB_INVALID_ARGUMENT
  assert lib.pb_nbody_ensemble_solve_to_trajectories(
      null, 0.0, null, None) == _capi.PB_INVALID_ARGUMENT
  assert len(lib.pb_last_error()) > 0
