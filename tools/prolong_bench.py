#!/usr/bin/env python3
"""Times Ephemeris::Prolong of the Sol system (34 bodies, QuinlanTremaine1990
Order12 at 10 min, fitting tolerance 1 mm) through the C++ facade: the massive
bodies are integrated on the GPU, the Newhall fits run on the host.

  python tools/prolong_bench.py [days]
"""
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

from principia_b200 import _capi, workloads  # noqa: E402
import oracle_lib  # noqa: E402


def main():
  days = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
  directory = os.path.join(REPO, 'tests', 'cpp')
  subprocess.check_call(
      ['/usr/bin/g++', '-std=c++17', '-O2', '-ffp-contract=off', '-fPIC',
       '-shared', '-o', 'libfacadetest.so', 'facade_capi.cpp',
       '-L../../principia_b200', '-lprincipia_b200',
       '-Wl,-rpath,$ORIGIN/../../principia_b200'], cwd=directory)
  lib = ctypes.CDLL(os.path.join(directory, 'libfacadetest.so'))
  _d, _i32, _vp = ctypes.c_double, ctypes.c_int32, ctypes.c_void_p
  lib.ft_create.restype = _vp
  lib.ft_create.argtypes = [_i32, ctypes.POINTER(_capi.Body), _capi.c_double_p,
                            _capi.c_double_p, _d, _d, _d, _i32, _d]
  lib.ft_prolong.argtypes = [_vp, _d]
  lib.ft_t_max.restype = _d
  lib.ft_t_max.argtypes = [_vp]
  system = workloads.sol_system()
  bodies, keep = system.c_bodies()
  q0 = np.ascontiguousarray(system.initial_positions())
  v0 = np.ascontiguousarray(system.initial_velocities())
  handle = lib.ft_create(len(system.bodies), bodies, oracle_lib.dptr(q0),
                         oracle_lib.dptr(v0), 0.0, 1e-3, 2.0**-24,
                         _capi.QUINLAN_TREMAINE_1990_ORDER_12, 600.0)
  assert handle
  assert lib.ft_prolong(handle, 86400.0) == 0  # Startup and warm-up.
  t0 = time.perf_counter()
  assert lib.ft_prolong(handle, (1.0 + days) * 86400.0) == 0
  seconds = time.perf_counter() - t0
  steps = days * 144
  print(json.dumps({'workload': 'Ephemeris::Prolong, Sol 34 bodies, QT-12 at 10 min',
                    'days': days, 'seconds': seconds,
                    'steps_per_second': steps / seconds,
                    't_max': lib.ft_t_max(handle)}))


if __name__ == '__main__':
  main()
