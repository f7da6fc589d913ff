"""principia_b200: B200-native batched gravitational-flow engine behind the
surface of mockingbirdnest/Principia's physics::Ephemeris.

The compute path is the CUDA library `libprincipia_b200.so` (sources under
`principia_b200/csrc`, C ABI in `include/principia_b200.h`); this package is the
thin Python mirror used by the tests and by bench.py.  There is no CPU
fallback: constructing an `Ephemeris` without the built library, or without a
CUDA device, raises.
"""
from . import _capi
from .solar_system import SolarSystem

__all__ = ['SolarSystem', '_capi']
