"""Python mirror of the massive-body flows of the C ABI: ensembles of small
systems sharing one gravity model (the reference runs one Ephemeris per
ensemble member, astronomy/trappist_dynamics_test.cpp:275-297) and the
large-N all-pairs system sharded by target block."""
import ctypes

import numpy as np

from . import _capi
from .ephemeris import _check, _dptr

_vp = ctypes.c_void_p


class NBodyEnsemble:
  """pb_nbody_ensemble_t: n_systems copies of the ephemeris' bodies, advanced
  by a fixed-step integrator instance (Integrator::Instance::Solve)."""

  def __init__(self, ephemeris, method, step, state, initial_time):
    """state: (12, n_bodies, n_systems) float64, bodies in INTERNAL order."""
    self.lib = _capi.library()
    self.ephemeris = ephemeris
    state = np.ascontiguousarray(state, dtype=np.float64)
    assert state.shape[0] == 12 and state.shape[1] == ephemeris.n_bodies
    self.shape = state.shape
    handle = _vp()
    _check(self.lib.pb_nbody_ensemble_create(
        ephemeris.handle, method, step, state.shape[2], _dptr(state),
        initial_time, ctypes.byref(handle)))
    self.handle = handle

  def close(self):
    if getattr(self, 'handle', None):
      self.lib.pb_nbody_ensemble_destroy(self.handle)
      self.handle = None

  def __del__(self):
    self.close()

  def solve(self, t_final):
    n_steps = ctypes.c_int64()
    _check(self.lib.pb_nbody_ensemble_solve(self.handle, t_final,
                                            ctypes.byref(n_steps)))
    return n_steps.value

  def solve_to_trajectories(self, t_final, trajectories):
    """Solve(t_final) with every appended state going to `trajectories` (a
    ContinuousTrajectories of n_bodies * n_systems) on the device."""
    n_steps = ctypes.c_int64()
    _check(self.lib.pb_nbody_ensemble_solve_to_trajectories(
        self.handle, t_final, trajectories.handle, ctypes.byref(n_steps)))
    return n_steps.value

  def state(self):
    state = np.zeros(self.shape)
    time = np.zeros(2)
    _check(self.lib.pb_nbody_ensemble_get_state(self.handle, _dptr(state),
                                                _dptr(time)))
    return state, time


class ContinuousTrajectories:
  """pb_continuous_trajectories_t: n ContinuousTrajectory objects appended in
  lockstep, fitted (Newhall) and evaluated on the device."""

  def __init__(self, n, step, tolerance, max_segments, device=0):
    self.lib = _capi.library()
    self.n = n
    handle = _vp()
    _check(self.lib.pb_continuous_trajectories_create(
        device, n, step, tolerance, max_segments, ctypes.byref(handle)))
    self.handle = handle

  def close(self):
    if getattr(self, 'handle', None):
      self.lib.pb_continuous_trajectories_destroy(self.handle)
      self.handle = None

  def __del__(self):
    self.close()

  def append(self, time, q, v):
    """q, v: (3, n).  Raises StatusError(PB_INVALID_ARGUMENT) on the
    "apocalypse" of any trajectory."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    assert q.shape == v.shape == (3, self.n)
    _check(self.lib.pb_continuous_trajectories_append(self.handle, time,
                                                      _dptr(q), _dptr(v)))

  def append_device(self, time, q_pointer, v_pointer, stream=None):
    _check(self.lib.pb_continuous_trajectories_append_device(
        self.handle, time, q_pointer, v_pointer, stream))

  def status(self):
    return self.lib.pb_continuous_trajectories_status(self.handle)

  def t_min(self):
    return self.lib.pb_continuous_trajectories_t_min(self.handle)

  def t_max(self):
    return self.lib.pb_continuous_trajectories_t_max(self.handle)

  def segment_count(self):
    return self.lib.pb_continuous_trajectories_segment_count(self.handle)

  def segments(self, trajectory):
    count = self.segment_count()
    t_min = ctypes.c_double()
    t_max = np.zeros(count)
    origin = np.zeros(count)
    degree = np.zeros(count, dtype=np.int32)
    coefficients = np.zeros((count, 18, 3))
    _check(self.lib.pb_continuous_trajectories_get_segments(
        self.handle, trajectory, ctypes.byref(t_min), _dptr(t_max),
        _dptr(origin), degree.ctypes.data_as(_capi.c_int32_p),
        _dptr(coefficients)))
    return dict(t_min=t_min.value, t_max=t_max, origin=origin, degree=degree,
                coefficients=coefficients)

  def evaluate(self, t):
    """((3, n) positions, (3, n) velocities) at t."""
    positions = np.zeros((3, self.n))
    velocities = np.zeros((3, self.n))
    _check(self.lib.pb_continuous_trajectories_evaluate(
        self.handle, t, _dptr(positions), _dptr(velocities)))
    return positions, velocities

  def evaluate_device(self, t, positions_pointer, velocities_pointer,
                      stream=None):
    _check(self.lib.pb_continuous_trajectories_evaluate_device(
        self.handle, t, positions_pointer, velocities_pointer, stream))


class NBodyAllPairs:
  """pb_nbody_allpairs_t: n point masses, this rank owning targets
  [i_begin, i_end)."""

  def __init__(self, method, step, mu, q, v, initial_time, i_begin=0,
               i_end=None, device=0, exchange=None):
    """mu: (n,), q, v: (3, n).  `exchange(sources_device_pointer, n, i_begin,
    i_end, stream)` is called after every position update when given (the
    sources are [n][4] doubles x, y, z, mu); prefer set_nccl, which makes the
    library issue one ncclAllGather itself."""
    self.lib = _capi.library()
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    self.n = len(mu)
    self.i_begin = i_begin
    self.i_end = self.n if i_end is None else i_end
    if exchange is None:
      self._callback = ctypes.cast(None, _capi.EXCHANGE_POSITIONS)
    else:
      self._exchange_error = None

      def trampoline(user, sources, n, begin, end, stream):
        # An exception must not escape into ctypes (it would only be printed
        # and the force would run on stale positions): report it through the
        # status and re-raise it once the C call has returned.
        try:
          exchange(sources, n, begin, end, stream)
          return _capi.PB_OK
        except BaseException as error:  # noqa: B902
          self._exchange_error = error
          return _capi.PB_ABORTED
      self._callback = _capi.EXCHANGE_POSITIONS(trampoline)
    handle = _vp()
    _check(self.lib.pb_nbody_allpairs_create(
        device, method, step, self.n, self.i_begin, self.i_end, _dptr(mu),
        _dptr(q), _dptr(v), initial_time, self._callback, None,
        ctypes.byref(handle)))
    self.handle = handle

  def close(self):
    if getattr(self, 'handle', None):
      self.lib.pb_nbody_allpairs_destroy(self.handle)
      self.handle = None

  def __del__(self):
    self.close()

  def _checked(self, code):
    error = getattr(self, '_exchange_error', None)
    if error is not None:
      self._exchange_error = None
      raise error
    _check(code)

  def set_nccl(self, communicator):
    """`communicator`: a sharding.NcclCommunicator (or None to detach): the
    library then all-gathers the owned slices itself, one ncclAllGather per
    force evaluation on the force stream."""
    self._communicator = communicator  # Keeps it alive.
    _check(self.lib.pb_nbody_allpairs_set_nccl(
        self.handle, None if communicator is None else communicator.handle))

  def set_arithmetic(self, arithmetic):
    """_capi.PB_ARITHMETIC_EXACT (default) or PB_ARITHMETIC_CONTRACTED."""
    _check(self.lib.pb_nbody_allpairs_set_arithmetic(self.handle, arithmetic))

  def solve(self, t_final):
    n_steps, n_forces = ctypes.c_int64(), ctypes.c_int64()
    self._checked(self.lib.pb_nbody_allpairs_solve(self.handle, t_final,
                                                   ctypes.byref(n_steps),
                                                   ctypes.byref(n_forces)))
    return n_steps.value, n_forces.value

  def state(self):
    n_local = self.i_end - self.i_begin
    q, v, time = np.zeros((3, n_local)), np.zeros((3, n_local)), np.zeros(2)
    _check(self.lib.pb_nbody_allpairs_get_state(self.handle, _dptr(q),
                                                _dptr(v), _dptr(time)))
    return q, v, time

  def accelerations(self):
    a = np.zeros((3, self.i_end - self.i_begin))
    self._checked(self.lib.pb_nbody_allpairs_accelerations(self.handle,
                                                           _dptr(a)))
    return a

  def last_force_ms(self):
    return self.lib.pb_nbody_allpairs_last_force_ms(self.handle)

  def last_exchange_ms(self):
    return self.lib.pb_nbody_allpairs_last_exchange_ms(self.handle)
