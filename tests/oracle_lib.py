"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.
"""
import ctypes
import os
import subprocess

import numpy as np

from principia_b200 import _capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, 'oracle')
ORACLE_PATH = os.path.join(ORACLE_DIR, 'liboracle.so')

_d, _i32, _i64, _vp = (ctypes.c_double, ctypes.c_int32, ctypes.c_int64,
                       ctypes.c_void_p)
_dp, _i32p, _i64p = _capi.c_double_p, _capi.c_int32_p, _capi.c_int64_p

_SYMBOLS = [
    ('orc_kat_fixed_step_oscillator', _i32,
     [_i32, _d, _d, _dp, _dp, _i64p, _i64p]),
    ('orc_oscillator_solve', _i32,
     [_i32, _d, _d, _d, _d, _d, _d, _d, _i64, _dp, _dp, _dp, _dp, _i64p,
      _i64p]),
    ('orc_kat_adaptive_oscillator', _i32,
     [_d, _d, _d, _d, _d, _d, _d, _d, _i64p, _i64p, _i64p, _i64p, _dp]),
    ('orc_kat_adaptive_oscillator_restart', _i32, [_d, _d, _d, _i64p, _dp]),
    ('orc_kat_adaptive_rocket', _i32, [_d, _d, _i64p, _dp, _dp]),
    ('orc_hermite3', None,
     [_d, _d, _dp, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _dp]),
    ('orc_harmonic_damping', None, [_d, _d, _d, _d, _dp, _dp, _dp]),
    ('orc_double_precision', None, [_i32, _dp, _dp, _dp]),
    ('orc_kat_best_newhall', _i32,
     [_d, _d, _i32, _i32p, _i32p, _dp, _i32p, _i32p, _dp, _i32p, _i32p]),
    ('orc_kat_legendre', _i32,
     [_d, _d, _i64, _dp, _dp, _dp, _i64p, _i64p, _i64p, _i64p]),
    ('orc_kat_adaptive_oscillator_max_steps', _i32,
     [_d, _d, _d, _d, _d, _d, _d, _d, _i64, _i64p, _i64p, _i64p, _i64p, _dp]),
    ('orc_cr_sin', _d, [_d]),
    ('orc_cr_cos', _d, [_d]),
    ('orc_root', _d, [_i32, _d]),
    ('orc_estrin', _d, [_dp, _i32, _d]),
    ('orc_estrin_derivative', _d, [_dp, _i32, _d]),
    ('orc_downsample', _i64,
     [_i64, _dp, _dp, _dp, _d, _dp, _dp, _dp, _dp, _i64, _dp, _dp, _dp]),
    ('orc_continuous_trajectory_create', _vp, [_d, _d]),
    ('orc_continuous_trajectory_destroy', None, [_vp]),
    ('orc_continuous_trajectory_append', _i32, [_vp, _d, _dp, _dp]),
    ('orc_continuous_trajectory_t_min', _d, [_vp]),
    ('orc_continuous_trajectory_t_max', _d, [_vp]),
    ('orc_continuous_trajectory_segment_count', _i32, [_vp]),
    ('orc_continuous_trajectory_degree', _i32, [_vp, _i32]),
    ('orc_continuous_trajectory_evaluate', None, [_vp, _d, _dp, _dp]),
    ('orc_continuous_trajectory_get_segment', None,
     [_vp, _i32, _dp, _dp, _i32p, _dp]),
    ('orc_chebyshev_evaluate', _d, [_dp, _i32, _d, _d, _d]),
    ('orc_chebyshev_evaluate_derivative', _d, [_dp, _i32, _d, _d, _d]),
    ('orc_newhall_fit', _i32,
     [_dp, _dp, _dp, _d, _d, _dp, _dp, _i32p, _dp]),
    ('orc_newhall_fixed_degree', _i32,
     [_i32, _dp, _dp, _d, _d, _dp, _i64, _dp, _dp, _dp]),
    ('orc_compute_gravitational_jerk_on_massless_bodies', _i32,
     [_vp, _d, _i64, _dp, _dp, _dp]),
    ('orc_compute_gravitational_jerk_on_massive_bodies', _i32,
     [_vp, _d, _dp, _dp, _dp]),
    ('orc_compute_jacobian_on_massive_bodies', _i32, [_vp, _d, _dp, _dp]),
    ('orc_ephemeris_create', _i32,
     [_i32, ctypes.POINTER(_capi.Body), _dp, _dp, _d, _d, _d, _i32, _d,
      ctypes.POINTER(_vp)]),
    ('orc_ephemeris_destroy', None, [_vp]),
    ('orc_ephemeris_internal_order', _i32, [_vp, _i32p, _i32p]),
    ('orc_geopotential_damping', _i32, [_vp, _i32, _i32p, _dp, _dp]),
    ('orc_geopotential_acceleration', _i32, [_vp, _i32, _d, _dp, _dp]),
    ('orc_surface_frame_rotate', _i32, [_vp, _i32, _d, _i32, _dp, _dp]),
    ('orc_geopotential_acceleration_explicit', _i32,
     [_vp, _i32, _d, _dp, _d, _d, _d, _dp]),
    ('orc_geopotential_potential', _d, [_vp, _i32, _d, _dp]),
    ('orc_ephemeris_prolong', _i32, [_vp, _d]),
    ('orc_ephemeris_instance_state', _i32, [_vp, _dp, _dp, _dp]),
    ('orc_ephemeris_segment_count', _i32, [_vp, _i32]),
    ('orc_ephemeris_get_segments', _i32,
     [_vp, _i32, _dp, _dp, _dp, _i32p, _dp]),
    ('orc_ephemeris_append_segments', _i32,
     [_vp, _i32, _i32, _d, _dp, _dp, _i32p, _dp]),
    ('orc_ephemeris_t_min', _d, [_vp]),
    ('orc_ephemeris_t_max', _d, [_vp]),
    ('orc_ephemeris_evaluate_all_positions', _i32, [_vp, _d, _dp]),
    ('orc_ephemeris_evaluate_all_velocities', _i32, [_vp, _d, _dp]),
    ('orc_compute_gravitational_acceleration_on_massless_bodies', _i32,
     [_vp, _d, _i64, _dp, _dp, _i32p]),
    ('orc_compute_gravitational_potential', _i32, [_vp, _d, _i64, _dp, _dp]),
    ('orc_compute_gravitational_acceleration_on_massive_bodies', _i32,
     [_vp, _d, _dp, _dp]),
    ('orc_compute_gravitational_acceleration_between_all_massive_bodies', _i32,
     [_vp, _d, _dp, _dp]),
    ('orc_flow_with_fixed_step', _i32,
     [_vp, ctypes.POINTER(_capi.FixedStepFlow), _i32]),
    ('orc_flow_with_adaptive_step', _i32,
     [_vp, ctypes.POINTER(_capi.AdaptiveStepFlow), _i32]),
    ('orc_nbody_ensemble_flow', _i32,
     [_vp, _i32, _d, _i64, _dp, _dp, _d, _i64p, _i32]),
    ('orc_allpairs_target_accelerations', _i32,
     [_i64, _dp, _dp, _i64, _i64p, _dp]),
    ('orc_max_threads', _i32, []),
]

_library = None


def build():
  subprocess.check_call(['make', '-s', '-C', ORACLE_DIR])


def library():
  global _library
  if _library is None:
    if not os.path.exists(ORACLE_PATH):
      build()
    lib = ctypes.CDLL(ORACLE_PATH)
    for name, restype, argtypes in _SYMBOLS:
      f = getattr(lib, name)
      f.restype = restype
      f.argtypes = argtypes
    _library = lib
  return _library


def dptr(a):
  return a.ctypes.data_as(_dp)


def allpairs_target_accelerations(mu, q, targets):
  """Accelerations (3, len(targets)) of the given targets of an all-pairs
  point-mass system (mu (n,), q (3, n)), in the product kernel's summation
  order: bit-comparable with NBodyAllPairs.accelerations()."""
  mu = np.ascontiguousarray(mu, dtype=np.float64)
  q = np.ascontiguousarray(q, dtype=np.float64)
  targets = np.ascontiguousarray(targets, dtype=np.int64)
  out = np.zeros((len(targets), 3))
  library().orc_allpairs_target_accelerations(
      len(mu), dptr(mu), dptr(q), len(targets),
      targets.ctypes.data_as(_i64p), dptr(out))
  return np.ascontiguousarray(out.T)


class OracleEphemeris:
  """The CPU restatement of physics::Ephemeris."""

  def __init__(self, system, method=_capi.QUINLAN_TREMAINE_1990_ORDER_12,
               step=600.0, fitting_tolerance=1e-3,
               geopotential_tolerance=2.0**-24, initial_positions=None,
               initial_velocities=None, initial_time=None):
    self.lib = library()
    self.system = system
    self.n = len(system.bodies)
    bodies, self._keep = system.c_bodies()
    q = np.ascontiguousarray(system.initial_positions() if initial_positions
                             is None else initial_positions, dtype=np.float64)
    v = np.ascontiguousarray(system.initial_velocities() if initial_velocities
                             is None else initial_velocities, dtype=np.float64)
    t0 = system.epoch if initial_time is None else initial_time
    handle = _vp()
    status = self.lib.orc_ephemeris_create(
        self.n, bodies, dptr(q), dptr(v), t0, fitting_tolerance,
        geopotential_tolerance, method, step, ctypes.byref(handle))
    assert status == 0
    self.handle = handle

  def __del__(self):
    if getattr(self, 'handle', None):
      self.lib.orc_ephemeris_destroy(self.handle)
      self.handle = None

  def internal_order(self):
    order = np.zeros(self.n, dtype=np.int32)
    n_oblate = _i32()
    self.lib.orc_ephemeris_internal_order(
        self.handle, order.ctypes.data_as(_i32p), ctypes.byref(n_oblate))
    return order, n_oblate.value

  def damping(self, body):
    n = _i32()
    inner = np.zeros(51)
    sectoral = _d()
    status = self.lib.orc_geopotential_damping(
        self.handle, body, ctypes.byref(n), dptr(inner), ctypes.byref(sectoral))
    assert status == 0
    return inner[:n.value].copy(), sectoral.value

  def geopotential_acceleration(self, body, t, r):
    r = np.ascontiguousarray(r, dtype=np.float64)
    out = np.zeros(3)
    self.lib.orc_geopotential_acceleration(self.handle, body, t, dptr(r),
                                           dptr(out))
    return out

  def geopotential_potential(self, body, t, r):
    r = np.ascontiguousarray(r, dtype=np.float64)
    return self.lib.orc_geopotential_potential(self.handle, body, t, dptr(r))

  def geopotential_acceleration_explicit(self, body, t, r, r_norm, r2,
                                         one_over_r3):
    r = np.ascontiguousarray(r, dtype=np.float64)
    out = np.zeros(3)
    self.lib.orc_geopotential_acceleration_explicit(
        self.handle, body, t, dptr(r), r_norm, r2, one_over_r3, dptr(out))
    return out

  def surface_frame_rotate(self, body, t, v, inverse=False):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(3)
    self.lib.orc_surface_frame_rotate(self.handle, body, t, int(inverse),
                                      dptr(v), dptr(out))
    return out

  def prolong(self, t):
    return self.lib.orc_ephemeris_prolong(self.handle, t)

  def instance_state(self):
    time = np.zeros(2)
    q = np.zeros(3 * self.n)
    v = np.zeros(3 * self.n)
    self.lib.orc_ephemeris_instance_state(self.handle, dptr(time), dptr(q),
                                          dptr(v))
    return time, q.reshape(self.n, 3), v.reshape(self.n, 3)

  def segments(self, body):
    count = self.lib.orc_ephemeris_segment_count(self.handle, body)
    t_min = _d()
    t_max = np.zeros(count)
    origin = np.zeros(count)
    degree = np.zeros(count, dtype=np.int32)
    coefficients = np.zeros((count, 18, 3))
    self.lib.orc_ephemeris_get_segments(
        self.handle, body, ctypes.byref(t_min), dptr(t_max), dptr(origin),
        degree.ctypes.data_as(_i32p), dptr(coefficients))
    return dict(t_min=t_min.value, t_max=t_max, origin=origin, degree=degree,
                coefficients=coefficients)

  def append_segments(self, body, seg):
    n = len(seg['t_max'])
    t_max = np.ascontiguousarray(seg['t_max'], dtype=np.float64)
    origin = np.ascontiguousarray(seg['origin'], dtype=np.float64)
    degree = np.ascontiguousarray(seg['degree'], dtype=np.int32)
    coefficients = np.ascontiguousarray(seg['coefficients'], dtype=np.float64)
    self.lib.orc_ephemeris_append_segments(
        self.handle, body, n, float(seg['t_min']), dptr(t_max), dptr(origin),
        degree.ctypes.data_as(_i32p), dptr(coefficients))

  def t_min(self):
    return self.lib.orc_ephemeris_t_min(self.handle)

  def t_max(self):
    return self.lib.orc_ephemeris_t_max(self.handle)

  def evaluate_all_positions(self, t):
    out = np.zeros((self.n, 3))
    self.lib.orc_ephemeris_evaluate_all_positions(self.handle, t, dptr(out))
    return out

  def evaluate_all_velocities(self, t):
    out = np.zeros((self.n, 3))
    self.lib.orc_ephemeris_evaluate_all_velocities(self.handle, t, dptr(out))
    return out

  def acceleration_on_massless_bodies(self, t, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    n = q.shape[1]
    a = np.zeros_like(q)
    status = _i32()
    self.lib.orc_compute_gravitational_acceleration_on_massless_bodies(
        self.handle, t, n, dptr(q), dptr(a), ctypes.byref(status))
    return a, status.value

  def potential(self, t, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    n = q.shape[1]
    out = np.zeros(n)
    self.lib.orc_compute_gravitational_potential(self.handle, t, n, dptr(q),
                                                 dptr(out))
    return out

  def acceleration_on_massive_bodies(self, t, positions=None):
    out = np.zeros((self.n, 3))
    if positions is None:
      p = None
    else:
      positions = np.ascontiguousarray(positions, dtype=np.float64)
      p = dptr(positions)
    self.lib.orc_compute_gravitational_acceleration_on_massive_bodies(
        self.handle, t, p, dptr(out))
    return out

  def jerk_on_massless_bodies(self, t, q, v):
    q = np.ascontiguousarray(q, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros_like(q)
    self.lib.orc_compute_gravitational_jerk_on_massless_bodies(
        self.handle, t, q.shape[1], dptr(q), dptr(v), dptr(out))
    return out

  def jerk_on_massive_bodies(self, t, positions=None, velocities=None):
    out = np.zeros((self.n, 3))
    p = v = None
    if positions is not None:
      positions = np.ascontiguousarray(positions, dtype=np.float64)
      velocities = np.ascontiguousarray(velocities, dtype=np.float64)
      p, v = dptr(positions), dptr(velocities)
    self.lib.orc_compute_gravitational_jerk_on_massive_bodies(
        self.handle, t, p, v, dptr(out))
    return out

  def jacobian_on_massive_bodies(self, t, positions=None):
    out = np.zeros((self.n, 3, 3))
    p = None
    if positions is not None:
      positions = np.ascontiguousarray(positions, dtype=np.float64)
      p = dptr(positions)
    self.lib.orc_compute_jacobian_on_massive_bodies(self.handle, t, p,
                                                    dptr(out))
    return out

  def acceleration_between_all_massive_bodies(self, t, positions_internal):
    positions = np.ascontiguousarray(positions_internal, dtype=np.float64)
    out = np.zeros_like(positions)
    self.lib.orc_compute_gravitational_acceleration_between_all_massive_bodies(
        self.handle, t, dptr(positions), dptr(out))
    return out

  def flow_with_fixed_step(self, method, step, t_final, state, time,
                           sample_every=0, max_samples=0, threads=0,
                           burns=None):
    """state: (12, n) array, time: (2,) array; both updated in place."""
    assert state.flags.c_contiguous and state.dtype == np.float64
    n = state.shape[1]
    flow = _capi.FixedStepFlow()
    flow.method = method
    flow.step = step
    flow.t_final = t_final
    flow.n = n
    flow.state = state.ctypes.data
    flow.time = dptr(time)
    samples = sample_times = None
    n_samples, n_steps, status = _i64(), _i64(), _i32()
    if sample_every > 0:
      samples = np.zeros((max_samples, 6, n))
      sample_times = np.zeros(max_samples)
      flow.sample_every = sample_every
      flow.max_samples = max_samples
      flow.samples = samples.ctypes.data
      flow.sample_times = dptr(sample_times)
    flow.n_samples = ctypes.pointer(n_samples)
    flow.n_steps = ctypes.pointer(n_steps)
    flow.status = ctypes.pointer(status)
    if burns is not None:
      burns = np.ascontiguousarray(burns, dtype=np.float64)
      assert burns.shape == (8, n)
      flow.burns = burns.ctypes.data
    self.lib.orc_flow_with_fixed_step(self.handle, ctypes.byref(flow), threads)
    return dict(n_steps=n_steps.value, status=status.value,
                samples=None if samples is None else samples[:n_samples.value],
                sample_times=None if samples is None
                else sample_times[:n_samples.value])

  def flow_with_adaptive_step(self, t_final, state, time, length_tolerance,
                              speed_tolerance, max_steps=2**62,
                              step_log_capacity=0, threads=0,
                              state_log=False, burns=None,
                              method=_capi.DORMAND_ELMIKKAWY_PRINCE_1986_RKN_434FM,
                              frenet_centres=None):
    """state: (12, n), time: (2, n); updated in place."""
    n = state.shape[1]
    flow = _capi.AdaptiveStepFlow()
    flow.method = method
    flow.t_final = t_final
    flow.length_integration_tolerance = length_tolerance
    flow.speed_integration_tolerance = speed_tolerance
    flow.max_steps = max_steps
    flow.n = n
    flow.state = state.ctypes.data
    flow.time = time.ctypes.data
    status = np.zeros(n, dtype=np.int32)
    accepted = np.zeros(n, dtype=np.int64)
    rejected = np.zeros(n, dtype=np.int64)
    evaluations = np.zeros(n, dtype=np.int64)
    flow.status = status.ctypes.data
    flow.accepted_steps = accepted.ctypes.data
    flow.rejected_steps = rejected.ctypes.data
    flow.evaluations = evaluations.ctypes.data
    step_log = None
    if step_log_capacity > 0:
      step_log = np.full((step_log_capacity, n), np.nan)
      flow.step_log_capacity = step_log_capacity
      flow.step_log = step_log.ctypes.data
    if burns is not None:
      burns = np.ascontiguousarray(burns, dtype=np.float64)
      assert burns.shape == (8, n)
      flow.burns = burns.ctypes.data
    if frenet_centres is not None:
      frenet_centres = np.ascontiguousarray(frenet_centres, dtype=np.int32)
      flow.frenet_centres = frenet_centres.ctypes.data
    states = None
    if state_log and step_log_capacity > 0:
      states = np.full((step_log_capacity, 6, n), np.nan)
      flow.state_log = states.ctypes.data
    self.lib.orc_flow_with_adaptive_step(self.handle, ctypes.byref(flow),
                                         threads)
    return dict(status=status, accepted=accepted, rejected=rejected,
                evaluations=evaluations, step_log=step_log, state_log=states)

  def nbody_ensemble_flow(self, method, step, state, time, t_final, threads=0):
    """state: (12, n_bodies, n_systems) internal order; in place."""
    n_systems = state.shape[2]
    n_steps = _i64()
    self.lib.orc_nbody_ensemble_flow(self.handle, method, step, n_systems,
                                     dptr(state), dptr(time), t_final,
                                     ctypes.byref(n_steps), threads)
    return n_steps.value
