"""ctypes declarations of include/principia_b200.h and the loader of the CUDA
library.  There is no CPU fallback: if `libprincipia_b200.so` is missing the
import of the product fails loudly (build it with `__graft_entry__.build()` or
`make -C principia_b200/csrc`).
"""
import ctypes
import os

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)

PB_OK = 0
PB_CANCELLED = 1
PB_INVALID_ARGUMENT = 3
PB_DEADLINE_EXCEEDED = 4
PB_RESOURCE_EXHAUSTED = 8
PB_FAILED_PRECONDITION = 9
PB_ABORTED = 10
PB_OUT_OF_RANGE = 11
PB_INTERNAL = 13

BLANES_MOAN_2002_SRKN_14A = 0
MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL = 1
QUINLAN_1999_ORDER_8A = 2
QUINLAN_TREMAINE_1990_ORDER_12 = 3
DORMAND_ELMIKKAWY_PRINCE_1986_RKN_434FM = 16
FINE_1987_RKNG_34 = 17


class Body(ctypes.Structure):
  """pb_body_t."""
  _fields_ = [
      ('gravitational_parameter', ctypes.c_double),
      ('min_radius', ctypes.c_double),
      ('mean_radius', ctypes.c_double),
      ('reference_angle', ctypes.c_double),
      ('reference_instant', ctypes.c_double),
      ('angular_frequency', ctypes.c_double),
      ('right_ascension_of_pole', ctypes.c_double),
      ('declination_of_pole', ctypes.c_double),
      ('reference_radius', ctypes.c_double),
      ('is_rotating', ctypes.c_int32),
      ('is_oblate', ctypes.c_int32),
      ('geopotential_degree', ctypes.c_int32),
      ('is_zonal', ctypes.c_int32),
      ('cos', c_double_p),
      ('sin', c_double_p),
  ]


class FixedStepFlow(ctypes.Structure):
  """pb_fixed_step_flow_t."""
  _fields_ = [
      ('method', ctypes.c_int32),
      ('step', ctypes.c_double),
      ('t_final', ctypes.c_double),
      ('n', ctypes.c_int64),
      ('state', ctypes.c_void_p),
      ('time', c_double_p),
      ('sample_every', ctypes.c_int64),
      ('max_samples', ctypes.c_int64),
      ('samples', ctypes.c_void_p),
      ('sample_times', c_double_p),
      ('n_samples', c_int64_p),
      ('n_steps', c_int64_p),
      ('status', c_int32_p),
      ('burns', ctypes.c_void_p),
      ('arithmetic', ctypes.c_int32),
  ]


class AdaptiveStepFlow(ctypes.Structure):
  """pb_adaptive_step_flow_t."""
  _fields_ = [
      ('method', ctypes.c_int32),
      ('t_final', ctypes.c_double),
      ('length_integration_tolerance', ctypes.c_double),
      ('speed_integration_tolerance', ctypes.c_double),
      ('max_steps', ctypes.c_int64),
      ('n', ctypes.c_int64),
      ('state', ctypes.c_void_p),
      ('time', ctypes.c_void_p),
      ('status', ctypes.c_void_p),
      ('accepted_steps', ctypes.c_void_p),
      ('rejected_steps', ctypes.c_void_p),
      ('evaluations', ctypes.c_void_p),
      ('step_log_capacity', ctypes.c_int64),
      ('step_log', ctypes.c_void_p),
      ('state_log', ctypes.c_void_p),
      ('sink', ctypes.c_void_p),
      ('burns', ctypes.c_void_p),
      ('frenet_centres', ctypes.c_void_p),
  ]


PB_ARITHMETIC_EXACT = 0
PB_ARITHMETIC_CONTRACTED = 1

EXCHANGE_POSITIONS = ctypes.CFUNCTYPE(ctypes.c_int32, ctypes.c_void_p,
                                      ctypes.c_void_p,
                                      ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_void_p)

# PB_LIBRARY_PATH: another build of the same library (kernel experiments,
# tools/flow_variants.sh); there is still no fallback of any kind.
LIBRARY_PATH = os.environ.get('PB_LIBRARY_PATH') or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), 'libprincipia_b200.so')

# Every symbol include/principia_b200.h declares: (name, restype, argtypes).
_d, _i32, _i64, _vp = (ctypes.c_double, ctypes.c_int32, ctypes.c_int64,
                       ctypes.c_void_p)
SYMBOLS = [
    ('pb_ephemeris_create', _i32,
     [_i32, ctypes.POINTER(Body), _d, _i32, ctypes.POINTER(_vp)]),
    ('pb_ephemeris_destroy', None, [_vp]),
    ('pb_last_error', ctypes.c_char_p, []),
    ('pb_ephemeris_internal_order', _i32, [_vp, c_int32_p, c_int32_p]),
    ('pb_geopotential_damping', _i32,
     [_vp, _i32, c_int32_p, c_double_p, c_double_p]),
    ('pb_ephemeris_append_segments', _i32,
     [_vp, _i32, _i32, _d, c_double_p, c_double_p, c_int32_p, c_double_p]),
    ('pb_ephemeris_t_min', _d, [_vp]),
    ('pb_ephemeris_t_max', _d, [_vp]),
    ('pb_ephemeris_evaluate_all_positions', _i32, [_vp, _d, c_double_p]),
    ('pb_compute_gravitational_acceleration_on_massless_bodies', _i32,
     [_vp, _d, _i64, c_double_p, c_double_p, c_int32_p]),
    ('pb_compute_gravitational_acceleration_on_massless_bodies_device', _i32,
     [_vp, _d, _i64, _vp, _vp, _vp, _vp]),
    ('pb_compute_gravitational_potential', _i32,
     [_vp, _d, _i64, c_double_p, c_double_p]),
    ('pb_compute_gravitational_acceleration_on_massive_bodies', _i32,
     [_vp, _d, c_double_p, c_double_p]),
    ('pb_ephemeris_evaluate_all_velocities', _i32, [_vp, _d, c_double_p]),
    ('pb_compute_gravitational_jerk_on_massless_bodies', _i32,
     [_vp, _d, _i64, c_double_p, c_double_p, c_double_p]),
    ('pb_compute_gravitational_jerk_on_massless_bodies_device', _i32,
     [_vp, _d, _i64, _vp, _vp, _vp, _vp]),
    ('pb_compute_gravitational_jerk_and_jacobian_on_massive_bodies', _i32,
     [_vp, _d, c_double_p, c_double_p, c_double_p, c_double_p]),
    ('pb_continuous_trajectories_create', _i32,
     [_i32, _i64, _d, _d, _i64, ctypes.POINTER(_vp)]),
    ('pb_continuous_trajectories_destroy', None, [_vp]),
    ('pb_continuous_trajectories_append', _i32,
     [_vp, _d, c_double_p, c_double_p]),
    ('pb_continuous_trajectories_append_device', _i32,
     [_vp, _d, _vp, _vp, _vp]),
    ('pb_continuous_trajectories_status', _i32, [_vp]),
    ('pb_continuous_trajectories_t_min', _d, [_vp]),
    ('pb_continuous_trajectories_t_max', _d, [_vp]),
    ('pb_continuous_trajectories_segment_count', _i64, [_vp]),
    ('pb_continuous_trajectories_get_segments', _i32,
     [_vp, _i64, c_double_p, c_double_p, c_double_p, c_int32_p, c_double_p]),
    ('pb_continuous_trajectories_evaluate', _i32,
     [_vp, _d, c_double_p, c_double_p]),
    ('pb_continuous_trajectories_evaluate_device', _i32,
     [_vp, _d, _vp, _vp, _vp]),
    ('pb_nbody_ensemble_solve_to_trajectories', _i32,
     [_vp, _d, _vp, ctypes.POINTER(_i64)]),
    ('pb_flow_with_fixed_step', _i32, [_vp, ctypes.POINTER(FixedStepFlow)]),
    ('pb_flow_with_fixed_step_device', _i32,
     [_vp, ctypes.POINTER(FixedStepFlow), _vp]),
    ('pb_fixed_step_instance_create', _i32,
     [_vp, _i32, _d, _i64, c_double_p, c_double_p, ctypes.POINTER(_vp)]),
    ('pb_fixed_step_instance_destroy', None, [_vp]),
    ('pb_fixed_step_instance_set_intrinsic_accelerations', _i32,
     [_vp, c_double_p]),
    ('pb_fixed_step_instance_set_arithmetic', _i32, [_vp, _i32]),
    ('pb_fixed_step_instance_flow', _i32,
     [_vp, _d, _i64, _i64, c_double_p, c_double_p, c_int64_p, c_int64_p,
      c_int32_p]),
    ('pb_fixed_step_instance_flow_to_downsampler', _i32,
     [_vp, _d, _vp, c_int64_p, c_int32_p]),
    ('pb_fixed_step_instance_checkpoint_size', _i64, [_vp]),
    ('pb_fixed_step_instance_write_checkpoint', _i32, [_vp, c_double_p]),
    ('pb_fixed_step_instance_read_checkpoint', _i32,
     [_vp, c_double_p, _i64, ctypes.POINTER(_vp)]),
    ('pb_fixed_step_instance_get_state', _i32, [_vp, c_double_p, c_double_p]),
    ('pb_flow_with_adaptive_step', _i32,
     [_vp, ctypes.POINTER(AdaptiveStepFlow)]),
    ('pb_flow_with_adaptive_step_device', _i32,
     [_vp, ctypes.POINTER(AdaptiveStepFlow), _vp]),
    ('pb_nbody_ensemble_create', _i32,
     [_vp, _i32, _d, _i64, c_double_p, _d, ctypes.POINTER(_vp)]),
    ('pb_nbody_ensemble_destroy', None, [_vp]),
    ('pb_nbody_ensemble_solve', _i32, [_vp, _d, c_int64_p]),
    ('pb_nbody_ensemble_solve_recording', _i32,
     [_vp, _d, _i64, c_double_p, c_double_p, c_int64_p]),
    ('pb_nbody_ensemble_get_state', _i32, [_vp, c_double_p, c_double_p]),
    ('pb_nbody_allpairs_create', _i32,
     [_i32, _i32, _d, _i64, _i64, _i64, c_double_p, c_double_p, c_double_p, _d,
      EXCHANGE_POSITIONS, _vp, ctypes.POINTER(_vp)]),
    ('pb_nbody_allpairs_destroy', None, [_vp]),
    ('pb_nbody_allpairs_solve', _i32, [_vp, _d, c_int64_p, c_int64_p]),
    ('pb_nbody_allpairs_get_state', _i32,
     [_vp, c_double_p, c_double_p, c_double_p]),
    ('pb_nbody_allpairs_accelerations', _i32, [_vp, c_double_p]),
    ('pb_nbody_allpairs_last_force_ms', _d, [_vp]),
    ('pb_nbody_allpairs_last_exchange_ms', _d, [_vp]),
    ('pb_nbody_allpairs_set_nccl', _i32, [_vp, _vp]),
    ('pb_nbody_allpairs_set_arithmetic', _i32, [_vp, _i32]),
    ('pb_nccl_get_unique_id', _i32, [ctypes.c_char_p]),
    ('pb_nccl_comm_create', _i32,
     [ctypes.c_char_p, _i32, _i32, _i32, ctypes.POINTER(_vp)]),
    ('pb_nccl_comm_destroy', None, [_vp]),
    ('pb_downsampler_create', _i32,
     [_i32, _i64, _d, _i64, ctypes.POINTER(_vp)]),
    ('pb_downsampler_destroy', None, [_vp]),
    ('pb_downsampler_append', _i32, [_vp, _i64, c_double_p, c_double_p]),
    ('pb_downsampler_append_device', _i32,
     [_vp, _i64, c_double_p, _vp, _vp]),
    ('pb_downsampler_append_logs_device', _i32,
     [_vp, _i64, _vp, _vp, _vp, _vp]),
    ('pb_downsampler_get', _i32,
     [_vp, c_int64_p, c_double_p, c_double_p, c_double_p]),
    ('pb_downsampler_evaluate', _i32,
     [_vp, _d, c_double_p, c_double_p, c_int32_p]),
    ('pb_downsampler_evaluate_device', _i32, [_vp, _d, _vp, _vp, _vp, _vp]),
    ('pb_chebyshev_series_evaluate', _i32,
     [_i32, _i64, _i32, c_int32_p, c_double_p, c_double_p, c_double_p, _i64,
      c_int32_p, c_double_p, c_double_p, c_double_p]),
    ('pb_ephemeris_last_flow_kernel_ms', _d, [_vp, c_int64_p]),
    ('pb_ephemeris_last_flow_path', _i32, [_vp, c_int32_p, c_int32_p]),
    ('pb_ephemeris_request_stop', _i32, [_vp]),
    ('pb_ephemeris_clear_stop', _i32, [_vp]),
    ('pb_measure_fp64_peak', _i32, [_i32, c_double_p, c_double_p]),
    ('pb_measure_sm_clock', _i32, [_i32, ctypes.c_double, c_double_p]),
    ('pb_sm_clock_probe_start', _i32,
     [_i32, ctypes.c_double, ctypes.c_double]),
    ('pb_sm_clock_probe_read', _i32, [_i32, c_double_p]),
    ('pb_kernel_launch_count', _i64, []),
]

_library = None


def library():
  """Loads libprincipia_b200.so (once) and types every entry point."""
  global _library
  if _library is None:
    if not os.path.exists(LIBRARY_PATH):
      raise ImportError(
          'principia_b200: %s is missing; there is no CPU fallback.  Build the '
          'CUDA library with `python -c "import __graft_entry__ as g; '
          'g.build()"` or `make -C principia_b200/csrc`.' % LIBRARY_PATH)
    lib = ctypes.CDLL(LIBRARY_PATH)
    for name, restype, argtypes in SYMBOLS:
      f = getattr(lib, name)
      f.restype = restype
      f.argtypes = argtypes
    _library = lib
  return _library


def last_error():
  return library().pb_last_error().decode('utf-8', 'replace')
