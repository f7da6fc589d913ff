"""Python mirror of the reference's physics::Ephemeris surface for the hot path
(physics/ephemeris.hpp:67-561), over the C ABI of libprincipia_b200.so.

Method names follow the reference (snake_case); arguments are numpy arrays in
the SoA layouts documented in include/principia_b200.h.  Everything computes
on the GPU; there is no CPU path.
"""
import ctypes

import numpy as np

from . import _capi

_vp = ctypes.c_void_p


class StatusError(RuntimeError):

  def __init__(self, code, message):
    super().__init__('pb_status %d: %s' % (code, message))
    self.code = code


def _check(code):
  if code != _capi.PB_OK:
    raise StatusError(code, _capi.last_error())


def _check_flow(code):
  """A flow cut short by request_stop is not an error: its outputs describe the
  last completed step.  Returns whether the flow was cancelled."""
  if code == _capi.PB_CANCELLED:
    return True
  _check(code)
  return False


def _dptr(array):
  return array.ctypes.data_as(_capi.c_double_p)


class Ephemeris:
  """Ephemeris<Frame>: owns the massive bodies and their ContinuousTrajectory
  segments on one GPU."""

  def __init__(self, system, geopotential_tolerance=2.0**-24, device=0):
    self.lib = _capi.library()
    self.system = system
    self.n_bodies = len(system.bodies)
    bodies, self._keep = system.c_bodies()
    handle = _vp()
    _check(self.lib.pb_ephemeris_create(self.n_bodies, bodies,
                                        geopotential_tolerance, device,
                                        ctypes.byref(handle)))
    self.handle = handle
    self.device = device

  def close(self):
    if getattr(self, 'handle', None):
      self.lib.pb_ephemeris_destroy(self.handle)
      self.handle = None

  def __del__(self):
    self.close()

  # -- structure ------------------------------------------------------------
  def internal_order(self):
    order = np.zeros(self.n_bodies, dtype=np.int32)
    n_oblate = ctypes.c_int32()
    _check(self.lib.pb_ephemeris_internal_order(
        self.handle, order.ctypes.data_as(_capi.c_int32_p),
        ctypes.byref(n_oblate)))
    return order, n_oblate.value

  def geopotential_damping(self, body):
    n = ctypes.c_int32()
    inner = np.zeros(51)
    sectoral = ctypes.c_double()
    _check(self.lib.pb_geopotential_damping(self.handle, body, ctypes.byref(n),
                                            _dptr(inner),
                                            ctypes.byref(sectoral)))
    return inner[:n.value].copy(), sectoral.value

  # -- ContinuousTrajectory -------------------------------------------------
  def append_segments(self, body, segments):
    t_max = np.ascontiguousarray(segments['t_max'], dtype=np.float64)
    origin = np.ascontiguousarray(segments['origin'], dtype=np.float64)
    degree = np.ascontiguousarray(segments['degree'], dtype=np.int32)
    coefficients = np.ascontiguousarray(segments['coefficients'],
                                        dtype=np.float64)
    _check(self.lib.pb_ephemeris_append_segments(
        self.handle, body, len(t_max), float(segments['t_min']), _dptr(t_max),
        _dptr(origin), degree.ctypes.data_as(_capi.c_int32_p),
        _dptr(coefficients)))

  def t_min(self):
    return self.lib.pb_ephemeris_t_min(self.handle)

  def t_max(self):
    return self.lib.pb_ephemeris_t_max(self.handle)

  def evaluate_all_positions(self, t):
    out = np.zeros((self.n_bodies, 3))
    _check(self.lib.pb_ephemeris_evaluate_all_positions(self.handle, t,
                                                        _dptr(out)))
    return out

  # -- accelerations --------------------------------------------------------
  def compute_gravitational_acceleration_on_massless_bodies(self, t, q):
    """q: (3, n) positions.  Returns ((3, n) accelerations, status)."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    a = np.zeros_like(q)
    status = ctypes.c_int32()
    _check(self.lib.pb_compute_gravitational_acceleration_on_massless_bodies(
        self.handle, t, q.shape[1], _dptr(q), _dptr(a), ctypes.byref(status)))
    return a, status.value

  def compute_gravitational_acceleration_on_massless_bodies_device(
      self, t, n, q_pointer, a_pointer, status_pointer=None, stream=None):
    _check(
        self.lib.pb_compute_gravitational_acceleration_on_massless_bodies_device(
            self.handle, t, n, q_pointer, a_pointer, status_pointer, stream))

  def compute_gravitational_potential(self, t, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.zeros(q.shape[1])
    _check(self.lib.pb_compute_gravitational_potential(
        self.handle, t, q.shape[1], _dptr(q), _dptr(out)))
    return out

  def compute_gravitational_acceleration_on_massive_bodies(self, t,
                                                           positions=None):
    out = np.zeros((self.n_bodies, 3))
    pointer = None
    if positions is not None:
      positions = np.ascontiguousarray(positions, dtype=np.float64)
      pointer = _dptr(positions)
    _check(self.lib.pb_compute_gravitational_acceleration_on_massive_bodies(
        self.handle, t, pointer, _dptr(out)))
    return out

  def evaluate_all_velocities(self, t):
    out = np.zeros((self.n_bodies, 3))
    _check(self.lib.pb_ephemeris_evaluate_all_velocities(self.handle, t,
                                                         _dptr(out)))
    return out

  def compute_gravitational_jerk_on_massless_bodies(self, t, q, v):
    """q, v: (3, n) degrees of freedom.  Returns the (3, n) jerks
    (Ephemeris::ComputeGravitationalJerkOnMasslessBody, batched)."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    assert q.shape == v.shape
    out = np.zeros_like(q)
    _check(self.lib.pb_compute_gravitational_jerk_on_massless_bodies(
        self.handle, t, q.shape[1], _dptr(q), _dptr(v), _dptr(out)))
    return out

  def compute_gravitational_jerk_on_massless_bodies_device(
      self, t, n, q_pointer, v_pointer, jerk_pointer, stream=None):
    _check(self.lib.pb_compute_gravitational_jerk_on_massless_bodies_device(
        self.handle, t, n, q_pointer, v_pointer, jerk_pointer, stream))

  def _jerk_and_jacobian(self, t, positions, velocities, want_jerks,
                         want_jacobians):
    jerks = np.zeros((self.n_bodies, 3)) if want_jerks else None
    jacobians = np.zeros((self.n_bodies, 3, 3)) if want_jacobians else None
    p = v = None
    if positions is not None:
      positions = np.ascontiguousarray(positions, dtype=np.float64)
      p = _dptr(positions)
      if velocities is not None:
        velocities = np.ascontiguousarray(velocities, dtype=np.float64)
        v = _dptr(velocities)
    _check(
        self.lib.pb_compute_gravitational_jerk_and_jacobian_on_massive_bodies(
            self.handle, t, p, v, None if jerks is None else _dptr(jerks),
            None if jacobians is None else _dptr(jacobians)))
    return jerks, jacobians

  def compute_gravitational_jerk_on_massive_bodies(self, t, positions=None,
                                                   velocities=None):
    """Ephemeris::ComputeGravitationalJerkOnMassiveBody for every body: at t
    from the trajectories, or on the given degrees of freedom ((n, 3), caller
    order)."""
    return self._jerk_and_jacobian(t, positions, velocities, True, False)[0]

  def compute_jacobian_on_massive_bodies(self, t, positions=None):
    """Ephemeris::ComputeJacobianOnMassiveBody for every body: (n, 3, 3)."""
    return self._jerk_and_jacobian(t, positions, None, False, True)[1]

  # -- flows ----------------------------------------------------------------
  def _fixed_step_flow(self, method, step, t_final, n, state_pointer, time,
                       sample_every, max_samples, samples_pointer,
                       sample_times, arithmetic=_capi.PB_ARITHMETIC_EXACT):
    flow = _capi.FixedStepFlow()
    flow.arithmetic = arithmetic
    flow.method = method
    flow.step = step
    flow.t_final = t_final
    flow.n = n
    flow.state = state_pointer
    flow.time = _dptr(time)
    flow.sample_every = sample_every
    flow.max_samples = max_samples
    flow.samples = samples_pointer
    flow.sample_times = None if sample_times is None else _dptr(sample_times)
    outs = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
    flow.n_samples = ctypes.pointer(outs[0])
    flow.n_steps = ctypes.pointer(outs[1])
    flow.status = ctypes.pointer(outs[2])
    return flow, outs

  def flow_with_fixed_step(self, method, step, t_final, state, time,
                           sample_every=0, max_samples=0, burns=None,
                           arithmetic=_capi.PB_ARITHMETIC_EXACT):
    """NewInstance + FlowWithFixedStep on n probes.  state: (12, n) float64,
    time: (2,) float64; both updated in place (host buffers).  burns: the
    intrinsic accelerations, tabulated (8, n), or None.  arithmetic:
    PB_ARITHMETIC_EXACT (bit-identical to the reference's arithmetic) or
    PB_ARITHMETIC_CONTRACTED (fused multiply-adds, within 1e-12)."""
    assert state.flags.c_contiguous and state.dtype == np.float64
    assert time.flags.c_contiguous and time.dtype == np.float64
    assert state.shape[0] == 12 and time.shape == (2,)
    n = state.shape[1]
    samples = sample_times = None
    if sample_every > 0 and max_samples > 0:
      samples = np.zeros((max_samples, 6, n))
      sample_times = np.zeros(max_samples)
    flow, outs = self._fixed_step_flow(
        method, step, t_final, n, state.ctypes.data, time, sample_every,
        max_samples, None if samples is None else samples.ctypes.data,
        sample_times, arithmetic)
    if burns is not None:
      burns = np.ascontiguousarray(burns, dtype=np.float64)
      assert burns.shape == (8, n)
      flow.burns = burns.ctypes.data
    cancelled = _check_flow(
        self.lib.pb_flow_with_fixed_step(self.handle, ctypes.byref(flow)))
    n_samples, n_steps, status = (o.value for o in outs)
    return dict(n_steps=n_steps, status=status, cancelled=cancelled,
                samples=None if samples is None else samples[:n_samples],
                sample_times=None if samples is None
                else sample_times[:n_samples])

  def flow_with_fixed_step_device(self, method, step, t_final, n,
                                  state_pointer, time, stream=None,
                                  arithmetic=_capi.PB_ARITHMETIC_EXACT):
    """Same, state already resident in HBM (device pointer to 12 * n
    doubles)."""
    flow, outs = self._fixed_step_flow(method, step, t_final, n, state_pointer,
                                       time, 0, 0, None, None, arithmetic)
    cancelled = _check_flow(self.lib.pb_flow_with_fixed_step_device(
        self.handle, ctypes.byref(flow), stream))
    return dict(n_steps=outs[1].value, status=outs[2].value,
                cancelled=cancelled)

  def new_instance(self, method, step, state, time):
    """Ephemeris::NewInstance over n massless trajectories (state (12, n),
    time (2,)); the instance keeps its state on the device."""
    return FixedStepInstance(self, method, step, state, time)

  def last_flow_kernel_ms(self):
    """(device ms, launches) of the flow kernels of the last flow call."""
    launches = ctypes.c_int64()
    ms = self.lib.pb_ephemeris_last_flow_kernel_ms(self.handle,
                                                   ctypes.byref(launches))
    return ms, launches.value

  def last_flow_path(self):
    """(hot body (caller index, -1: none), whether some probe took the slow
    generic harmonics) of the last fixed-step flow."""
    hot, generic = ctypes.c_int32(), ctypes.c_int32()
    _check(self.lib.pb_ephemeris_last_flow_path(self.handle, ctypes.byref(hot),
                                                ctypes.byref(generic)))
    return hot.value, bool(generic.value)

  def request_stop(self):
    """RETURN_IF_STOPPED: the flows in progress return PB_CANCELLED with the
    state of their last completed step.  Any thread."""
    _check(self.lib.pb_ephemeris_request_stop(self.handle))

  def clear_stop(self):
    _check(self.lib.pb_ephemeris_clear_stop(self.handle))

  def flow_with_adaptive_step(self, t_final, state, time, length_tolerance,
                              speed_tolerance, max_steps=2**62,
                              step_log_capacity=0, sink=None, burns=None,
                              method=_capi.DORMAND_ELMIKKAWY_PRINCE_1986_RKN_434FM,
                              frenet_centres=None):
    """FlowWithAdaptiveStep, one trajectory per probe.  state: (12, n),
    time: (2, n); updated in place (host buffers).  frenet_centres: (n,) body
    indices (-1: inertially fixed) for burns given in the Frenet frame of the
    body-centred non-rotating frame of that body (needs FINE_1987_RKNG_34)."""
    assert state.flags.c_contiguous and time.flags.c_contiguous
    assert state.dtype == np.float64 and time.dtype == np.float64
    assert state.shape[0] == 12 and time.shape == (2, state.shape[1])
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
      # (8, n): direction xyz, thrust, initial mass, mass flow, initial and
      # final time of an inertially fixed burn per probe.
      burns = np.ascontiguousarray(burns, dtype=np.float64)
      assert burns.shape == (8, n)
      flow.burns = burns.ctypes.data
    if frenet_centres is not None:
      frenet_centres = np.ascontiguousarray(frenet_centres, dtype=np.int32)
      assert frenet_centres.shape == (n,) and burns is not None
      flow.frenet_centres = frenet_centres.ctypes.data
    if sink is not None:
      # A trajectory.Downsampler: every accepted state is appended to it inside
      # the kernel.
      flow.sink = sink.handle
    cancelled = _check_flow(
        self.lib.pb_flow_with_adaptive_step(self.handle, ctypes.byref(flow)))
    return dict(status=status, accepted=accepted, rejected=rejected,
                evaluations=evaluations, step_log=step_log,
                cancelled=cancelled)


class FixedStepInstance:
  """Integrator<NewtonianMotionEquation>::Instance of the reference for a
  fixed-step method (integrators/integrators.hpp:40-94): flow(t) is
  Ephemeris::FlowWithFixedStep(t, instance)."""

  @classmethod
  def from_checkpoint(cls, ephemeris, checkpoint):
    """Instance::ReadFromMessage."""
    checkpoint = np.ascontiguousarray(checkpoint, dtype=np.float64)
    self = cls.__new__(cls)
    self.ephemeris = ephemeris
    self.lib = ephemeris.lib
    self.handle = ctypes.c_void_p()
    _check(self.lib.pb_fixed_step_instance_read_checkpoint(
        ephemeris.handle, _dptr(checkpoint), len(checkpoint),
        ctypes.byref(self.handle)))
    self.n = int(checkpoint[3])
    return self

  def checkpoint(self):
    """Instance::WriteToMessage: a flat array of doubles."""
    size = self.lib.pb_fixed_step_instance_checkpoint_size(self.handle)
    out = np.zeros(size)
    _check(self.lib.pb_fixed_step_instance_write_checkpoint(self.handle,
                                                            _dptr(out)))
    return out

  def __init__(self, ephemeris, method, step, state, time):
    self.ephemeris = ephemeris  # Keeps the ephemeris alive.
    self.lib = ephemeris.lib
    self.n = state.shape[1]
    state = np.ascontiguousarray(state, dtype=np.float64)
    time = np.ascontiguousarray(time, dtype=np.float64)
    self.handle = ctypes.c_void_p()
    _check(self.lib.pb_fixed_step_instance_create(
        ephemeris.handle, method, step, self.n, _dptr(state), _dptr(time),
        ctypes.byref(self.handle)))

  def __del__(self):
    if getattr(self, 'handle', None):
      self.lib.pb_fixed_step_instance_destroy(self.handle)
      self.handle = None

  def set_arithmetic(self, arithmetic):
    """PB_ARITHMETIC_EXACT (default) or PB_ARITHMETIC_CONTRACTED."""
    _check(self.lib.pb_fixed_step_instance_set_arithmetic(self.handle,
                                                          arithmetic))

  def set_intrinsic_accelerations(self, burns):
    """The intrinsic_accelerations of NewInstance, tabulated: (8, n) or None."""
    if burns is not None:
      burns = np.ascontiguousarray(burns, dtype=np.float64)
      assert burns.shape == (8, self.n)
    _check(self.lib.pb_fixed_step_instance_set_intrinsic_accelerations(
        self.handle, None if burns is None else _dptr(burns)))

  def flow(self, t_final, sample_every=0, max_samples=0):
    samples = sample_times = None
    if sample_every > 0 and max_samples > 0:
      samples = np.zeros((max_samples, 6, self.n))
      sample_times = np.zeros(max_samples)
    n_steps, n_samples = ctypes.c_int64(), ctypes.c_int64()
    status = ctypes.c_int32()
    cancelled = _check_flow(self.lib.pb_fixed_step_instance_flow(
        self.handle, t_final, sample_every, max_samples,
        None if samples is None else _dptr(samples),
        None if samples is None else _dptr(sample_times),
        ctypes.byref(n_steps), ctypes.byref(n_samples), ctypes.byref(status)))
    return dict(n_steps=n_steps.value, status=status.value,
                cancelled=cancelled,
                samples=None if samples is None
                else samples[:n_samples.value],
                sample_times=None if samples is None
                else sample_times[:n_samples.value])

  def flow_to_downsampler(self, t_final, downsampler):
    """flow(t_final) with every appended state going to a
    trajectory.Downsampler on the device."""
    n_steps = ctypes.c_int64()
    status = ctypes.c_int32()
    _check(self.lib.pb_fixed_step_instance_flow_to_downsampler(
        self.handle, t_final, downsampler.handle, ctypes.byref(n_steps),
        ctypes.byref(status)))
    return dict(n_steps=n_steps.value, status=status.value)

  def state_into(self, state, time):
    """The instance's state into caller-provided (12, n) and (2,) arrays."""
    assert state.flags.c_contiguous and state.shape == (12, self.n)
    _check(self.lib.pb_fixed_step_instance_get_state(self.handle, _dptr(state),
                                                     _dptr(time)))

  def close(self):
    if getattr(self, 'handle', None):
      self.lib.pb_fixed_step_instance_destroy(self.handle)
      self.handle = None

  def state(self):
    """(state (12, n), time (2,)) of the instance."""
    state = np.zeros((12, self.n))
    time = np.zeros(2)
    _check(self.lib.pb_fixed_step_instance_get_state(self.handle, _dptr(state),
                                                     _dptr(time)))
    return state, time


def measure_fp64_peak(device=0):
  flops = ctypes.c_double()
  ms = ctypes.c_double()
  _check(_capi.library().pb_measure_fp64_peak(device, ctypes.byref(flops),
                                              ctypes.byref(ms)))
  return flops.value, ms.value


def measure_sm_clock(device=0, microseconds=200.0):
  """SM clock in MHz under the current load, measured on the device."""
  mhz = ctypes.c_double()
  _check(_capi.library().pb_measure_sm_clock(device, microseconds,
                                             ctypes.byref(mhz)))
  return mhz.value


def sm_clock_probe_start(device=0, delay_microseconds=5000.0,
                         microseconds=1000.0):
  """Enqueues a device-side SM clock measurement over `microseconds`, starting
  `delay_microseconds` from now; returns at once."""
  _check(_capi.library().pb_sm_clock_probe_start(device, delay_microseconds,
                                                 microseconds))


def sm_clock_probe_read(device=0):
  """MHz measured by the probe started last on this thread."""
  mhz = ctypes.c_double()
  _check(_capi.library().pb_sm_clock_probe_read(device, ctypes.byref(mhz)))
  return mhz.value


def kernel_launch_count():
  return _capi.library().pb_kernel_launch_count()
