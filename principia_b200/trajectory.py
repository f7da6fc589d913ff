"""The trajectory sink of the flows: DiscreteTrajectorySegment with
downsampling (physics/discrete_trajectory_segment.hpp of the reference) for n
trajectories on the GPU, through the C ABI."""
import ctypes

import numpy as np

from . import _capi
from .ephemeris import _check, _dptr


class Downsampler:
  """n timelines of at most `capacity` points; SetDownsampling({tolerance})
  (a negative tolerance keeps every point)."""

  def __init__(self, n, tolerance, capacity, device=0):
    self.lib = _capi.library()
    self.n = n
    self.capacity = capacity
    self.handle = ctypes.c_void_p()
    _check(self.lib.pb_downsampler_create(device, n, tolerance, capacity,
                                          ctypes.byref(self.handle)))

  def __del__(self):
    if getattr(self, 'handle', None):
      self.lib.pb_downsampler_destroy(self.handle)
      self.handle = None

  def append(self, sample_times, samples):
    """samples: (n_samples, 6, n) host array, as returned by the fixed-step
    flows; sample_times: (n_samples,)."""
    sample_times = np.ascontiguousarray(sample_times, dtype=np.float64)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    assert samples.shape == (len(sample_times), 6, self.n)
    _check(self.lib.pb_downsampler_append(self.handle, len(sample_times),
                                          _dptr(sample_times), _dptr(samples)))

  def append_logs_device(self, log_capacity, step_log_pointer,
                         accepted_steps_pointer, state_log_pointer,
                         stream=None):
    """The device-resident logs of an adaptive flow."""
    _check(self.lib.pb_downsampler_append_logs_device(
        self.handle, log_capacity, step_log_pointer, accepted_steps_pointer,
        state_log_pointer, stream))

  def timelines(self):
    """(counts (n,), times (capacity, n), points (capacity, 6, n),
    errors (capacity, n))."""
    counts = np.zeros(self.n, dtype=np.int64)
    times = np.zeros((self.capacity, self.n))
    points = np.zeros((self.capacity, 6, self.n))
    errors = np.zeros((self.capacity, self.n))
    _check(self.lib.pb_downsampler_get(
        self.handle, counts.ctypes.data_as(_capi.c_int64_p), _dptr(times),
        _dptr(points), _dptr(errors)))
    return counts, times, points, errors

  def counts(self):
    """The number of points of every timeline, (n,)."""
    counts = np.zeros(self.n, dtype=np.int64)
    _check(self.lib.pb_downsampler_get(
        self.handle, counts.ctypes.data_as(_capi.c_int64_p), None, None, None))
    return counts

  def evaluate(self, t):
    """DiscreteTrajectory::EvaluatePosition / EvaluateVelocity of every
    timeline at t: ((3, n) positions, (3, n) velocities, (n,) in_range)."""
    q = np.zeros((3, self.n))
    v = np.zeros((3, self.n))
    in_range = np.zeros(self.n, dtype=np.int32)
    _check(self.lib.pb_downsampler_evaluate(
        self.handle, t, _dptr(q), _dptr(v),
        in_range.ctypes.data_as(_capi.c_int32_p)))
    return q, v, in_range.astype(bool)

  def evaluate_device(self, t, q_pointer, v_pointer, in_range_pointer=None,
                      stream=None):
    _check(self.lib.pb_downsampler_evaluate_device(
        self.handle, t, q_pointer, v_pointer, in_range_pointer, stream))
