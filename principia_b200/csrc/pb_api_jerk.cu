// C ABI: the jerk and Jacobian queries of physics::Ephemeris
// (include/principia_b200.h).
#include <cuda_runtime.h>

#include <vector>

#include "../../include/principia_b200.h"
#include "pb_kernels_jerk.cuh"
#include "pb_runtime.cuh"

using namespace pb;

namespace {

// Positions (from the stage table) and velocities of all the massive bodies
// at t, internal order, in w->stage_table[0 .. 3n) and w->scratch_c[0 .. 3n).
// The caller holds the workspace and the segments.
pb_status_t EvaluateDegreesOfFreedom(pb_ephemeris_t* e, FlowWorkspace* w,
                                     double t, cudaStream_t stream) {
  EphemerisView const view = e->View();
  int const n = view.n_bodies;
  PB_RETURN_IF_ERROR(w->stage_table.Reserve(view.stage_stride));
  PB_RETURN_IF_ERROR(BuildStageTable(e, w, &t, 1, w->stage_table.data, stream));
  PB_RETURN_IF_ERROR(w->scratch_c.Reserve(3 * n));
  VelocityTableKernel<<<(n + 63) / 64, 64, 0, stream>>>(e->View(), t,
                                                        w->scratch_c.data);
  PB_LAUNCHED();
  return PB_OK;
}

// ComputeGravitationalJerkOnMasslessBody on device arrays, asynchronous.
pb_status_t MasslessJerks(pb_ephemeris_t* e, FlowWorkspace* w, double t,
                          int64_t n, const double* q, const double* v,
                          double* jerk, cudaStream_t stream) {
  PB_RETURN_IF_ERROR(EvaluateDegreesOfFreedom(e, w, t, stream));
  if (n > 0) {
    MasslessJerkKernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0,
                         stream>>>(e->View(), w->stage_table.data,
                                   w->scratch_c.data, n, q, v, jerk);
    PB_LAUNCHED();
  }
  return PB_OK;
}

// Caller order (AoS, `width` doubles per body) <-> internal order.
void ToInternal(pb_ephemeris_t const* e, double const* caller, int width,
                std::vector<double>& internal) {
  int const n = static_cast<int>(e->model.bodies.size());
  internal.resize(static_cast<std::size_t>(width) * n);
  for (int b = 0; b < n; ++b) {
    int const c = e->model.order[b];
    for (int k = 0; k < width; ++k) {
      internal[width * b + k] = caller[width * c + k];
    }
  }
}

void ToCaller(pb_ephemeris_t const* e, std::vector<double> const& internal,
              int width, double* caller) {
  int const n = static_cast<int>(e->model.bodies.size());
  for (int b = 0; b < n; ++b) {
    int const c = e->model.order[b];
    for (int k = 0; k < width; ++k) {
      caller[width * c + k] = internal[width * b + k];
    }
  }
}

}  // namespace

extern "C" {

pb_status_t pb_ephemeris_evaluate_all_velocities(pb_ephemeris_t* e, double t,
                                                 double* velocities) {
  PB_API_RANGE();
  if (e == nullptr || velocities == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, false, nullptr));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  cudaStream_t const stream = w.stream();
  PB_RETURN_IF_ERROR(EvaluateDegreesOfFreedom(e, w.get(), t, stream));
  int const n = static_cast<int>(e->model.bodies.size());
  std::vector<double> internal(3 * n);
  PB_CUDA(cudaMemcpyAsync(internal.data(), w->scratch_c.data,
                          internal.size() * sizeof(double),
                          cudaMemcpyDeviceToHost, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  ToCaller(e, internal, 3, velocities);
  return PB_OK;
}

pb_status_t pb_compute_gravitational_jerk_on_massless_bodies_device(
    pb_ephemeris_t* e, double t, int64_t n, const double* q, const double* v,
    double* jerk, void* stream_pointer) {
  PB_API_RANGE();
  if (e == nullptr || n < 0 ||
      (n > 0 && (q == nullptr || v == nullptr || jerk == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = static_cast<cudaStream_t>(stream_pointer);
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, true, stream));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  return MasslessJerks(e, w.get(), t, n, q, v, jerk, stream);
}

pb_status_t pb_compute_gravitational_jerk_on_massless_bodies(
    pb_ephemeris_t* e, double t, int64_t n, const double* q, const double* v,
    double* jerk) {
  PB_API_RANGE();
  if (e == nullptr || n < 0 ||
      (n > 0 && (q == nullptr || v == nullptr || jerk == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, false, nullptr));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  cudaStream_t const stream = w.stream();
  // scratch_a: q then v; scratch_b: the jerks.
  PB_RETURN_IF_ERROR(w->scratch_a.Reserve(6 * n));
  PB_RETURN_IF_ERROR(w->scratch_b.Reserve(3 * n));
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(w->scratch_a.data, q, 3 * n * sizeof(double),
                            cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemcpyAsync(w->scratch_a.data + 3 * n, v,
                            3 * n * sizeof(double), cudaMemcpyHostToDevice,
                            stream));
  }
  PB_RETURN_IF_ERROR(MasslessJerks(e, w.get(), t, n, w->scratch_a.data,
                                   w->scratch_a.data + 3 * n,
                                   w->scratch_b.data, stream));
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(jerk, w->scratch_b.data, 3 * n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  return PB_OK;
}

pb_status_t pb_compute_gravitational_jerk_and_jacobian_on_massive_bodies(
    pb_ephemeris_t* e, double t, const double* positions,
    const double* velocities, double* jerks, double* jacobians) {
  PB_API_RANGE();
  if (e == nullptr || (jerks == nullptr && jacobians == nullptr) ||
      (positions != nullptr && jerks != nullptr && velocities == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, false, nullptr));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  cudaStream_t const stream = w.stream();
  EphemerisView const view = e->View();
  int const n = view.n_bodies;
  if (positions == nullptr) {
    PB_RETURN_IF_ERROR(EvaluateDegreesOfFreedom(e, w.get(), t, stream));
  } else {
    std::vector<double> internal;
    PB_RETURN_IF_ERROR(w->stage_table.Reserve(view.stage_stride));
    PB_RETURN_IF_ERROR(w->scratch_c.Reserve(3 * n));
    ToInternal(e, positions, 3, internal);
    PB_CUDA(cudaMemcpyAsync(w->stage_table.data, internal.data(),
                            internal.size() * sizeof(double),
                            cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    if (jerks != nullptr) {
      ToInternal(e, velocities, 3, internal);
      PB_CUDA(cudaMemcpyAsync(w->scratch_c.data, internal.data(),
                              internal.size() * sizeof(double),
                              cudaMemcpyHostToDevice, stream));
      PB_CUDA(cudaStreamSynchronize(stream));
    }
  }
  // scratch_b: 3n jerks then 9n Jacobian entries.
  PB_RETURN_IF_ERROR(w->scratch_b.Reserve(12 * n));
  MassiveJerkJacobianKernel<<<(n + 31) / 32, 32, 0, stream>>>(
      e->View(), w->stage_table.data, w->scratch_c.data,
      jerks != nullptr ? w->scratch_b.data : nullptr,
      jacobians != nullptr ? w->scratch_b.data + 3 * n : nullptr);
  PB_LAUNCHED();
  std::vector<double> internal(12 * n);
  PB_CUDA(cudaMemcpyAsync(internal.data(), w->scratch_b.data,
                          internal.size() * sizeof(double),
                          cudaMemcpyDeviceToHost, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  if (jerks != nullptr) {
    ToCaller(e, std::vector<double>(internal.begin(), internal.begin() + 3 * n),
             3, jerks);
  }
  if (jacobians != nullptr) {
    ToCaller(e, std::vector<double>(internal.begin() + 3 * n, internal.end()), 9,
             jacobians);
  }
  return PB_OK;
}

}  // extern "C"
