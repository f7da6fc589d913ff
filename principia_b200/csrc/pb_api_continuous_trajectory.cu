// C ABI: batches of ContinuousTrajectory produced and evaluated on the device
// (include/principia_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <limits>
#include <vector>

#include "../../include/principia_b200.h"
#include "generated_tables.h"
#include "pb_kernels_newhall.cuh"
#include "pb_runtime.cuh"

using namespace pb;

struct pb_continuous_trajectories {
  int device = 0;
  long long n = 0;
  double step = 0;
  double tolerance = 0;
  long long max_segments = 0;

  // last_points_: the ring of the last (at most 9) points and their times.
  DeviceBuffer<double> ring;  // [9][6][n].
  double ring_times[kNewhallRing];
  int first_slot = 0;
  int n_points = 0;
  bool has_first_time = false;
  double first_time = 0;

  DeviceBuffer<double> adjusted_tolerance;
  DeviceBuffer<int32_t> degree, degree_age, is_unstable;
  DeviceBuffer<int32_t> status_word;

  DeviceBuffer<double> monomial, chebyshev_last_row;

  // polynomials_: the times are shared by the trajectories.
  std::vector<double> t_max, origin;
  DeviceBuffer<int32_t> degrees;      // [segment][n].
  DeviceBuffer<double> coefficients;  // [segment][xyz][18][n].

  DeviceBuffer<double> scratch;  // Host-facing staging, 6 n.

  NewhallTables Tables() const {
    NewhallTables tables;
    tables.monomial = monomial.data;
    tables.chebyshev_last_row = chebyshev_last_row.data;
    for (int i = 0; i < 15; ++i) {
      tables.monomial_offset[i] = pb_newhall_monomial_offset[i];
    }
    return tables;
  }
};

namespace pb {

namespace {

__global__ void FillDoubleKernel(double* const out, long long const n,
                                 double const value) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    out[i] = value;
  }
}

__global__ void FillIntKernel(int32_t* const out, long long const n,
                              int32_t const value) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    out[i] = value;
  }
}

}  // namespace

// ContinuousTrajectory::Append for every trajectory: q and v are 3 planes of n
// in device memory.  Asynchronous on `stream`.
pb_status_t ContinuousTrajectoriesAppendDevice(pb_continuous_trajectories* h,
                                               double const time,
                                               double const* q, double const* v,
                                               cudaStream_t stream) {
  long long const n = h->n;
  if (h->n_points > 0 &&
      !(time > h->ring_times[(h->first_slot + h->n_points - 1) % kNewhallRing])) {
    return Fail(PB_INVALID_ARGUMENT, "Append out of order");
  }
  if (!h->has_first_time) {
    h->first_time = time;
    h->has_first_time = true;
  }
  bool const fit = h->n_points == kNewhallDivisions;
  if (fit && static_cast<long long>(h->t_max.size()) >= h->max_segments) {
    return Fail(PB_RESOURCE_EXHAUSTED, "max_segments reached");
  }
  int const slot = (h->first_slot + h->n_points) % kNewhallRing;
  double* const point = h->ring.data + static_cast<long long>(slot) * 6 * n;
  PB_CUDA(cudaMemcpyAsync(point, q, 3 * n * sizeof(double),
                          cudaMemcpyDeviceToDevice, stream));
  PB_CUDA(cudaMemcpyAsync(point + 3 * n, v, 3 * n * sizeof(double),
                          cudaMemcpyDeviceToDevice, stream));
  h->ring_times[slot] = time;
  ++h->n_points;
  if (fit) {
    double const t_min = h->ring_times[h->first_slot];
    long long const segment = static_cast<long long>(h->t_max.size());
    NewhallState const state{h->adjusted_tolerance.data, h->degree.data,
                             h->degree_age.data, h->is_unstable.data};
    NewhallPoints const points{h->ring.data, h->first_slot};
    NewhallFitKernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(
        h->Tables(), state, points, n, h->tolerance, t_min, time, segment,
        h->degrees.data, h->coefficients.data, h->status_word.data);
    PB_LAUNCHED();
    h->t_max.push_back(time);
    // Barycentre({t_min, t_max}), geometry/barycentre_calculator_body.hpp:77-96.
    double total = 0.0;
    total += t_min - 0.0;
    total += time - 0.0;
    h->origin.push_back(0.0 + total / 2);
    // last_points_.clear(); last_points_.emplace_back(time, ...).
    h->first_slot = slot;
    h->n_points = 1;
  }
  return PB_OK;
}

long long ContinuousTrajectoriesSize(pb_continuous_trajectories const* h) {
  return h->n;
}

bool ContinuousTrajectoriesEmpty(pb_continuous_trajectories const* h) {
  return !h->has_first_time;
}

}  // namespace pb

extern "C" {

pb_status_t pb_continuous_trajectories_create(
    int32_t device, int64_t n_trajectories, double step, double tolerance,
    int64_t max_segments, pb_continuous_trajectories_t** out) {
  if (out == nullptr || n_trajectories <= 0 || max_segments <= 0 ||
      !(step > 0)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(device));
  auto* const h = new pb_continuous_trajectories;
  h->device = device;
  h->n = n_trajectories;
  h->step = step;
  h->tolerance = tolerance;
  h->max_segments = max_segments;
  std::size_t const n = static_cast<std::size_t>(n_trajectories);
  std::size_t const segments = static_cast<std::size_t>(max_segments);
  pb_status_t status = PB_OK;
  auto reserve = [&status](auto& buffer, std::size_t const size) {
    if (status == PB_OK) {
      status = buffer.Reserve(size);
    }
  };
  reserve(h->ring, kNewhallRing * 6 * n);
  reserve(h->adjusted_tolerance, n);
  reserve(h->degree, n);
  reserve(h->degree_age, n);
  reserve(h->is_unstable, n);
  reserve(h->status_word, 1);
  reserve(h->degrees, segments * n);
  reserve(h->coefficients, segments * 3 * kSegmentCoefficients * n);
  reserve(h->scratch, 6 * n);
  if (status == PB_OK) {
    status = h->monomial.Upload(pb_newhall_monomial, 2970, nullptr);
  }
  if (status == PB_OK) {
    status = h->chebyshev_last_row.Upload(pb_newhall_chebyshev_last_row, 270,
                                          nullptr);
  }
  if (status != PB_OK) {
    delete h;
    return status;
  }
  unsigned const grid = static_cast<unsigned>((n + 255) / 256);
  FillDoubleKernel<<<grid, 256>>>(h->adjusted_tolerance.data, h->n, tolerance);
  FillIntKernel<<<grid, 256>>>(h->degree.data, h->n, kNewhallMinDegree);
  FillIntKernel<<<grid, 256>>>(h->degree_age.data, h->n, 0);
  FillIntKernel<<<grid, 256>>>(h->is_unstable.data, h->n, 0);
  FillIntKernel<<<1, 32>>>(h->status_word.data, 1, 0);
  PB_CUDA(cudaDeviceSynchronize());
  *out = h;
  return PB_OK;
}

void pb_continuous_trajectories_destroy(pb_continuous_trajectories_t* h) {
  if (h != nullptr) {
    cudaSetDevice(h->device);
    delete h;
  }
}

pb_status_t pb_continuous_trajectories_append_device(
    pb_continuous_trajectories_t* h, double time, const double* q,
    const double* v, void* stream) {
  if (h == nullptr || q == nullptr || v == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  PB_CUDA(cudaSetDevice(h->device));
  return ContinuousTrajectoriesAppendDevice(h, time, q, v,
                                            static_cast<cudaStream_t>(stream));
}

pb_status_t pb_continuous_trajectories_status(pb_continuous_trajectories_t* h) {
  if (h == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  PB_CUDA(cudaSetDevice(h->device));
  // The appends may be in flight on any stream.
  PB_CUDA(cudaDeviceSynchronize());
  int32_t word = 0;
  PB_CUDA(cudaMemcpy(&word, h->status_word.data, sizeof(word),
                     cudaMemcpyDeviceToHost));
  if (word != 0) {
    return Fail(PB_INVALID_ARGUMENT, "Error extending trajectory: apocalypse");
  }
  return PB_OK;
}

pb_status_t pb_continuous_trajectories_append(pb_continuous_trajectories_t* h,
                                              double time, const double* q,
                                              const double* v) {
  if (h == nullptr || q == nullptr || v == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  PB_CUDA(cudaSetDevice(h->device));
  std::size_t const n = static_cast<std::size_t>(h->n);
  PB_CUDA(cudaMemcpy(h->scratch.data, q, 3 * n * sizeof(double),
                     cudaMemcpyHostToDevice));
  PB_CUDA(cudaMemcpy(h->scratch.data + 3 * n, v, 3 * n * sizeof(double),
                     cudaMemcpyHostToDevice));
  PB_RETURN_IF_ERROR(ContinuousTrajectoriesAppendDevice(
      h, time, h->scratch.data, h->scratch.data + 3 * n, nullptr));
  return pb_continuous_trajectories_status(h);
}

double pb_continuous_trajectories_t_min(const pb_continuous_trajectories_t* h) {
  // ContinuousTrajectory::t_min_locked, continuous_trajectory_body.hpp:201-208.
  return h->t_max.empty() ? std::numeric_limits<double>::infinity()
                          : h->first_time;
}

double pb_continuous_trajectories_t_max(const pb_continuous_trajectories_t* h) {
  return h->t_max.empty() ? -std::numeric_limits<double>::infinity()
                          : h->t_max.back();
}

int64_t pb_continuous_trajectories_segment_count(
    const pb_continuous_trajectories_t* h) {
  return static_cast<int64_t>(h->t_max.size());
}

pb_status_t pb_continuous_trajectories_get_segments(
    pb_continuous_trajectories_t* h, int64_t trajectory, double* t_min,
    double* t_max, double* origin, int32_t* degree, double* coefficients) {
  if (h == nullptr || trajectory < 0 || trajectory >= h->n) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(h->device));
  std::size_t const segments = h->t_max.size();
  std::size_t const n = static_cast<std::size_t>(h->n);
  if (t_min != nullptr) {
    *t_min = h->first_time;
  }
  if (segments == 0) {
    return PB_OK;
  }
  PB_CUDA(cudaDeviceSynchronize());
  if (t_max != nullptr) {
    std::copy(h->t_max.begin(), h->t_max.end(), t_max);
  }
  if (origin != nullptr) {
    std::copy(h->origin.begin(), h->origin.end(), origin);
  }
  if (degree != nullptr) {
    PB_CUDA(cudaMemcpy2D(degree, sizeof(int32_t), h->degrees.data + trajectory,
                         n * sizeof(int32_t), sizeof(int32_t), segments,
                         cudaMemcpyDeviceToHost));
  }
  if (coefficients != nullptr) {
    // Device [segment][xyz][18][n] -> host [segment][18][xyz].
    std::vector<double> staged(segments * 3 * kSegmentCoefficients);
    PB_CUDA(cudaMemcpy2D(staged.data(), sizeof(double),
                         h->coefficients.data + trajectory, n * sizeof(double),
                         sizeof(double), staged.size(),
                         cudaMemcpyDeviceToHost));
    for (std::size_t s = 0; s < segments; ++s) {
      for (int c = 0; c < 3; ++c) {
        for (int k = 0; k < kSegmentCoefficients; ++k) {
          coefficients[(s * kSegmentCoefficients + k) * 3 + c] =
              staged[(s * 3 + c) * kSegmentCoefficients + k];
        }
      }
    }
  }
  return PB_OK;
}

pb_status_t pb_continuous_trajectories_evaluate_device(
    pb_continuous_trajectories_t* h, double t, double* positions,
    double* velocities, void* stream_pointer) {
  if (h == nullptr || (positions == nullptr && velocities == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  // CHECK_LE(t_min_locked(), time); CHECK_GE(t_max_locked(), time).
  if (h->t_max.empty() || !(t >= h->first_time) || !(t <= h->t_max.back())) {
    return Fail(PB_INVALID_ARGUMENT,
                "time outside [t_min, t_max] of the trajectories");
  }
  PB_CUDA(cudaSetDevice(h->device));
  // FindPolynomialForInstantLocked, :860-885: the first polynomial with
  // t <= t_max.
  long long const segment =
      std::lower_bound(h->t_max.begin(), h->t_max.end(), t) - h->t_max.begin();
  double const x = t - h->origin[segment];
  ContinuousTrajectoriesEvaluateKernel<<<
      static_cast<unsigned>((h->n + 127) / 128), 128, 0,
      static_cast<cudaStream_t>(stream_pointer)>>>(
      h->n, segment, x, h->degrees.data, h->coefficients.data, positions,
      velocities);
  PB_LAUNCHED();
  return PB_OK;
}

pb_status_t pb_continuous_trajectories_evaluate(pb_continuous_trajectories_t* h,
                                                double t, double* positions,
                                                double* velocities) {
  if (h == nullptr || (positions == nullptr && velocities == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(h->device));
  std::size_t const n = static_cast<std::size_t>(h->n);
  PB_RETURN_IF_ERROR(pb_continuous_trajectories_evaluate_device(
      h, t, positions != nullptr ? h->scratch.data : nullptr,
      velocities != nullptr ? h->scratch.data + 3 * n : nullptr, nullptr));
  if (positions != nullptr) {
    PB_CUDA(cudaMemcpy(positions, h->scratch.data, 3 * n * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  if (velocities != nullptr) {
    PB_CUDA(cudaMemcpy(velocities, h->scratch.data + 3 * n,
                       3 * n * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return PB_OK;
}

}  // extern "C"
