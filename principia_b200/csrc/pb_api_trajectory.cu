// C ABI: the trajectory sink of the flows, DiscreteTrajectorySegment::Append
// with downsampling for n trajectories on the device
// (include/principia_b200.h).
#include <cuda_runtime.h>

#include <vector>

#include "../../include/principia_b200.h"
#include "pb_downsampling.cuh"
#include "pb_runtime.cuh"

using namespace pb;

struct pb_downsampler {
  int device = 0;
  long long n = 0;
  long long capacity = 0;
  double tolerance = 0;
  DeviceBuffer<double> times, points, errors, downsampling_error, last_a2a3;
  DeviceBuffer<long long> counts;
  DeviceBuffer<int32_t> overflow;
  DeviceBuffer<double> staged_times, staged_samples;
  DownsamplerView View() {
    return {n,           capacity,
            tolerance,   times.data,
            points.data, errors.data,
            counts.data, downsampling_error.data,
            last_a2a3.data, overflow.data};
  }
};

pb::DownsamplerView PbDownsamplerView(pb_downsampler* const downsampler) {
  return downsampler->View();
}

namespace {

// One thread per trajectory; the samples of a trajectory are appended in
// order.  `times` is [n_samples] (one time for all trajectories: the fixed-step
// flows) or, with per_trajectory, [n_samples][n] with valid[i] samples for
// trajectory i (the adaptive flow's step and state logs).
__global__ void __launch_bounds__(128)
DownsampleKernel(DownsamplerView const d, long long const n_samples,
                 double const* __restrict__ times, bool const per_trajectory,
                 long long const* __restrict__ valid,
                 double const* __restrict__ samples) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= d.n) {
    return;
  }
  long long const n = d.n;
  long long const count =
      valid != nullptr ? (valid[i] < n_samples ? valid[i] : n_samples)
                       : n_samples;
  for (long long s = 0; s < count; ++s) {
    double const t = per_trajectory ? times[s * n + i] : times[s];
    double const* const sample = samples + s * 6 * n + i;
    DownsamplerAppend(d, i, t,
                      Vec3{sample[0], sample[n], sample[2 * n]},
                      Vec3{sample[3 * n], sample[4 * n], sample[5 * n]});
  }
}

pb_status_t LaunchDownsample(pb_downsampler* d, long long const n_samples,
                             double const* times, bool const per_trajectory,
                             long long const* valid, double const* samples,
                             cudaStream_t stream) {
  if (d->n == 0 || n_samples == 0) {
    return PB_OK;
  }
  DownsampleKernel<<<static_cast<unsigned>((d->n + 127) / 128), 128, 0,
                     stream>>>(d->View(), n_samples, times, per_trajectory,
                               valid, samples);
  PB_LAUNCHED();
  return PB_OK;
}

}  // namespace

namespace {

// DiscreteTrajectorySegment::EvaluatePosition / EvaluateVelocity,
// physics/discrete_trajectory_segment_body.hpp:132-155, for every timeline at
// one time: lower_bound(t), the point itself if its time is t, else the
// Hermite interpolation between the previous point and it (get_hermite3,
// :493-509).  A timeline that does not cover t (the reference CHECKs
// t_min() <= t <= t_max()) gets NaNs and in_range[i] = 0.
__global__ void __launch_bounds__(128)
DownsamplerEvaluateKernel(DownsamplerView const d, double const t,
                          double* __restrict__ q, double* __restrict__ v,
                          int32_t* __restrict__ in_range) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= d.n) {
    return;
  }
  long long const n = d.n;
  long long const count = d.counts[i] < d.capacity ? d.counts[i] : d.capacity;
  double const nan = __longlong_as_double(0x7ff8000000000000LL);
  Vec3 position{nan, nan, nan};
  Vec3 velocity{nan, nan, nan};
  bool ok = count > 0 && t >= d.times[i] && t <= d.times[(count - 1) * n + i];
  if (ok) {
    long long lo = 0;
    long long hi = count;
    while (lo < hi) {
      long long const mid = (lo + hi) / 2;
      if (d.times[mid * n + i] < t) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    auto const load = [&](long long const slot, int const plane) {
      return d.points[(slot * 6 + plane) * n + i];
    };
    double const t2 = d.times[lo * n + i];
    Vec3 const q2{load(lo, 0), load(lo, 1), load(lo, 2)};
    Vec3 const v2{load(lo, 3), load(lo, 4), load(lo, 5)};
    if (t2 == t) {
      position = q2;
      velocity = v2;
    } else {
      double const t1 = d.times[(lo - 1) * n + i];
      Vec3 const q1{load(lo - 1, 0), load(lo - 1, 1), load(lo - 1, 2)};
      Vec3 const v1{load(lo - 1, 3), load(lo - 1, 4), load(lo - 1, 5)};
      Hermite3 const h = MakeHermite3(t1, t2, q1, q2, v1, v2);
      position = Hermite3Evaluate(h, t);
      velocity = Hermite3EvaluateDerivative(h, t);
    }
  }
  if (q != nullptr) {
    q[i] = position.x;
    q[n + i] = position.y;
    q[2 * n + i] = position.z;
  }
  if (v != nullptr) {
    v[i] = velocity.x;
    v[n + i] = velocity.y;
    v[2 * n + i] = velocity.z;
  }
  if (in_range != nullptr) {
    in_range[i] = ok ? 1 : 0;
  }
}

}  // namespace

extern "C" {

pb_status_t pb_downsampler_create(int32_t cuda_device, int64_t n,
                                  double tolerance, int64_t capacity,
                                  pb_downsampler_t** downsampler) {
  if (downsampler == nullptr || n < 0 || capacity < 2 ||
      tolerance != tolerance) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  *downsampler = nullptr;
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0) {
    return Fail(PB_FAILED_PRECONDITION,
                "no CUDA device; this library has no CPU fallback");
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  auto* d = new pb_downsampler;
  d->device = cuda_device;
  d->n = n;
  d->capacity = capacity;
  d->tolerance = tolerance;
  std::size_t const slots =
      static_cast<std::size_t>(capacity) * static_cast<std::size_t>(n);
  std::size_t const count = static_cast<std::size_t>(n);
  pb_status_t status = d->times.Reserve(slots + 1);
  if (status == PB_OK) status = d->points.Reserve(6 * slots + 1);
  if (status == PB_OK) status = d->errors.Reserve(slots + 1);
  if (status == PB_OK) status = d->downsampling_error.Reserve(count + 1);
  if (status == PB_OK) status = d->last_a2a3.Reserve(6 * count + 1);
  if (status == PB_OK) status = d->counts.Reserve(count + 1);
  if (status == PB_OK) status = d->overflow.Reserve(count + 1);
  if (status == PB_OK) {
    cudaError_t error = cudaMemset(d->counts.data, 0,
                                   (count + 1) * sizeof(long long));
    if (error == cudaSuccess) {
      error = cudaMemset(d->overflow.data, 0, (count + 1) * sizeof(int32_t));
    }
    if (error == cudaSuccess) {
      error = cudaMemset(d->downsampling_error.data, 0,
                         (count + 1) * sizeof(double));
    }
    if (error != cudaSuccess) {
      status = Fail(PB_INTERNAL, cudaGetErrorString(error));
    }
  }
  if (status != PB_OK) {
    delete d;
    return status;
  }
  *downsampler = d;
  return PB_OK;
}

void pb_downsampler_destroy(pb_downsampler_t* d) {
  if (d != nullptr) {
    cudaSetDevice(d->device);
    delete d;
  }
}

pb_status_t pb_downsampler_append_device(pb_downsampler_t* d,
                                         int64_t n_samples,
                                         const double* sample_times,
                                         const double* samples,
                                         void* stream_pointer) {
  if (d == nullptr || n_samples < 0 ||
      (n_samples > 0 && (sample_times == nullptr || samples == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(d->device));
  cudaStream_t const stream = static_cast<cudaStream_t>(stream_pointer);
  PB_RETURN_IF_ERROR(d->staged_times.Upload(
      sample_times, static_cast<std::size_t>(n_samples), stream));
  PB_RETURN_IF_ERROR(LaunchDownsample(d, n_samples, d->staged_times.data,
                                      /*per_trajectory=*/false, nullptr,
                                      samples, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  return PB_OK;
}

pb_status_t pb_downsampler_append(pb_downsampler_t* d, int64_t n_samples,
                                  const double* sample_times,
                                  const double* samples) {
  if (d == nullptr || n_samples < 0 ||
      (n_samples > 0 && (sample_times == nullptr || samples == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(d->device));
  PB_RETURN_IF_ERROR(d->staged_samples.Upload(
      samples,
      static_cast<std::size_t>(n_samples) * 6 * static_cast<std::size_t>(d->n),
      nullptr));
  return pb_downsampler_append_device(d, n_samples, sample_times,
                                      d->staged_samples.data, nullptr);
}

pb_status_t pb_downsampler_append_logs_device(pb_downsampler_t* d,
                                              int64_t log_capacity,
                                              const double* step_log,
                                              const int64_t* accepted_steps,
                                              const double* state_log,
                                              void* stream_pointer) {
  if (d == nullptr || log_capacity < 0 ||
      (log_capacity > 0 && d->n > 0 &&
       (step_log == nullptr || accepted_steps == nullptr ||
        state_log == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(d->device));
  cudaStream_t const stream = static_cast<cudaStream_t>(stream_pointer);
  PB_RETURN_IF_ERROR(LaunchDownsample(
      d, log_capacity, step_log, /*per_trajectory=*/true,
      reinterpret_cast<long long const*>(accepted_steps), state_log, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  return PB_OK;
}

pb_status_t pb_downsampler_get(pb_downsampler_t* d, int64_t* counts,
                               double* times, double* points, double* errors) {
  if (d == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null downsampler");
  }
  PB_CUDA(cudaSetDevice(d->device));
  std::size_t const n = static_cast<std::size_t>(d->n);
  std::size_t const slots = static_cast<std::size_t>(d->capacity) * n;
  if (n == 0) {
    return PB_OK;
  }
  if (counts != nullptr) {
    PB_CUDA(cudaMemcpy(counts, d->counts.data, n * sizeof(int64_t),
                       cudaMemcpyDeviceToHost));
  }
  if (times != nullptr) {
    PB_CUDA(cudaMemcpy(times, d->times.data, slots * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  if (points != nullptr) {
    PB_CUDA(cudaMemcpy(points, d->points.data, 6 * slots * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  if (errors != nullptr) {
    PB_CUDA(cudaMemcpy(errors, d->errors.data, slots * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  std::vector<int32_t> overflow(n);
  PB_CUDA(cudaMemcpy(overflow.data(), d->overflow.data, n * sizeof(int32_t),
                     cudaMemcpyDeviceToHost));
  for (int32_t const o : overflow) {
    if (o != 0) {
      return Fail(PB_RESOURCE_EXHAUSTED,
                  "a timeline outgrew the capacity of the downsampler");
    }
  }
  return PB_OK;
}

pb_status_t pb_downsampler_evaluate_device(pb_downsampler_t* d, double t,
                                           double* q, double* v,
                                           int32_t* in_range, void* stream) {
  if (d == nullptr || (q == nullptr && v == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(d->device));
  if (d->n > 0) {
    DownsamplerEvaluateKernel<<<static_cast<unsigned>((d->n + 127) / 128), 128,
                                0, static_cast<cudaStream_t>(stream)>>>(
        d->View(), t, q, v, in_range);
    PB_LAUNCHED();
  }
  return PB_OK;
}

pb_status_t pb_downsampler_evaluate(pb_downsampler_t* d, double t, double* q,
                                    double* v, int32_t* in_range) {
  if (d == nullptr || (q == nullptr && v == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(d->device));
  std::size_t const n = static_cast<std::size_t>(d->n);
  if (n == 0) {
    return PB_OK;
  }
  // staged_samples: q then v; the flags after them.
  PB_RETURN_IF_ERROR(d->staged_samples.Reserve(7 * n));
  double* const device_q = d->staged_samples.data;
  double* const device_v = device_q + 3 * n;
  int32_t* const device_flags = reinterpret_cast<int32_t*>(device_v + 3 * n);
  PB_RETURN_IF_ERROR(pb_downsampler_evaluate_device(
      d, t, device_q, device_v, device_flags, nullptr));
  if (q != nullptr) {
    PB_CUDA(cudaMemcpy(q, device_q, 3 * n * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  if (v != nullptr) {
    PB_CUDA(cudaMemcpy(v, device_v, 3 * n * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  if (in_range != nullptr) {
    PB_CUDA(cudaMemcpy(in_range, device_flags, n * sizeof(int32_t),
                       cudaMemcpyDeviceToHost));
  }
  return PB_OK;
}

}  // extern "C"
