// Runtime plumbing of the CUDA library: error reporting, device buffers, the
// ephemeris handle.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/principia_b200.h"
#include "pb_host_model.hpp"
#include "pb_model.h"

namespace pb {

inline std::string& LastError() {
  static thread_local std::string error;
  return error;
}

inline std::atomic<std::int64_t>& LaunchCount() {
  static std::atomic<std::int64_t> count{0};
  return count;
}

inline std::uint64_t NextSerial() {
  static std::atomic<std::uint64_t> serial{0};
  return ++serial;
}

// The flow entry points share the constant bank of their kernels (per-body
// scalars, hot-body tables) and the scratch buffers of the ephemeris: they are
// serialised.  (The reference lets flows on distinct objects run concurrently,
// ephemeris.hpp:64-66; here concurrency is inside a call, across the probes.)
inline std::recursive_mutex& FlowMutex() {
  static std::recursive_mutex mutex;
  return mutex;
}

inline pb_status_t Fail(pb_status_t const status, std::string const& message) {
  LastError() = message;
  return status;
}

#define PB_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t const pb_cuda_error = (call);                                \
    if (pb_cuda_error != cudaSuccess) {                                      \
      return ::pb::Fail(                                                     \
          pb_cuda_error == cudaErrorNoDevice ||                              \
                  pb_cuda_error == cudaErrorInsufficientDriver               \
              ? PB_FAILED_PRECONDITION                                       \
              : PB_INTERNAL,                                                 \
          std::string(#call) + ": " + cudaGetErrorString(pb_cuda_error));    \
    }                                                                        \
  } while (0)

#define PB_RETURN_IF_ERROR(expr)               \
  do {                                         \
    pb_status_t const pb_status = (expr);      \
    if (pb_status != PB_OK) return pb_status;  \
  } while (0)

#define PB_LAUNCHED()                                  \
  do {                                                 \
    ::pb::LaunchCount().fetch_add(1);                  \
    PB_CUDA(cudaGetLastError());                       \
  } while (0)

// A grow-only device allocation.
template<typename T>
struct DeviceBuffer {
  T* data = nullptr;
  std::size_t capacity = 0;

  ~DeviceBuffer() {
    if (data != nullptr) {
      cudaFree(data);
    }
  }
  DeviceBuffer() = default;
  DeviceBuffer(DeviceBuffer const&) = delete;
  DeviceBuffer& operator=(DeviceBuffer const&) = delete;

  pb_status_t Reserve(std::size_t const n) {
    if (n <= capacity) {
      return PB_OK;
    }
    if (data != nullptr) {
      PB_CUDA(cudaFree(data));
      data = nullptr;
      capacity = 0;
    }
    std::size_t const wanted = n + n / 4;
    PB_CUDA(cudaMalloc(&data, wanted * sizeof(T)));
    capacity = wanted;
    return PB_OK;
  }

  pb_status_t Upload(T const* const host, std::size_t const n,
                     cudaStream_t const stream) {
    PB_RETURN_IF_ERROR(Reserve(n));
    if (n > 0) {
      PB_CUDA(cudaMemcpyAsync(data, host, n * sizeof(T), cudaMemcpyHostToDevice,
                              stream));
    }
    return PB_OK;
  }
  pb_status_t Upload(std::vector<T> const& host, cudaStream_t const stream) {
    return Upload(host.data(), host.size(), stream);
  }
};

// Page-locked host staging, grow-only.
template<typename T>
struct PinnedBuffer {
  T* data = nullptr;
  std::size_t capacity = 0;
  ~PinnedBuffer() {
    if (data != nullptr) {
      cudaFreeHost(data);
    }
  }
  PinnedBuffer() = default;
  PinnedBuffer(PinnedBuffer const&) = delete;
  PinnedBuffer& operator=(PinnedBuffer const&) = delete;
  pb_status_t Reserve(std::size_t const n) {
    if (n <= capacity) {
      return PB_OK;
    }
    if (data != nullptr) {
      PB_CUDA(cudaFreeHost(data));
      data = nullptr;
      capacity = 0;
    }
    std::size_t const wanted = n + n / 4;
    PB_CUDA(cudaMallocHost(&data, wanted * sizeof(T)));
    capacity = wanted;
    return PB_OK;
  }
};

struct HostTrajectory {
  bool has_t_min = false;
  double t_min = 0;
  std::vector<double> t_max;
  std::vector<double> origin;
  std::vector<int32_t> degree;
  std::vector<double> coefficients;  // [segment][18][3].
};

}  // namespace pb

// The opaque handle of the C ABI.
struct pb_ephemeris {
  int device = 0;
  // Distinguishes handles whose storage the allocator reuses (caches of
  // constant-bank uploads are keyed by it).
  std::uint64_t const serial = pb::NextSerial();
  pb::HostModel model;
  std::vector<pb::HostTrajectory> trajectories;  // Internal order.
  bool segments_dirty = true;
  bool uniform_segments = false;  // Set by UploadSegments.

  pb::DeviceBuffer<pb::DeviceBody> bodies;
  pb::DeviceBuffer<pb::Damping> dampings;
  pb::DeviceBuffer<double> cos, sin, normalization;
  pb::DeviceBuffer<int32_t> segment_offset;
  pb::DeviceBuffer<double> segment_t_max, segment_origin, segment_coefficients;
  pb::DeviceBuffer<int32_t> segment_degree;

  // Scratch.
  pb::DeviceBuffer<double> stage_times;
  pb::DeviceBuffer<double> stage_table;
  pb::DeviceBuffer<double> scratch_a, scratch_b, scratch_c;
  pb::DeviceBuffer<double> burns_scratch;
  pb::DeviceBuffer<int32_t> status_word;
  pb::DeviceBuffer<long long> counters;
  pb::PinnedBuffer<double> pinned_times;
  pb::PinnedBuffer<int32_t> pinned_status;

  // Adaptive flow: scheduling keys, execution order, sort storage, queue head.
  pb::DeviceBuffer<float> adaptive_cost;
  pb::DeviceBuffer<int> adaptive_order;
  pb::DeviceBuffer<unsigned char> adaptive_sort_storage;
  pb::DeviceBuffer<unsigned long long> adaptive_queue;

  // CUDA-event timing of the dominant kernel of the last flow call.
  std::vector<cudaEvent_t> timing_events;
  double last_flow_kernel_ms = 0;
  std::int64_t last_flow_kernel_launches = 0;

  // Helper streams of the sliced flow launches and their join events.
  std::vector<cudaStream_t> flow_streams;
  std::vector<cudaEvent_t> flow_events;

  ~pb_ephemeris() {
    for (cudaEvent_t event : timing_events) {
      cudaEventDestroy(event);
    }
    for (cudaEvent_t event : flow_events) {
      cudaEventDestroy(event);
    }
    for (cudaStream_t helper : flow_streams) {
      cudaStreamDestroy(helper);
    }
  }

  pb::EphemerisView View() const {
    pb::EphemerisView v;
    v.n_bodies = static_cast<int32_t>(model.bodies.size());
    v.n_oblate = model.n_oblate;
    v.n_frames = model.n_frames;
    v.stage_stride = pb::StageStride(v.n_bodies, v.n_frames);
    v.uniform_segments = uniform_segments ? 1 : 0;
    v.bodies = bodies.data;
    v.dampings = dampings.data;
    v.cos = cos.data;
    v.sin = sin.data;
    v.normalization = normalization.data;
    v.segment_offset = segment_offset.data;
    v.segment_t_max = segment_t_max.data;
    v.segment_origin = segment_origin.data;
    v.segment_degree = segment_degree.data;
    v.segment_coefficients = segment_coefficients.data;
    return v;
  }
};

namespace pb {
// Defined in principia_b200.cu.
pb_status_t UploadSegments(pb_ephemeris* e, cudaStream_t stream);
double EphemerisTMin(pb_ephemeris const* e);
double EphemerisTMax(pb_ephemeris const* e);
pb_status_t BuildStageTable(pb_ephemeris* e, double const* times_host,
                            int n_times, double* table, cudaStream_t stream);
// Defined in pb_api_continuous_trajectory.cu.
pb_status_t ContinuousTrajectoriesAppendDevice(pb_continuous_trajectories* h,
                                               double time, double const* q,
                                               double const* v,
                                               cudaStream_t stream);
long long ContinuousTrajectoriesSize(pb_continuous_trajectories const* h);
bool ContinuousTrajectoriesEmpty(pb_continuous_trajectories const* h);
}  // namespace pb
