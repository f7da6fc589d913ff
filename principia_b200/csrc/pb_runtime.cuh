// Runtime plumbing of the CUDA library: error reporting, device buffers, the
// ephemeris handle.
#pragma once

#include <cuda_runtime.h>

#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <mutex>
#include <shared_mutex>
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

// An NVTX range per C-ABI call (nsys / ncu timelines name the entry points).
struct ApiRange {
  explicit ApiRange(char const* const name) { nvtxRangePushA(name); }
  ~ApiRange() { nvtxRangePop(); }
  ApiRange(ApiRange const&) = delete;
  ApiRange& operator=(ApiRange const&) = delete;
};
#define PB_API_RANGE() ::pb::ApiRange const pb_api_range(__func__)

// A bank of __constant__ tables that several kernels of one translation unit
// read as instruction operands (per-body scalars, the hot body's harmonics).
// Calls that need the same contents (the same ephemeris, the same hot body)
// share it and run concurrently; a call that needs other contents waits until
// no kernel that reads the current ones is in flight (every reader
// synchronises its stream before it releases the bank), then rewrites it.
class ConstantBank {
 public:
  // `upload` copies the tables to the symbols and synchronises; it runs under
  // the bank's lock, only when the contents change.
  template<typename Upload>
  pb_status_t Acquire(std::uint64_t const ephemeris_serial, int const variant,
                      Upload&& upload) {
    std::unique_lock<std::mutex> lock(mutex_);
    changed_.wait(lock, [&] {
      return readers_ == 0 ||
             (valid_ && serial_ == ephemeris_serial && variant_ == variant);
    });
    if (!(valid_ && serial_ == ephemeris_serial && variant_ == variant)) {
      valid_ = false;
      pb_status_t const status = upload();
      if (status != PB_OK) {
        return status;
      }
      valid_ = true;
      serial_ = ephemeris_serial;
      variant_ = variant;
    }
    ++readers_;
    return PB_OK;
  }
  void Release() {
    std::lock_guard<std::mutex> const lock(mutex_);
    if (--readers_ == 0) {
      changed_.notify_all();
    }
  }

 private:
  std::mutex mutex_;
  std::condition_variable changed_;
  bool valid_ = false;
  std::uint64_t serial_ = 0;
  int variant_ = 0;
  int readers_ = 0;
};

// Releases a ConstantBank on scope exit.
class ConstantBankLease {
 public:
  ConstantBankLease() = default;
  ~ConstantBankLease() {
    if (bank_ != nullptr) {
      bank_->Release();
    }
  }
  ConstantBankLease(ConstantBankLease const&) = delete;
  ConstantBankLease& operator=(ConstantBankLease const&) = delete;
  void Adopt(ConstantBank* const bank) { bank_ = bank; }
  bool held() const { return bank_ != nullptr; }

 private:
  ConstantBank* bank_ = nullptr;
};

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

// Device allocations come from the device's default memory pool
// (cudaMallocAsync) with a release threshold of 8 GB, so that objects that are
// created and destroyed repeatedly (Ephemeris::NewInstance per flow, the
// workspaces of concurrent calls) reuse their memory instead of going through
// the OS: cudaFree of a 10 MB state was measured at up to 0.7 s, cudaMalloc at
// up to 0.1 s (profiles/r02p_srkn_flow_round2.md).  The semantics of cudaMalloc
// / cudaFree are kept: the allocation is valid on every stream when
// PooledMalloc returns, and PooledFree waits for the device like cudaFree does.
inline bool MemoryPoolsUsable() {
  static bool const usable = [] {
    int device = 0;
    int supported = 0;
    if (cudaGetDevice(&device) != cudaSuccess ||
        cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported,
                               device) != cudaSuccess ||
        supported == 0) {
      return false;
    }
    int count = 0;
    cudaGetDeviceCount(&count);
    for (int d = 0; d < count; ++d) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) {
        std::uint64_t threshold = 8ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold,
                                &threshold);
      }
    }
    return true;
  }();
  return usable;
}

inline cudaError_t PooledMalloc(void** const pointer, std::size_t const bytes) {
  if (!MemoryPoolsUsable()) {
    return cudaMalloc(pointer, bytes);
  }
  cudaError_t const error = cudaMallocAsync(pointer, bytes, cudaStreamLegacy);
  if (error != cudaSuccess) {
    return error;
  }
  return cudaStreamSynchronize(cudaStreamLegacy);
}

inline cudaError_t PooledFree(void* const pointer) {
  if (!MemoryPoolsUsable()) {
    return cudaFree(pointer);
  }
  cudaError_t const error = cudaDeviceSynchronize();
  if (error != cudaSuccess) {
    return error;
  }
  return cudaFreeAsync(pointer, cudaStreamLegacy);
}

// A grow-only device allocation.
template<typename T>
struct DeviceBuffer {
  T* data = nullptr;
  std::size_t capacity = 0;

  ~DeviceBuffer() {
    if (data != nullptr) {
      PooledFree(data);
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
      PB_CUDA(PooledFree(data));
      data = nullptr;
      capacity = 0;
    }
    std::size_t const wanted = n + n / 4;
    void* pointer = nullptr;
    PB_CUDA(PooledMalloc(&pointer, wanted * sizeof(T)));
    data = static_cast<T*>(pointer);
    capacity = wanted;
    return PB_OK;
  }

  void Swap(DeviceBuffer& other) {
    T* const d = data;
    std::size_t const c = capacity;
    data = other.data;
    capacity = other.capacity;
    other.data = d;
    other.capacity = c;
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
  // [segment][xyz][18]: the device layout (the host API's [segment][18][3]
  // transposed), the slots above the degree holding -0.0.
  std::vector<double> coefficients;
};

}  // namespace pb

namespace pb {

// Everything one call on an ephemeris needs besides the ephemeris itself:
// scratch buffers, staging, events, helper streams.  Calls lease a workspace
// from the ephemeris' pool for their duration, so that calls on distinct
// objects (two flows, a flow and a query) run concurrently
// (physics/ephemeris.hpp:64-66).
struct FlowWorkspace {
  // The stream of the host-pointer entry points (the *_device ones run on the
  // caller's).  Non-blocking: it does not synchronise with the legacy default
  // stream.
  cudaStream_t stream = nullptr;
  // Recorded on the stream of the last user when the lease ends; the next user
  // makes its stream wait for it (the asynchronous *_device queries return
  // while their kernels still read the scratch).
  cudaEvent_t last_use = nullptr;
  bool used = false;

  DeviceBuffer<double> stage_times;
  DeviceBuffer<double> stage_table;
  DeviceBuffer<double> scratch_a, scratch_b, scratch_c;
  DeviceBuffer<double> burns_scratch;
  // [0]: OR of the right-hand-side status codes; [1]: path flags
  // (kPathGenericHarmonics).
  DeviceBuffer<int32_t> status_word;
  // Per launch of SrknFlowSpecializedKernel, per block: completed there.
  DeviceBuffer<int32_t> block_done;
  DeviceBuffer<long long> counters;
  DeviceBuffer<int32_t> frenet_scratch;
  PinnedBuffer<double> pinned_times;
  PinnedBuffer<int32_t> pinned_status;

  // Adaptive flow: scheduling keys, execution order, sort storage, queue head.
  DeviceBuffer<float> adaptive_cost;
  DeviceBuffer<int> adaptive_order;
  DeviceBuffer<unsigned char> adaptive_sort_storage;
  DeviceBuffer<unsigned long long> adaptive_queue;
  DeviceBuffer<int32_t> adaptive_votes;

  // CUDA-event timing of the dominant kernel of a flow call.
  std::vector<cudaEvent_t> timing_events;
  // Helper streams of the sliced flow launches and their join events.
  std::vector<cudaStream_t> flow_streams;
  std::vector<cudaEvent_t> flow_events;

  FlowWorkspace() = default;
  FlowWorkspace(FlowWorkspace const&) = delete;
  FlowWorkspace& operator=(FlowWorkspace const&) = delete;
  ~FlowWorkspace() {
    for (cudaEvent_t event : timing_events) {
      cudaEventDestroy(event);
    }
    for (cudaEvent_t event : flow_events) {
      cudaEventDestroy(event);
    }
    for (cudaStream_t helper : flow_streams) {
      cudaStreamDestroy(helper);
    }
    if (last_use != nullptr) {
      cudaEventDestroy(last_use);
    }
    if (stream != nullptr) {
      cudaStreamDestroy(stream);
    }
  }
};

}  // namespace pb

// The opaque handle of the C ABI.
struct pb_ephemeris {
  int device = 0;
  // Distinguishes handles whose storage the allocator reuses (the constant
  // banks are keyed by it).
  std::uint64_t const serial = pb::NextSerial();
  pb::HostModel model;

  // The ContinuousTrajectory segments: host copy (internal order) and device
  // arrays.  Readers (flows, queries) hold `segments_mutex` shared for the
  // whole call; pb_ephemeris_append_segments and the upload hold it exclusive.
  mutable std::shared_mutex segments_mutex;
  std::vector<pb::HostTrajectory> trajectories;
  bool segments_dirty = true;
  bool uniform_segments = false;  // Set by UploadSegments.
  bool uniform_broken = false;    // Some boundary differs between bodies.
  std::size_t uniform_prefix = 0; // Boundaries compared so far.
  // Device layout: body b's segments at [b * segment_capacity, ... + count).
  int32_t segment_capacity = 0;
  std::vector<int32_t> uploaded_segments;  // Per body.

  pb::DeviceBuffer<pb::DeviceBody> bodies;
  pb::DeviceBuffer<pb::Damping> dampings;
  pb::DeviceBuffer<double> cos, sin, normalization;
  pb::DeviceBuffer<int32_t> internal_of_caller;
  pb::DeviceBuffer<int32_t> segment_offset, segment_end;
  pb::DeviceBuffer<double> segment_t_max, segment_origin, segment_coefficients;
  pb::DeviceBuffer<int32_t> segment_degree;

  // Workspaces not in use.
  std::mutex pool_mutex;
  std::vector<std::unique_ptr<pb::FlowWorkspace>> idle_workspaces;

  // Cooperative cancellation (RETURN_IF_STOPPED): set by
  // pb_ephemeris_request_stop from any thread; the fixed-step flows poll the
  // host flag between chunks of steps, the adaptive kernel polls the device
  // word after every accepted step.
  std::atomic<int> stop_requested{0};
  int32_t* stop_word_device = nullptr;
  cudaStream_t control_stream = nullptr;

  // The hot oblate body of the fixed-step flow kernels (internal index, -1:
  // none), chosen from the first probe of the first flow and kept until a flow
  // reports kPathGenericHarmonics.
  std::mutex hot_mutex;
  bool hot_body_valid = false;
  int hot_body = -1;
  int warm_body = -1;  // See HotBody, pb_kernels_massless.cuh.

  // Measurement of the last completed flow call.
  mutable std::mutex statistics_mutex;
  double last_flow_kernel_ms = 0;
  std::int64_t last_flow_kernel_launches = 0;
  int32_t last_flow_hot_body = -1;  // Caller index.
  int32_t last_flow_path_flags = 0;

  // Live pb_fixed_step_instance / pb_nbody_ensemble handles: the ephemeris
  // refuses to be destroyed under them.
  std::atomic<int> dependants{0};

  ~pb_ephemeris() {
    idle_workspaces.clear();
    if (stop_word_device != nullptr) {
      cudaFree(stop_word_device);
    }
    if (control_stream != nullptr) {
      cudaStreamDestroy(control_stream);
    }
  }

  pb::EphemerisView View() const {
    pb::EphemerisView v;
    v.n_bodies = static_cast<int32_t>(model.bodies.size());
    v.n_oblate = model.n_oblate;
    v.n_frames = model.n_frames;
    v.stage_stride = pb::StageStride(v.n_bodies, v.n_frames);
    v.uniform_segments = uniform_segments ? 1 : 0;
    v.segment_capacity = segment_capacity;
    v.ensemble_tables = 0;
    v.bodies = bodies.data;
    v.dampings = dampings.data;
    v.cos = cos.data;
    v.sin = sin.data;
    v.normalization = normalization.data;
    v.segment_offset = segment_offset.data;
    v.segment_end = segment_end.data;
    v.segment_t_max = segment_t_max.data;
    v.segment_origin = segment_origin.data;
    v.segment_degree = segment_degree.data;
    v.segment_coefficients = segment_coefficients.data;
    v.internal_of_caller = internal_of_caller.data;
    return v;
  }
};

namespace pb {

// A workspace of `e` for the duration of a call whose device work is ordered
// on `stream` (null: the workspace's own stream, see stream()).
class WorkspaceLease {
 public:
  WorkspaceLease() = default;
  ~WorkspaceLease() { Release(); }
  WorkspaceLease(WorkspaceLease const&) = delete;
  WorkspaceLease& operator=(WorkspaceLease const&) = delete;

  // `user_stream_given`: the call runs on the caller's stream
  // (`user_stream`, possibly the legacy default stream) rather than on the
  // workspace's own.
  pb_status_t Acquire(pb_ephemeris* const e, bool const user_stream_given,
                      cudaStream_t const user_stream) {
    e_ = e;
    {
      std::lock_guard<std::mutex> const lock(e->pool_mutex);
      if (!e->idle_workspaces.empty()) {
        w_ = std::move(e->idle_workspaces.back());
        e->idle_workspaces.pop_back();
      }
    }
    if (w_ == nullptr) {
      w_ = std::make_unique<FlowWorkspace>();
      PB_CUDA(cudaStreamCreateWithFlags(&w_->stream, cudaStreamNonBlocking));
      PB_CUDA(cudaEventCreateWithFlags(&w_->last_use, cudaEventDisableTiming));
      PB_RETURN_IF_ERROR(w_->status_word.Reserve(4));
      PB_RETURN_IF_ERROR(w_->pinned_status.Reserve(4));
    }
    stream_ = user_stream_given ? user_stream : w_->stream;
    if (w_->used) {
      PB_CUDA(cudaStreamWaitEvent(stream_, w_->last_use, 0));
    }
    return PB_OK;
  }

  void Release() {
    if (w_ == nullptr) {
      return;
    }
    cudaEventRecord(w_->last_use, stream_);
    w_->used = true;
    std::lock_guard<std::mutex> const lock(e_->pool_mutex);
    e_->idle_workspaces.push_back(std::move(w_));
  }

  FlowWorkspace* get() const { return w_.get(); }
  FlowWorkspace* operator->() const { return w_.get(); }
  cudaStream_t stream() const { return stream_; }

 private:
  pb_ephemeris* e_ = nullptr;
  std::unique_ptr<FlowWorkspace> w_;
  cudaStream_t stream_ = nullptr;
};

// Shared ownership of the segments for the duration of a call, with the device
// copy up to date.
class SegmentsReadLock {
 public:
  SegmentsReadLock() = default;
  SegmentsReadLock(SegmentsReadLock const&) = delete;
  SegmentsReadLock& operator=(SegmentsReadLock const&) = delete;
  pb_status_t Acquire(pb_ephemeris* e);

 private:
  std::shared_lock<std::shared_mutex> lock_;
};

}  // namespace pb

namespace pb {
// Defined in principia_b200.cu.  The callers hold a SegmentsReadLock.
double EphemerisTMin(pb_ephemeris const* e);
double EphemerisTMax(pb_ephemeris const* e);
pb_status_t BuildStageTable(pb_ephemeris* e, FlowWorkspace* w,
                            double const* times_host, int n_times,
                            double* table, cudaStream_t stream);
// Defined in pb_api_continuous_trajectory.cu.
pb_status_t ContinuousTrajectoriesAppendDevice(pb_continuous_trajectories* h,
                                               double time, double const* q,
                                               double const* v,
                                               cudaStream_t stream);
long long ContinuousTrajectoriesSize(pb_continuous_trajectories const* h);
bool ContinuousTrajectoriesEmpty(pb_continuous_trajectories const* h);
}  // namespace pb
