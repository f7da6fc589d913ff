// C ABI: Ephemeris::FlowWithAdaptiveStep, batched (include/principia_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <limits>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/principia_b200.h"
#include "pb_kernels_adaptive.cuh"
#include "pb_runtime.cuh"

using namespace pb;

namespace {

// Resident blocks per SM of the adaptive kernel (register budget 255 / 168;
// measured: 2 blocks without spills beat 3 blocks with).
int AdaptiveMinBlocks() {
  static int const value = [] {
    char const* const text = std::getenv("PB_ADAPTIVE_MIN_BLOCKS");
    int const v = text != nullptr ? std::atoi(text) : 2;
    return v == 3 ? 3 : 2;
  }();
  return value;
}

}  // namespace

extern "C" {

pb_status_t pb_flow_with_adaptive_step_device(
    pb_ephemeris_t* e, const pb_adaptive_step_flow_t* flow,
    void* stream_pointer) {
  std::lock_guard<std::recursive_mutex> const flow_lock(FlowMutex());
  if (e == nullptr || flow == nullptr || flow->n < 0 ||
      (flow->n > 0 && (flow->state == nullptr || flow->time == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  if (flow->method != PB_DORMAND_ELMIKKAWY_PRINCE_1986_RKN_434FM &&
      flow->method != PB_FINE_1987_RKNG_34) {
    return Fail(PB_INVALID_ARGUMENT,
                "adaptive massless flow supports "
                "DormandElMikkawyPrince1986RKN434FM and Fine1987RKNG34");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = static_cast<cudaStream_t>(stream_pointer);
  PB_RETURN_IF_ERROR(UploadSegments(e, stream));
  AdaptiveParameters parameters;
  parameters.t_requested = flow->t_final;
  // FlowODEWithAdaptiveStep: t_final = min(t, t_max()) (after Prolong, which is
  // the facade's job here).
  parameters.t_final = std::min(flow->t_final, EphemerisTMax(e));
  parameters.length_integration_tolerance = flow->length_integration_tolerance;
  parameters.speed_integration_tolerance = flow->speed_integration_tolerance;
  parameters.max_steps = flow->max_steps;
  parameters.step_log_capacity =
      flow->step_log != nullptr || flow->state_log != nullptr
          ? flow->step_log_capacity
          : 0;
  if (flow->n == 0) {
    return PB_OK;
  }
  if (flow->n > std::numeric_limits<int>::max()) {
    return Fail(PB_INVALID_ARGUMENT, "too many trajectories in one call");
  }
  int const n = static_cast<int>(flow->n);
  int const n_bodies = static_cast<int>(e->model.bodies.size());
  if (n_bodies > kMaxConstantBodies) {
    return Fail(PB_INVALID_ARGUMENT, "too many bodies for the flow kernel");
  }
  static BodyScalars scalars;
  for (int b = 0; b < n_bodies; ++b) {
    DeviceBody const& body = e->model.bodies[b];
    scalars.mu[b] = body.mu;
    scalars.collision_radius[b] = body.collision_radius;
    scalars.outer_threshold_degree_2[b] =
        body.is_oblate && body.n_damping > 2
            ? e->model.dampings[body.damping_offset + 2].outer_threshold
            : -std::numeric_limits<double>::infinity();
  }
  PB_CUDA(cudaMemcpyToSymbolAsync(c_adaptive_bodies, &scalars, sizeof(scalars),
                                  0, cudaMemcpyHostToDevice, stream));
  PB_CUDA(cudaStreamSynchronize(stream));

  // Execution order: decreasing estimated cost.
  PB_RETURN_IF_ERROR(e->adaptive_cost.Reserve(2 * static_cast<std::size_t>(n)));
  PB_RETURN_IF_ERROR(e->adaptive_order.Reserve(2 * static_cast<std::size_t>(n)));
  PB_RETURN_IF_ERROR(e->adaptive_queue.Reserve(1));
  float* const cost_in = e->adaptive_cost.data;
  float* const cost_out = cost_in + n;
  int* const index_in = e->adaptive_order.data;
  int* const order = index_in + n;
  std::size_t sort_bytes = 0;
  PB_CUDA(cub::DeviceRadixSort::SortPairsDescending(
      nullptr, sort_bytes, cost_in, cost_out, index_in, order, n, 0, 32,
      stream));
  PB_RETURN_IF_ERROR(e->adaptive_sort_storage.Reserve(sort_bytes));
  sort_bytes = e->adaptive_sort_storage.capacity;
  AdaptiveCostKernel<<<(n + 127) / 128, 128, 0, stream>>>(
      e->View(), parameters.t_final, flow->n, flow->state, flow->time, cost_in,
      index_in);
  PB_LAUNCHED();
  PB_CUDA(cub::DeviceRadixSort::SortPairsDescending(
      e->adaptive_sort_storage.data, sort_bytes, cost_in, cost_out, index_in,
      order, n, 0, 32, stream));
  PB_CUDA(cudaMemsetAsync(e->adaptive_queue.data, 0,
                          sizeof(unsigned long long), stream));
  int sm_count = 0;
  PB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount,
                                 e->device));
  int const min_blocks = AdaptiveMinBlocks();
  unsigned const blocks = static_cast<unsigned>(
      std::min<long long>((flow->n + kAdaptiveThreads - 1) / kAdaptiveThreads,
                          static_cast<long long>(sm_count) * min_blocks));
  while (e->timing_events.size() < 2) {
    cudaEvent_t event;
    PB_CUDA(cudaEventCreate(&event));
    e->timing_events.push_back(event);
  }
  PB_CUDA(cudaEventRecord(e->timing_events[0], stream));
  DownsamplerView sink{};
  if (flow->sink != nullptr) {
    sink = PbDownsamplerView(flow->sink);
    if (sink.n != flow->n) {
      return Fail(PB_INVALID_ARGUMENT,
                  "the sink must have one timeline per trajectory");
    }
  }
  auto const kernel =
      flow->method == PB_FINE_1987_RKNG_34
          ? (min_blocks == 2 ? ErknFlowKernel<Fine1987RKNG34, 2>
                             : ErknFlowKernel<Fine1987RKNG34, 3>)
          : (min_blocks == 2
                 ? ErknFlowKernel<DormandElMikkawyPrince1986RKN434FM, 2>
                 : ErknFlowKernel<DormandElMikkawyPrince1986RKN434FM, 3>);
  kernel<<<blocks, kAdaptiveThreads, 0, stream>>>(
      e->View(), parameters, flow->n, order, e->adaptive_queue.data,
      flow->state, flow->time, flow->status,
      reinterpret_cast<long long*>(flow->accepted_steps),
      reinterpret_cast<long long*>(flow->rejected_steps),
      reinterpret_cast<long long*>(flow->evaluations), flow->step_log,
      flow->state_log, sink, flow->burns);
  PB_LAUNCHED();
  PB_CUDA(cudaEventRecord(e->timing_events[1], stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  float elapsed = 0;
  PB_CUDA(cudaEventElapsedTime(&elapsed, e->timing_events[0],
                               e->timing_events[1]));
  e->last_flow_kernel_ms = elapsed;
  e->last_flow_kernel_launches = 1;
  return PB_OK;
}

pb_status_t pb_flow_with_adaptive_step(pb_ephemeris_t* e,
                                       const pb_adaptive_step_flow_t* flow) {
  std::lock_guard<std::recursive_mutex> const flow_lock(FlowMutex());
  if (e == nullptr || flow == nullptr || flow->n < 0 ||
      (flow->n > 0 && (flow->state == nullptr || flow->time == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = nullptr;
  std::size_t const n = static_cast<std::size_t>(flow->n);
  DeviceBuffer<double> state, time, step_log, state_log;
  DeviceBuffer<int32_t> status;
  DeviceBuffer<long long> counters;
  PB_RETURN_IF_ERROR(state.Upload(flow->state, 12 * n, stream));
  PB_RETURN_IF_ERROR(time.Upload(flow->time, 2 * n, stream));
  PB_RETURN_IF_ERROR(status.Reserve(n + 1));
  PB_RETURN_IF_ERROR(counters.Reserve(3 * n + 1));
  pb_adaptive_step_flow_t device_flow = *flow;
  DeviceBuffer<double> burns;
  if (flow->burns != nullptr) {
    PB_RETURN_IF_ERROR(burns.Upload(flow->burns, PB_BURN_PLANES * n, stream));
    device_flow.burns = burns.data;
  }
  device_flow.state = state.data;
  device_flow.time = time.data;
  device_flow.status = status.data;
  device_flow.accepted_steps = reinterpret_cast<int64_t*>(counters.data);
  device_flow.rejected_steps = reinterpret_cast<int64_t*>(counters.data + n);
  device_flow.evaluations = reinterpret_cast<int64_t*>(counters.data + 2 * n);
  bool const logging = flow->step_log != nullptr && flow->step_log_capacity > 0;
  if (logging) {
    std::size_t const size = static_cast<std::size_t>(flow->step_log_capacity) * n;
    PB_RETURN_IF_ERROR(step_log.Upload(flow->step_log, size, stream));
    device_flow.step_log = step_log.data;
  } else {
    device_flow.step_log = nullptr;
  }
  bool const logging_states =
      flow->state_log != nullptr && flow->step_log_capacity > 0;
  if (logging_states) {
    PB_RETURN_IF_ERROR(state_log.Reserve(
        static_cast<std::size_t>(flow->step_log_capacity) * 6 * n));
    device_flow.state_log = state_log.data;
  } else {
    device_flow.state_log = nullptr;
  }
  if (!logging && !logging_states) {
    device_flow.step_log_capacity = 0;
  }
  PB_RETURN_IF_ERROR(
      pb_flow_with_adaptive_step_device(e, &device_flow, stream));
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(flow->state, state.data, 12 * n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaMemcpyAsync(flow->time, time.data, 2 * n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
    if (flow->status != nullptr) {
      PB_CUDA(cudaMemcpyAsync(flow->status, status.data, n * sizeof(int32_t),
                              cudaMemcpyDeviceToHost, stream));
    }
    if (flow->accepted_steps != nullptr) {
      PB_CUDA(cudaMemcpyAsync(flow->accepted_steps, counters.data,
                              n * sizeof(int64_t), cudaMemcpyDeviceToHost,
                              stream));
    }
    if (flow->rejected_steps != nullptr) {
      PB_CUDA(cudaMemcpyAsync(flow->rejected_steps, counters.data + n,
                              n * sizeof(int64_t), cudaMemcpyDeviceToHost,
                              stream));
    }
    if (flow->evaluations != nullptr) {
      PB_CUDA(cudaMemcpyAsync(flow->evaluations, counters.data + 2 * n,
                              n * sizeof(int64_t), cudaMemcpyDeviceToHost,
                              stream));
    }
    if (logging) {
      PB_CUDA(cudaMemcpyAsync(
          flow->step_log, step_log.data,
          static_cast<std::size_t>(flow->step_log_capacity) * n * sizeof(double),
          cudaMemcpyDeviceToHost, stream));
    }
    if (logging_states) {
      PB_CUDA(cudaMemcpyAsync(
          flow->state_log, state_log.data,
          static_cast<std::size_t>(flow->step_log_capacity) * 6 * n *
              sizeof(double),
          cudaMemcpyDeviceToHost, stream));
    }
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  return PB_OK;
}

}  // extern "C"
