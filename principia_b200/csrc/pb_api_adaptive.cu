// C ABI: Ephemeris::FlowWithAdaptiveStep, batched (include/principia_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <limits>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include <cstring>

#include "../../include/principia_b200.h"
#include "generated_tables.h"
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

// The constant bank of this translation unit (c_adaptive_bodies).
ConstantBank& AdaptiveConstantBank() {
  static ConstantBank bank;
  return bank;
}

// The flow on device arrays.  The caller holds the workspace and the segments.
pb_status_t AdaptiveFlowDevice(pb_ephemeris_t* e, FlowWorkspace* w,
                               const pb_adaptive_step_flow_t* flow,
                               cudaStream_t stream) {
  AdaptiveParameters parameters;
  parameters.t_requested = flow->t_final;
  // FlowODEWithAdaptiveStep: t_final = min(t, t_max()) (after Prolong, which is
  // the facade's job here).
  parameters.t_final = std::min(flow->t_final, EphemerisTMax(e));
  parameters.t_min = EphemerisTMin(e);
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
  DownsamplerView sink{};
  if (flow->sink != nullptr) {
    sink = PbDownsamplerView(flow->sink);
    if (sink.n != flow->n) {
      return Fail(PB_INVALID_ARGUMENT,
                  "the sink must have one timeline per trajectory");
    }
  }
  // Execution order: decreasing estimated cost.
  PB_RETURN_IF_ERROR(w->adaptive_cost.Reserve(2 * static_cast<std::size_t>(n)));
  PB_RETURN_IF_ERROR(w->adaptive_order.Reserve(2 * static_cast<std::size_t>(n)));
  PB_RETURN_IF_ERROR(w->adaptive_queue.Reserve(1));
  float* const cost_in = w->adaptive_cost.data;
  float* const cost_out = cost_in + n;
  int* const index_in = w->adaptive_order.data;
  int* const order = index_in + n;
  std::size_t sort_bytes = 0;
  PB_CUDA(cub::DeviceRadixSort::SortPairsDescending(
      nullptr, sort_bytes, cost_in, cost_out, index_in, order, n, 0, 32,
      stream));
  PB_RETURN_IF_ERROR(w->adaptive_sort_storage.Reserve(sort_bytes));
  sort_bytes = w->adaptive_sort_storage.capacity;
  PB_RETURN_IF_ERROR(w->adaptive_votes.Reserve(kMaxConstantBodies));
  PB_CUDA(cudaMemsetAsync(w->adaptive_votes.data, 0,
                          kMaxConstantBodies * sizeof(int32_t), stream));
  AdaptiveCostKernel<<<(n + 127) / 128, 128, 0, stream>>>(
      e->View(), parameters.t_final, flow->n, flow->state, flow->time, cost_in,
      index_in, w->adaptive_votes.data);
  PB_LAUNCHED();
  // The deep body: among the oblate bodies whose model fits the tables, the
  // one that is the strongest tide of most trajectories.  Only the speed
  // depends on the choice.
  int deep = -1;
  {
    int32_t votes[kMaxConstantBodies];
    PB_CUDA(cudaMemcpyAsync(votes, w->adaptive_votes.data, sizeof(votes),
                            cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    int32_t best = 0;
    int const n_candidates = std::min(e->model.n_oblate, kAdaptiveTableBodies);
    for (int b = 0; b < n_candidates; ++b) {
      DeviceBody const& body = e->model.bodies[b];
      if (body.degree > kAdaptiveTableDegree &&
          body.degree <= kMaxUnrolledDegree &&
          body.n_damping <= kMaxUnrolledDegree + 1 && votes[b] > best) {
        best = votes[b];
        deep = b;
      }
    }
  }
  ConstantBankLease constants;
  PB_RETURN_IF_ERROR(AdaptiveConstantBank().Acquire(
      e->serial, deep + 1, [&]() -> pb_status_t {
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
        PB_CUDA(cudaMemcpyToSymbolAsync(c_adaptive_bodies, &scalars,
                                        sizeof(scalars), 0,
                                        cudaMemcpyHostToDevice, stream));
        // The oblate bodies' records and degree <= 3 harmonics.
        static AdaptiveConstants tables;
        std::memset(&tables, 0, sizeof(tables));
        tables.n_tables = std::min(e->model.n_oblate, kAdaptiveTableBodies);
        for (int k = 0; k < kAdaptiveTableSize; ++k) {
          tables.normalization[k] = pb_legendre_normalization_factor[k];
        }
        for (int b = 0; b < tables.n_tables; ++b) {
          DeviceBody const& body = e->model.bodies[b];
          AdaptiveBodyTables& entry = tables.bodies[b];
          entry.data = body;
          entry.n_damping = std::min(body.n_damping, kAdaptiveTableDampings);
          for (int d = 0; d < entry.n_damping; ++d) {
            entry.degree_damping[d] =
                e->model.dampings[body.damping_offset + d];
          }
          int const size = std::min((body.degree + 1) * (body.degree + 2) / 2,
                                    kAdaptiveTableSize);
          for (int k = 0; k < size; ++k) {
            entry.cos[k] = e->model.cos[body.coefficient_offset + k];
            entry.sin[k] = e->model.sin[body.coefficient_offset + k];
          }
        }
        PB_CUDA(cudaMemcpyToSymbolAsync(c_adaptive_tables, &tables,
                                        sizeof(tables), 0,
                                        cudaMemcpyHostToDevice, stream));
        static AdaptiveDeepBody deep_body;
        std::memset(&deep_body, 0, sizeof(deep_body));
        deep_body.body = deep;
        if (deep >= 0) {
          DeviceBody const& body = e->model.bodies[deep];
          deep_body.n_damping = body.n_damping;
          for (int d = 0; d < body.n_damping; ++d) {
            deep_body.degree_damping[d] =
                e->model.dampings[body.damping_offset + d];
          }
          int const size = (body.degree + 1) * (body.degree + 2) / 2;
          for (int k = 0; k < size; ++k) {
            deep_body.cos[k] = e->model.cos[body.coefficient_offset + k];
            deep_body.sin[k] = e->model.sin[body.coefficient_offset + k];
            deep_body.normalization[k] = pb_legendre_normalization_factor[k];
          }
        }
        PB_CUDA(cudaMemcpyToSymbolAsync(c_adaptive_deep, &deep_body,
                                        sizeof(deep_body), 0,
                                        cudaMemcpyHostToDevice, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        return PB_OK;
      }));
  constants.Adopt(&AdaptiveConstantBank());

  PB_CUDA(cub::DeviceRadixSort::SortPairsDescending(
      w->adaptive_sort_storage.data, sort_bytes, cost_in, cost_out, index_in,
      order, n, 0, 32, stream));
  PB_CUDA(cudaMemsetAsync(w->adaptive_queue.data, 0,
                          sizeof(unsigned long long), stream));
  int sm_count = 0;
  PB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount,
                                 e->device));
  int const min_blocks = AdaptiveMinBlocks();
  unsigned const blocks = static_cast<unsigned>(
      std::min<long long>((flow->n + kAdaptiveThreads - 1) / kAdaptiveThreads,
                          static_cast<long long>(sm_count) * min_blocks));
  while (w->timing_events.size() < 2) {
    cudaEvent_t event;
    PB_CUDA(cudaEventCreate(&event));
    w->timing_events.push_back(event);
  }
  PB_CUDA(cudaEventRecord(w->timing_events[0], stream));
  // A velocity-dependent intrinsic acceleration takes the generalized Solve
  // (the velocity stages exist only there).
  bool const generalized =
      flow->burns != nullptr && flow->frenet_centres != nullptr;
  if (generalized && flow->method != PB_FINE_1987_RKNG_34) {
    return Fail(PB_INVALID_ARGUMENT,
                "Frenet burns need the generalized integrator "
                "(PB_FINE_1987_RKNG_34)");
  }
  auto const kernel =
      flow->method == PB_FINE_1987_RKNG_34
          ? (generalized
                 ? ErknFlowKernel<Fine1987RKNG34, 2, true>
                 : (min_blocks == 2 ? ErknFlowKernel<Fine1987RKNG34, 2>
                                    : ErknFlowKernel<Fine1987RKNG34, 3>))
          : (min_blocks == 2
                 ? ErknFlowKernel<DormandElMikkawyPrince1986RKN434FM, 2>
                 : ErknFlowKernel<DormandElMikkawyPrince1986RKN434FM, 3>);
  kernel<<<blocks, kAdaptiveThreads, 0, stream>>>(
      e->View(), parameters, flow->n, order, w->adaptive_queue.data,
      e->stop_word_device, flow->state, flow->time, flow->status,
      reinterpret_cast<long long*>(flow->accepted_steps),
      reinterpret_cast<long long*>(flow->rejected_steps),
      reinterpret_cast<long long*>(flow->evaluations), flow->step_log,
      flow->state_log, sink, flow->burns,
      generalized ? flow->frenet_centres : nullptr);
  PB_LAUNCHED();
  PB_CUDA(cudaEventRecord(w->timing_events[1], stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  float elapsed = 0;
  PB_CUDA(cudaEventElapsedTime(&elapsed, w->timing_events[0],
                               w->timing_events[1]));
  {
    std::lock_guard<std::mutex> const lock(e->statistics_mutex);
    e->last_flow_kernel_ms = elapsed;
    e->last_flow_kernel_launches = 1;
  }
  // RETURN_IF_STOPPED (embedded_explicit_runge_kutta_nyström_integrator_body
  // .hpp:243): the trajectories that were cut short carry PB_CANCELLED in their
  // status and the state of their last accepted step.
  return e->stop_requested.load() != 0 ? Fail(PB_CANCELLED, "stop requested")
                                       : PB_OK;
}

pb_status_t ValidateAdaptiveFlow(pb_ephemeris_t* e,
                                 const pb_adaptive_step_flow_t* flow) {
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
  // AdaptiveStepParameters: CHECK_LT(0, max_steps), tolerances positive
  // (physics/ephemeris_body.hpp:62-75).
  if (!(flow->max_steps > 0)) {
    return Fail(PB_INVALID_ARGUMENT, "CHECK_LT(0, max_steps)");
  }
  if (!(flow->length_integration_tolerance > 0) ||
      !(flow->speed_integration_tolerance > 0)) {
    return Fail(PB_INVALID_ARGUMENT, "the tolerances must be positive");
  }
  if (flow->t_final != flow->t_final) {
    return Fail(PB_INVALID_ARGUMENT, "NaN time");
  }
  if (flow->frenet_centres != nullptr && flow->burns == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "Frenet centres without burns");
  }
  return PB_OK;
}

}  // namespace

extern "C" {

pb_status_t pb_flow_with_adaptive_step_device(
    pb_ephemeris_t* e, const pb_adaptive_step_flow_t* flow,
    void* stream_pointer) {
  PB_API_RANGE();
  PB_RETURN_IF_ERROR(ValidateAdaptiveFlow(e, flow));
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = static_cast<cudaStream_t>(stream_pointer);
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, true, stream));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  return AdaptiveFlowDevice(e, w.get(), flow, stream);
}

pb_status_t pb_flow_with_adaptive_step(pb_ephemeris_t* e,
                                       const pb_adaptive_step_flow_t* flow) {
  PB_API_RANGE();
  PB_RETURN_IF_ERROR(ValidateAdaptiveFlow(e, flow));
  PB_CUDA(cudaSetDevice(e->device));
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, false, nullptr));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  cudaStream_t const stream = w.stream();
  std::size_t const n = static_cast<std::size_t>(flow->n);
  // The workspace's grow-only scratch: scratch_a the state (12 n) and the times
  // (2 n), scratch_b the step log, scratch_c the state log, counters the three
  // per-trajectory counts and (as 32-bit words behind them) the statuses.
  PB_RETURN_IF_ERROR(w->scratch_a.Reserve(14 * n + 1));
  PB_RETURN_IF_ERROR(w->counters.Reserve(4 * n + 1));
  double* const state = w->scratch_a.data;
  double* const time = state + 12 * n;
  long long* const counters = w->counters.data;
  int32_t* const status = reinterpret_cast<int32_t*>(counters + 3 * n);
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(state, flow->state, 12 * n * sizeof(double),
                            cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemcpyAsync(time, flow->time, 2 * n * sizeof(double),
                            cudaMemcpyHostToDevice, stream));
  }
  pb_adaptive_step_flow_t device_flow = *flow;
  if (flow->burns != nullptr) {
    PB_RETURN_IF_ERROR(
        w->burns_scratch.Upload(flow->burns, PB_BURN_PLANES * n, stream));
    device_flow.burns = w->burns_scratch.data;
  }
  if (flow->frenet_centres != nullptr) {
    int32_t const n_bodies = static_cast<int32_t>(e->model.bodies.size());
    for (std::size_t i = 0; i < n; ++i) {
      if (flow->frenet_centres[i] >= n_bodies) {
        return Fail(PB_INVALID_ARGUMENT, "Frenet centre out of range");
      }
    }
    PB_RETURN_IF_ERROR(
        w->frenet_scratch.Upload(flow->frenet_centres, n, stream));
    device_flow.frenet_centres = w->frenet_scratch.data;
  }
  device_flow.state = state;
  device_flow.time = time;
  device_flow.status = status;
  device_flow.accepted_steps = reinterpret_cast<int64_t*>(counters);
  device_flow.rejected_steps = reinterpret_cast<int64_t*>(counters + n);
  device_flow.evaluations = reinterpret_cast<int64_t*>(counters + 2 * n);
  bool const logging = flow->step_log != nullptr && flow->step_log_capacity > 0;
  if (logging) {
    std::size_t const size = static_cast<std::size_t>(flow->step_log_capacity) * n;
    PB_RETURN_IF_ERROR(w->scratch_b.Upload(flow->step_log, size, stream));
    device_flow.step_log = w->scratch_b.data;
  } else {
    device_flow.step_log = nullptr;
  }
  bool const logging_states =
      flow->state_log != nullptr && flow->step_log_capacity > 0;
  if (logging_states) {
    PB_RETURN_IF_ERROR(w->scratch_c.Reserve(
        static_cast<std::size_t>(flow->step_log_capacity) * 6 * n));
    device_flow.state_log = w->scratch_c.data;
  } else {
    device_flow.state_log = nullptr;
  }
  if (!logging && !logging_states) {
    device_flow.step_log_capacity = 0;
  }
  pb_status_t const result =
      AdaptiveFlowDevice(e, w.get(), &device_flow, stream);
  if (result != PB_OK && result != PB_CANCELLED) {
    return result;
  }
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(flow->state, state, 12 * n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaMemcpyAsync(flow->time, time, 2 * n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
    if (flow->status != nullptr) {
      PB_CUDA(cudaMemcpyAsync(flow->status, status, n * sizeof(int32_t),
                              cudaMemcpyDeviceToHost, stream));
    }
    if (flow->accepted_steps != nullptr) {
      PB_CUDA(cudaMemcpyAsync(flow->accepted_steps, counters,
                              n * sizeof(int64_t), cudaMemcpyDeviceToHost,
                              stream));
    }
    if (flow->rejected_steps != nullptr) {
      PB_CUDA(cudaMemcpyAsync(flow->rejected_steps, counters + n,
                              n * sizeof(int64_t), cudaMemcpyDeviceToHost,
                              stream));
    }
    if (flow->evaluations != nullptr) {
      PB_CUDA(cudaMemcpyAsync(flow->evaluations, counters + 2 * n,
                              n * sizeof(int64_t), cudaMemcpyDeviceToHost,
                              stream));
    }
    if (logging) {
      PB_CUDA(cudaMemcpyAsync(
          flow->step_log, w->scratch_b.data,
          static_cast<std::size_t>(flow->step_log_capacity) * n * sizeof(double),
          cudaMemcpyDeviceToHost, stream));
    }
    if (logging_states) {
      PB_CUDA(cudaMemcpyAsync(
          flow->state_log, w->scratch_c.data,
          static_cast<std::size_t>(flow->step_log_capacity) * 6 * n *
              sizeof(double),
          cudaMemcpyDeviceToHost, stream));
    }
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  return result;
}

}  // extern "C"
