// C ABI: large-N all-pairs point-mass system (include/principia_b200.h),
// sharded by target block across ranks, Quinlan-Tremaine / SRKN steppers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/principia_b200.h"
#include "pb_kernels_allpairs.cuh"
#include "pb_kernels_nbody.cuh"
#include "pb_nccl.hpp"
#include "pb_runtime.cuh"

using namespace pb;

namespace pb {
bool MakeSrknTableau(int32_t method, SrknTableau& t);
void ComputeNodes(SrknTableau const& tableau, double* c);
bool MakeMultistepMethod(int32_t method, MultistepMethod& m);

constexpr int kElementwiseThreads = 256;

// Planes of the owned slice: 12 state planes (as everywhere), then dq (3),
// dv (3), g (3), new displacement value (3) / error (3).
struct AllPairsView {
  long long n;        // All bodies.
  long long n_local;  // Owned targets.
  long long i_begin;
  double* state;      // 12 * n_local.
  double* dq;         // 3 * n_local.
  double* dv;
  double* g;
  double* new_value;
  double* new_error;
  double* sources;    // [n][4]: x, y, z, mu; every rank's slices.
  MultistepHistory history;  // [slot][component][n_local].
};

__global__ void ApBeginStepKernel(AllPairsView const v, double const ha0,
                                  int const first_stage) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= v.n_local) {
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double dq = 0.0;
    double const dv = 0.0;
    if (first_stage == 1) {
      dq += ha0 * (v.state[(6 + c) * v.n_local + i] + dv);
    }
    v.dq[c * v.n_local + i] = dq;
    v.dv[c * v.n_local + i] = dv;
  }
}

// q_stage = q.value + dq into the owned slice of the source array
// (with_increment false: the state's own positions).
__global__ void ApStagePositionsKernel(AllPairsView const v,
                                       bool const with_increment) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= v.n_local) {
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double const q = v.state[c * v.n_local + i];
    v.sources[kApSourceDoubles * (v.i_begin + i) + c] =
        with_increment ? q + v.dq[c * v.n_local + i] : q;
  }
}

__global__ void ApStageUpdateKernel(AllPairsView const v, double const hb,
                                    double const ha) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= v.n_local) {
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    long long const k = c * v.n_local + i;
    double dv = v.dv[k];
    double dq = v.dq[k];
    dv += hb * v.g[k];
    dq += ha * (v.state[(6 + c) * v.n_local + i] + dv);
    v.dv[k] = dv;
    v.dq[k] = dq;
  }
}

__global__ void ApEndStepKernel(AllPairsView const v) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= v.n_local) {
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    DP q{v.state[c * v.n_local + i], v.state[(3 + c) * v.n_local + i]};
    Increment(q, v.dq[c * v.n_local + i]);
    v.state[c * v.n_local + i] = q.value;
    v.state[(3 + c) * v.n_local + i] = q.error;
    DP w{v.state[(6 + c) * v.n_local + i], v.state[(9 + c) * v.n_local + i]};
    Increment(w, v.dv[c * v.n_local + i]);
    v.state[(6 + c) * v.n_local + i] = w.value;
    v.state[(9 + c) * v.n_local + i] = w.error;
  }
}

// Starter::FillStepFromState: displacement = position - DoublePosition(),
// acceleration = the force just evaluated at the state's positions (in g).
__global__ void ApFillStepKernel(AllPairsView const v, int const slot) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= v.n_local) {
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    DP const p{v.state[c * v.n_local + i], v.state[(3 + c) * v.n_local + i]};
    DP const displacement = Subtract(p, DP{0.0, 0.0});
    long long const h = (static_cast<long long>(slot) * 3 + c) * v.n_local + i;
    v.history.displacement_value[h] = displacement.value;
    v.history.displacement_error[h] = displacement.error;
    v.history.acceleration[h] = v.g[c * v.n_local + i];
  }
}

// First half of a multistep step: the new displacement and position.
__global__ void ApPredictKernel(AllPairsView const v,
                                MultistepMethod const method, double const h,
                                int const oldest) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= v.n_local) {
    return;
  }
  int const k = method.order;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    auto index = [&](int const j) {
      return (static_cast<long long>((oldest + j) % k) * 3 + c) * v.n_local + i;
    };
    auto displacement = [&](int const j) {
      long long const hi = index(j);
      return DP{v.history.displacement_value[hi],
                v.history.displacement_error[hi]};
    };
    DP sum = Scale(-method.alpha[0], displacement(0));
    double weighted = method.beta_numerator[0] * v.history.acceleration[index(0)];
    for (int j = 1; j < k / 2; ++j) {
      sum = Subtract(sum, Scale(method.alpha[j], displacement(j)));
      sum = Subtract(sum, Scale(method.alpha[j], displacement(k - j)));
      weighted += method.beta_numerator[j] *
                  (v.history.acceleration[index(j)] +
                   v.history.acceleration[index(k - j)]);
    }
    sum = Subtract(sum, Scale(method.alpha[k / 2], displacement(k / 2)));
    weighted += method.beta_numerator[k / 2] *
                v.history.acceleration[index(k / 2)];
    Increment(sum, h * h * weighted / method.beta_denominator);
    v.new_value[c * v.n_local + i] = sum.value;
    v.new_error[c * v.n_local + i] = sum.error;
    DP const position = Add(DP{0.0, 0.0}, sum);
    v.state[c * v.n_local + i] = position.value;
    v.state[(3 + c) * v.n_local + i] = position.error;
    v.sources[kApSourceDoubles * (v.i_begin + i) + c] = position.value;
  }
}

// Second half: push the step, Cohen-Hubbard-Oesterwinter velocity.
__global__ void ApCorrectKernel(AllPairsView const v,
                                MultistepMethod const method, double const h,
                                int const oldest) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= v.n_local) {
    return;
  }
  int const k = method.order;
  int const newest = oldest;
  int const previous = (oldest + k - 1) % k;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    auto index = [&](int const slot) {
      return (static_cast<long long>(slot) * 3 + c) * v.n_local + i;
    };
    DP const displacement{v.new_value[c * v.n_local + i],
                          v.new_error[c * v.n_local + i]};
    double const a = v.g[c * v.n_local + i];
    v.history.displacement_value[index(newest)] = displacement.value;
    v.history.displacement_error[index(newest)] = displacement.error;
    v.history.acceleration[index(newest)] = a;
    DP const displacement_change = Subtract(
        displacement, DP{v.history.displacement_value[index(previous)],
                         v.history.displacement_error[index(previous)]});
    double velocity =
        (displacement_change.value + displacement_change.error) / h;
    double weighted_accelerations = 0.0;
    for (int j = 0; j < k; ++j) {
      int const slot = (newest + k - j) % k;
      double const aj = j == 0 ? a : v.history.acceleration[index(slot)];
      weighted_accelerations += method.cho_numerators[j] * aj;
    }
    velocity += weighted_accelerations * h / method.cho_denominator;
    v.state[(6 + c) * v.n_local + i] = velocity;
    v.state[(9 + c) * v.n_local + i] = 0.0;
  }
}

}  // namespace pb

struct pb_nbody_allpairs {
  int device = 0;
  bool is_multistep = false;
  SrknTableau tableau;
  double nodes[16];
  MultistepMethod multistep;
  double step = 0;
  long long n = 0, i_begin = 0, i_end = 0;
  DP time{0.0, 0.0};
  int previous_steps = 0;
  long long startup_step_index = 0;
  int oldest = 0;
  pb_exchange_positions_t exchange = nullptr;
  void* exchange_user = nullptr;
  // The communicator of pb_nbody_allpairs_set_nccl (not owned), or null.
  NcclComm nccl_comm = nullptr;
  int32_t arithmetic = PB_ARITHMETIC_EXACT;
  AllPairsSplit split;
  DeviceBuffer<double> state, dq, dv, g, new_value, new_error, sources;
  // Group partials of the force kernel when the sources are cut into more
  // than one group of sub-ranges.
  DeviceBuffer<double> partials;
  DeviceBuffer<double> displacement_value, displacement_error, acceleration;
  std::vector<cudaEvent_t> events;
  std::size_t events_used = 0;
  double last_force_ms = 0;
  double last_exchange_ms = 0;
  long long force_evaluations = 0;

  ~pb_nbody_allpairs() {
    for (cudaEvent_t event : events) {
      cudaEventDestroy(event);
    }
  }
  long long n_local() const { return i_end - i_begin; }
  AllPairsView View() {
    AllPairsView v;
    v.n = n;
    v.n_local = n_local();
    v.i_begin = i_begin;
    v.state = state.data;
    v.dq = dq.data;
    v.dv = dv.data;
    v.g = g.data;
    v.new_value = new_value.data;
    v.new_error = new_error.data;
    v.sources = sources.data;
    v.history = {displacement_value.data, displacement_error.data,
                 acceleration.data};
    return v;
  }
};

namespace {

unsigned ElementwiseGrid(long long const n) {
  return static_cast<unsigned>((n + kElementwiseThreads - 1) /
                               kElementwiseThreads);
}

// Makes the other ranks' slices of the source array visible -- one in-place
// ncclAllGather of the owned [i_begin, i_end) x {x, y, z, mu} slices on the
// force stream, or the caller's callback -- then evaluates the accelerations
// of the owned targets into g.  Events: [exchange begin, force begin, force
// end] per evaluation.
pb_status_t Force(pb_nbody_allpairs* s, cudaStream_t stream) {
  while (s->events.size() < s->events_used + 3) {
    cudaEvent_t event;
    PB_CUDA(cudaEventCreate(&event));
    s->events.push_back(event);
  }
  bool const timed = s->events_used < 3 * 4096;
  if (timed) {
    PB_CUDA(cudaEventRecord(s->events[s->events_used], stream));
  }
  long long const n_local = s->n_local();
  if (s->nccl_comm != nullptr) {
    int const result = Nccl().all_gather(
        s->sources.data + kApSourceDoubles * s->i_begin, s->sources.data,
        static_cast<std::size_t>(kApSourceDoubles * n_local), kNcclFloat64,
        s->nccl_comm, stream);
    if (result != kNcclSuccess) {
      return Fail(PB_INTERNAL,
                  std::string("ncclAllGather: ") +
                      (Nccl().get_error_string != nullptr
                           ? Nccl().get_error_string(result)
                           : "error"));
    }
  } else if (s->exchange != nullptr) {
    pb_status_t const status =
        s->exchange(s->exchange_user, s->sources.data, s->n, s->i_begin,
                    s->i_end, stream);
    if (status != PB_OK) {
      return Fail(status, "the position exchange callback failed");
    }
  }
  if (timed) {
    PB_CUDA(cudaEventRecord(s->events[s->events_used + 1], stream));
  }
  int const groups = s->split.sub_ranges / kApWarps;
  long long const tiles = (n_local + kApTargets - 1) / kApTargets;
  AllPairsForceView view;
  view.n = s->n;
  view.i_begin = s->i_begin;
  view.n_local = n_local;
  view.split = s->split;
  view.sources = s->sources.data;
  view.out = groups == 1 ? s->g.data : s->partials.data;
  unsigned const grid = static_cast<unsigned>(tiles * groups);
  if (s->arithmetic == PB_ARITHMETIC_CONTRACTED) {
    AllPairsAccelerationKernel<ContractedInteraction>
        <<<grid, kApThreads, kApSharedBytes, stream>>>(view);
  } else {
    AllPairsAccelerationKernel<ExactInteraction>
        <<<grid, kApThreads, kApSharedBytes, stream>>>(view);
  }
  PB_LAUNCHED();
  if (groups > 1) {
    AllPairsCombineKernel<<<ElementwiseGrid(3 * n_local), kElementwiseThreads,
                            0, stream>>>(n_local, groups, s->partials.data,
                                         s->g.data);
    PB_LAUNCHED();
  }
  if (timed) {
    PB_CUDA(cudaEventRecord(s->events[s->events_used + 2], stream));
    s->events_used += 3;
  }
  ++s->force_evaluations;
  return PB_OK;
}

pb_status_t SrknStep(pb_nbody_allpairs* s, double const h,
                     cudaStream_t stream) {
  AllPairsView const v = s->View();
  unsigned const grid = ElementwiseGrid(v.n_local);
  ApBeginStepKernel<<<grid, kElementwiseThreads, 0, stream>>>(
      v, h * s->tableau.a[0], s->tableau.first_stage);
  PB_LAUNCHED();
  for (int st = s->tableau.first_stage; st < s->tableau.stages; ++st) {
    ApStagePositionsKernel<<<grid, kElementwiseThreads, 0, stream>>>(v, true);
    PB_LAUNCHED();
    PB_RETURN_IF_ERROR(Force(s, stream));
    ApStageUpdateKernel<<<grid, kElementwiseThreads, 0, stream>>>(
        v, h * s->tableau.b[st], h * s->tableau.a[st]);
    PB_LAUNCHED();
  }
  ApEndStepKernel<<<grid, kElementwiseThreads, 0, stream>>>(v);
  PB_LAUNCHED();
  Increment(s->time, h);
  return PB_OK;
}

pb_status_t FillStep(pb_nbody_allpairs* s, int const slot,
                     cudaStream_t stream) {
  AllPairsView const v = s->View();
  unsigned const grid = ElementwiseGrid(v.n_local);
  ApStagePositionsKernel<<<grid, kElementwiseThreads, 0, stream>>>(v, false);
  PB_LAUNCHED();
  PB_RETURN_IF_ERROR(Force(s, stream));
  ApFillStepKernel<<<grid, kElementwiseThreads, 0, stream>>>(v, slot);
  PB_LAUNCHED();
  return PB_OK;
}

pb_status_t MultistepStep(pb_nbody_allpairs* s, cudaStream_t stream) {
  AllPairsView const v = s->View();
  unsigned const grid = ElementwiseGrid(v.n_local);
  ApPredictKernel<<<grid, kElementwiseThreads, 0, stream>>>(
      v, s->multistep, s->step, s->oldest);
  PB_LAUNCHED();
  PB_RETURN_IF_ERROR(Force(s, stream));
  ApCorrectKernel<<<grid, kElementwiseThreads, 0, stream>>>(
      v, s->multistep, s->step, s->oldest);
  PB_LAUNCHED();
  s->oldest = (s->oldest + 1) % s->multistep.order;
  Increment(s->time, s->step);
  return PB_OK;
}

pb_status_t CollectForceTime(pb_nbody_allpairs* s, cudaStream_t stream) {
  PB_CUDA(cudaStreamSynchronize(stream));
  s->last_force_ms = 0;
  s->last_exchange_ms = 0;
  for (std::size_t k = 0; k + 3 <= s->events_used; k += 3) {
    float elapsed = 0;
    PB_CUDA(cudaEventElapsedTime(&elapsed, s->events[k], s->events[k + 1]));
    s->last_exchange_ms += elapsed;
    PB_CUDA(cudaEventElapsedTime(&elapsed, s->events[k + 1], s->events[k + 2]));
    s->last_force_ms += elapsed;
  }
  s->events_used = 0;
  return PB_OK;
}

}  // namespace

extern "C" {

pb_status_t pb_nbody_allpairs_create(
    int32_t cuda_device, int32_t method, double step, int64_t n,
    int64_t i_begin, int64_t i_end, const double* gravitational_parameters,
    const double* q, const double* v, double initial_time,
    pb_exchange_positions_t exchange, void* exchange_user,
    pb_nbody_allpairs_t** system) {
  if (system == nullptr || n <= 0 || i_begin < 0 || i_end > n ||
      i_begin >= i_end || gravitational_parameters == nullptr ||
      q == nullptr || v == nullptr || !(step > 0)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  *system = nullptr;
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0) {
    return Fail(PB_FAILED_PRECONDITION,
                "no CUDA device; principia_b200 has no CPU fallback");
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  auto* s = new pb_nbody_allpairs;
  s->device = cuda_device;
  s->step = step;
  s->n = n;
  s->i_begin = i_begin;
  s->i_end = i_end;
  s->time = {initial_time, 0.0};
  s->exchange = exchange;
  s->exchange_user = exchange_user;
  if (MakeMultistepMethod(method, s->multistep)) {
    s->is_multistep = true;
    MakeSrknTableau(PB_BLANES_MOAN_2002_SRKN_14A, s->tableau);
  } else if (!MakeSrknTableau(method, s->tableau)) {
    delete s;
    return Fail(PB_INVALID_ARGUMENT, "unsupported fixed-step method");
  }
  ComputeNodes(s->tableau, s->nodes);
  std::size_t const n_local = static_cast<std::size_t>(i_end - i_begin);
  s->split = MakeAllPairsSplit(n);
  {
    // The dynamic shared memory of the force kernel (stages + partials +
    // barriers) is below the 48 KiB default; opt in all the same so that a
    // deeper ring only needs the constant changed.
    cudaError_t error = cudaFuncSetAttribute(
        AllPairsAccelerationKernel<ExactInteraction>,
        cudaFuncAttributeMaxDynamicSharedMemorySize, kApSharedBytes);
    if (error == cudaSuccess) {
      error = cudaFuncSetAttribute(
          AllPairsAccelerationKernel<ContractedInteraction>,
          cudaFuncAttributeMaxDynamicSharedMemorySize, kApSharedBytes);
    }
    if (error != cudaSuccess) {
      delete s;
      return Fail(PB_INTERNAL, cudaGetErrorString(error));
    }
  }
  std::vector<double> state(12 * n_local, 0.0);
  for (int c = 0; c < 3; ++c) {
    std::memcpy(&state[c * n_local], q + c * n + i_begin,
                n_local * sizeof(double));
    std::memcpy(&state[(6 + c) * n_local], v + c * n + i_begin,
                n_local * sizeof(double));
  }
  // [n][4]: x, y, z, mu.
  std::vector<double> sources(kApSourceDoubles * static_cast<std::size_t>(n));
  for (int64_t j = 0; j < n; ++j) {
    for (int c = 0; c < 3; ++c) {
      sources[kApSourceDoubles * j + c] = q[c * n + j];
    }
    sources[kApSourceDoubles * j + 3] = gravitational_parameters[j];
  }
  pb_status_t status = s->state.Upload(state, nullptr);
  if (status == PB_OK) status = s->sources.Upload(sources, nullptr);
  if (status == PB_OK && s->split.sub_ranges > kApWarps) {
    std::size_t const tiles = (n_local + kApTargets - 1) / kApTargets;
    status = s->partials.Reserve(tiles * (s->split.sub_ranges / kApWarps) * 3 *
                                 kApTargets);
  }
  if (status == PB_OK) status = s->dq.Reserve(3 * n_local);
  if (status == PB_OK) status = s->dv.Reserve(3 * n_local);
  if (status == PB_OK) status = s->g.Reserve(3 * n_local);
  if (status == PB_OK) status = s->new_value.Reserve(3 * n_local);
  if (status == PB_OK) status = s->new_error.Reserve(3 * n_local);
  if (status == PB_OK && s->is_multistep) {
    std::size_t const size = 3 * n_local * s->multistep.order;
    status = s->displacement_value.Reserve(size);
    if (status == PB_OK) status = s->displacement_error.Reserve(size);
    if (status == PB_OK) status = s->acceleration.Reserve(size);
  }
  if (status == PB_OK && cudaDeviceSynchronize() != cudaSuccess) {
    status = Fail(PB_INTERNAL, "upload failed");
  }
  if (status != PB_OK) {
    delete s;
    return status;
  }
  *system = s;
  return PB_OK;
}

void pb_nbody_allpairs_destroy(pb_nbody_allpairs_t* system) {
  if (system != nullptr) {
    cudaSetDevice(system->device);
    delete system;
  }
}

pb_status_t pb_nbody_allpairs_solve(pb_nbody_allpairs_t* s, double t_final,
                                    int64_t* n_steps,
                                    int64_t* n_force_evaluations) {
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null system");
  }
  PB_CUDA(cudaSetDevice(s->device));
  cudaStream_t const stream = nullptr;
  int64_t steps = 0;
  long long const evaluations_before = s->force_evaluations;
  double const h = s->step;
  bool finished_startup = true;
  if (!s->is_multistep) {
    while (h <= std::abs((t_final - s->time.value) - s->time.error)) {
      PB_RETURN_IF_ERROR(SrknStep(s, h, stream));
      ++steps;
    }
  } else {
    int const order = s->multistep.order;
    if (s->previous_steps < order) {
      if (s->previous_steps == 0) {
        PB_RETURN_IF_ERROR(FillStep(s, 0, stream));
        s->previous_steps = 1;
      }
      double const startup_step = h / kStartupStepDivisor;
      double const target =
          std::min(s->time.value + (order - s->previous_steps) * h + h / 2.0,
                   t_final);
      while (startup_step <=
                 std::abs((target - s->time.value) - s->time.error) &&
             s->previous_steps < order) {
        PB_RETURN_IF_ERROR(SrknStep(s, startup_step, stream));
        if (++s->startup_step_index % kStartupStepDivisor == 0) {
          PB_RETURN_IF_ERROR(FillStep(s, s->previous_steps, stream));
          ++s->previous_steps;
          ++steps;
        }
      }
      finished_startup = s->previous_steps == order;
      if (finished_startup) {
        s->oldest = 0;
      }
    }
    while (finished_startup &&
           h <= (t_final - s->time.value) - s->time.error) {
      PB_RETURN_IF_ERROR(MultistepStep(s, stream));
      ++steps;
    }
  }
  PB_RETURN_IF_ERROR(CollectForceTime(s, stream));
  if (n_steps != nullptr) {
    *n_steps = steps;
  }
  if (n_force_evaluations != nullptr) {
    *n_force_evaluations = s->force_evaluations - evaluations_before;
  }
  return PB_OK;
}

pb_status_t pb_nbody_allpairs_get_state(pb_nbody_allpairs_t* s, double* q,
                                        double* v, double* time) {
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null system");
  }
  PB_CUDA(cudaSetDevice(s->device));
  std::size_t const n_local = static_cast<std::size_t>(s->n_local());
  if (q != nullptr) {
    PB_CUDA(cudaMemcpy(q, s->state.data, 3 * n_local * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  if (v != nullptr) {
    PB_CUDA(cudaMemcpy(v, s->state.data + 6 * n_local,
                       3 * n_local * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (time != nullptr) {
    time[0] = s->time.value;
    time[1] = s->time.error;
  }
  return PB_OK;
}

pb_status_t pb_nbody_allpairs_accelerations(pb_nbody_allpairs_t* s, double* a) {
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null system");
  }
  PB_CUDA(cudaSetDevice(s->device));
  cudaStream_t const stream = nullptr;
  AllPairsView const v = s->View();
  ApStagePositionsKernel<<<ElementwiseGrid(v.n_local), kElementwiseThreads, 0,
                           stream>>>(v, false);
  PB_LAUNCHED();
  PB_RETURN_IF_ERROR(Force(s, stream));
  PB_RETURN_IF_ERROR(CollectForceTime(s, stream));
  if (a != nullptr) {
    PB_CUDA(cudaMemcpy(a, s->g.data, 3 * v.n_local * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  return PB_OK;
}

double pb_nbody_allpairs_last_force_ms(const pb_nbody_allpairs_t* s) {
  return s->last_force_ms;
}

double pb_nbody_allpairs_last_exchange_ms(const pb_nbody_allpairs_t* s) {
  return s->last_exchange_ms;
}

pb_status_t pb_nbody_allpairs_set_arithmetic(pb_nbody_allpairs_t* s,
                                             int32_t arithmetic) {
  if (s == nullptr || (arithmetic != PB_ARITHMETIC_EXACT &&
                       arithmetic != PB_ARITHMETIC_CONTRACTED)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  s->arithmetic = arithmetic;
  return PB_OK;
}

pb_status_t pb_nbody_allpairs_set_nccl(pb_nbody_allpairs_t* s,
                                       void* nccl_comm) {
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null system");
  }
  if (nccl_comm == nullptr) {
    s->nccl_comm = nullptr;
    return PB_OK;
  }
  if (!Nccl().loaded) {
    return Fail(PB_FAILED_PRECONDITION,
                "libnccl.so.2 is neither loaded in this process nor on the "
                "library path");
  }
  NcclComm const comm = static_cast<NcclComm>(nccl_comm);
  int world = 0, rank = -1;
  if (Nccl().comm_count(comm, &world) != kNcclSuccess ||
      Nccl().comm_user_rank(comm, &rank) != kNcclSuccess) {
    return Fail(PB_INVALID_ARGUMENT, "not a usable NCCL communicator");
  }
  // The in-place all-gather needs equal contributions in rank order.
  long long const n_local = s->n_local();
  if (n_local * world != s->n || s->i_begin != rank * n_local) {
    return Fail(PB_INVALID_ARGUMENT,
                "the NCCL exchange needs n = world size x block and "
                "i_begin = rank x block");
  }
  s->nccl_comm = comm;
  return PB_OK;
}

pb_status_t pb_nccl_get_unique_id(char* unique_id) {
  if (unique_id == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  if (!Nccl().loaded) {
    return Fail(PB_FAILED_PRECONDITION, "libnccl.so.2 not available");
  }
  NcclUniqueId id;
  int const result = Nccl().get_unique_id(&id);
  if (result != kNcclSuccess) {
    return Fail(PB_INTERNAL, "ncclGetUniqueId failed");
  }
  std::memcpy(unique_id, id.internal, kNcclUniqueIdBytes);
  return PB_OK;
}

pb_status_t pb_nccl_comm_create(const char* unique_id, int32_t world,
                                int32_t rank, int32_t cuda_device,
                                void** nccl_comm) {
  if (unique_id == nullptr || nccl_comm == nullptr || world <= 0 || rank < 0 ||
      rank >= world) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  *nccl_comm = nullptr;
  if (!Nccl().loaded) {
    return Fail(PB_FAILED_PRECONDITION, "libnccl.so.2 not available");
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  NcclUniqueId id;
  std::memcpy(id.internal, unique_id, kNcclUniqueIdBytes);
  NcclComm comm = nullptr;
  int const result = Nccl().comm_init_rank(&comm, world, id, rank);
  if (result != kNcclSuccess) {
    return Fail(PB_INTERNAL,
                std::string("ncclCommInitRank: ") +
                    (Nccl().get_error_string != nullptr
                         ? Nccl().get_error_string(result)
                         : "error"));
  }
  *nccl_comm = comm;
  return PB_OK;
}

void pb_nccl_comm_destroy(void* nccl_comm) {
  if (nccl_comm != nullptr && Nccl().loaded) {
    Nccl().comm_destroy(static_cast<NcclComm>(nccl_comm));
  }
}

}  // extern "C"
