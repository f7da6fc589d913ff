// C ABI: massive-body flows (include/principia_b200.h): ensembles of small
// systems and the large-N all-pairs system.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/principia_b200.h"
#include "generated_tables.h"
#include "pb_kernels_nbody.cuh"
#include "pb_runtime.cuh"

using namespace pb;

namespace pb {

bool MakeSrknTableau(int32_t method, SrknTableau& t);
void ComputeNodes(SrknTableau const& tableau, double* c);

bool MakeMultistepMethod(int32_t const method, MultistepMethod& m) {
  std::memset(&m, 0, sizeof(m));
  if (method == PB_QUINLAN_1999_ORDER_8A) {
    // integrators/methods.hpp:998-1007.
    m.order = 8;
    double const alpha[] = {1.0, -2.0, 2.0, -2.0, 2.0};
    double const beta[] = {0.0, 22081.0, -29418.0, 75183.0, -75212.0};
    double const cho[] = {416173.0,  950684.0, -1025097.0, 1059430.0,
                          -768805.0, 362112.0, -99359.0,   12062.0};
    std::memcpy(m.alpha, alpha, sizeof(alpha));
    std::memcpy(m.beta_numerator, beta, sizeof(beta));
    std::memcpy(m.cho_numerators, cho, sizeof(cho));
    m.beta_denominator = 15120.0;
    m.cho_denominator = 1814400.0;
    return true;
  }
  if (method == PB_QUINLAN_TREMAINE_1990_ORDER_12) {
    // integrators/methods.hpp:1040-1055.
    m.order = 12;
    double const alpha[] = {1.0, -2.0, 2.0, -1.0, 0.0, 0.0, 0.0};
    double const beta[] = {0.0,           90987349.0,   -229596838.0,
                           812627169.0,   -1628539944.0, 2714971338.0,
                           -3041896548.0};
    double const cho[] = {92158447389.0,    301307140046.0,  -554452444015.0,
                          1035372815340.0,  -1505150506950.0, 1655690777412.0,
                          -1363696062582.0, 828085590240.0,  -360089099415.0,
                          106193749950.0,   -19043781851.0,  1569102436.0};
    std::memcpy(m.alpha, alpha, sizeof(alpha));
    std::memcpy(m.beta_numerator, beta, sizeof(beta));
    std::memcpy(m.cho_numerators, cho, sizeof(cho));
    m.beta_denominator = 53222400.0;
    m.cho_denominator = 435891456000.0;
    return true;
  }
  return false;
}

}  // namespace pb

struct pb_nbody_ensemble {
  pb_ephemeris* ephemeris = nullptr;
  int device = 0;
  // The ensemble's own non-blocking stream: distinct ensembles flow
  // concurrently.  One Solve at a time per ensemble.
  cudaStream_t stream = nullptr;
  std::mutex mutex;
  int32_t method = 0;
  bool is_multistep = false;
  SrknTableau tableau;          // The method itself, or the SRKN14A starter.
  double nodes[16];
  MultistepMethod multistep;
  double step = 0;
  long long n_systems = 0;
  int n_bodies = 0;
  DP time{0.0, 0.0};
  // Starter state.
  int previous_steps = 0;       // previous_steps_.size().
  long long startup_step_index = 0;
  int oldest = 0;               // Ring slot of previous_steps_.front().
  bool startup_target_valid = false;
  double startup_target = 0;
  DeviceBuffer<double> state;
  DeviceBuffer<double> displacement_value, displacement_error, acceleration;
  DeviceBuffer<double> frame_times, frames;
  // The multistep steps build the surface frames of the next steps in one go:
  // the times frames.data holds (6 n_frames doubles each), and the entry the
  // next step expects to use.  Any other use of `frames` empties the cache.
  std::vector<double> frame_cache_times;
  std::size_t frame_cache_next = 0;
  // True while a Solve of this ensemble holds the constant bank with the
  // tables of its bodies (c_ensemble).
  bool constant_tables = false;
  EphemerisView View() const {
    EphemerisView view = ephemeris->View();
    view.ensemble_tables = constant_tables ? 1 : 0;
    return view;
  }
  MultistepHistory History() {
    return {displacement_value.data, displacement_error.data,
            acceleration.data};
  }
};

namespace {

// Launch geometry of the ensemble kernels (pb_kernels_nbody.cuh): 32 systems
// per block, one warp per body (or per two bodies).
struct EnsembleLaunch {
  unsigned grid;
  unsigned block;
  std::size_t shared_bytes;
  int own;
};

EnsembleLaunch MakeEnsembleLaunch(pb_nbody_ensemble const* s) {
  EnsembleShape const shape = MakeEnsembleShape(s->n_bodies, s->n_systems);
  EnsembleLaunch launch;
  launch.grid = static_cast<unsigned>((s->n_systems + kEnsembleLanes - 1) /
                                      kEnsembleLanes);
  launch.block = static_cast<unsigned>(shape.warps * kEnsembleLanes);
  launch.shared_bytes = static_cast<std::size_t>(3) * s->n_bodies *
                        kEnsembleLanes * sizeof(double);
  launch.own = shape.own;
  return launch;
}

// Surface-frame axes at the given times -> ensemble->frames.
pb_status_t BuildFrames(pb_nbody_ensemble* s, std::vector<double> const& times,
                        cudaStream_t stream) {
  s->frame_cache_times.clear();
  EphemerisView const view = s->ephemeris->View();
  std::size_t const size =
      std::max<std::size_t>(1, times.size() * 6 * view.n_frames);
  PB_RETURN_IF_ERROR(s->frames.Reserve(size));
  if (view.n_frames == 0) {
    return PB_OK;
  }
  PB_RETURN_IF_ERROR(s->frame_times.Upload(times, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  int const threads = static_cast<int>(times.size()) * view.n_bodies;
  FrameTableKernel<<<(threads + 63) / 64, 64, 0, stream>>>(
      view, s->frame_times.data, static_cast<int>(times.size()),
      s->frames.data);
  PB_LAUNCHED();
  return PB_OK;
}

// One SRKN step of size h from the current time; updates the time.
pb_status_t SrknStep(pb_nbody_ensemble* s, double const h,
                     cudaStream_t stream) {
  // (Without the constant-bank tables: the 14 evaluations of a launch find the
  // body records in the L1 after the first one, and the tables were measured
  // 48 % slower here -- while they gain 14 % on the one evaluation per launch
  // of the multistep step.)
  EphemerisView const view = s->ephemeris->View();
  std::vector<double> times;
  for (int st = s->tableau.first_stage; st < s->tableau.stages; ++st) {
    times.push_back(s->time.value + (s->time.error + s->nodes[st] * h));
  }
  PB_RETURN_IF_ERROR(BuildFrames(s, times, stream));
  EnsembleLaunch const launch = MakeEnsembleLaunch(s);
  auto const kernel = launch.own == 4       ? EnsembleSrknStepKernel<4, 512>
                      : launch.own == 2     ? EnsembleSrknStepKernel<2, 576>
                      : launch.block <= 256 ? EnsembleSrknStepKernel<1, 256>
                                            : EnsembleSrknStepKernel<1, 512>;
  kernel<<<launch.grid, launch.block, launch.shared_bytes, stream>>>(
      view, s->tableau, h, s->frames.data, s->n_systems, s->state.data);
  PB_LAUNCHED();
  Increment(s->time, h);
  return PB_OK;
}

// FillStepFromState into ring slot `slot`.
pb_status_t FillStep(pb_nbody_ensemble* s, int const slot,
                     cudaStream_t stream) {
  EphemerisView const view = s->View();
  PB_RETURN_IF_ERROR(BuildFrames(s, {s->time.value}, stream));
  EnsembleLaunch const launch = MakeEnsembleLaunch(s);
  auto const kernel = launch.own == 4       ? EnsembleFillStepKernel<4, 512>
                      : launch.own == 2     ? EnsembleFillStepKernel<2, 576>
                      : launch.block <= 256 ? EnsembleFillStepKernel<1, 256>
                                            : EnsembleFillStepKernel<1, 512>;
  kernel<<<launch.grid, launch.block, launch.shared_bytes, stream>>>(
      view, s->frames.data, s->n_systems, s->state.data, s->History(), slot);
  PB_LAUNCHED();
  return PB_OK;
}

pb_status_t MultistepStep(pb_nbody_ensemble* s, cudaStream_t stream) {
  EphemerisView const view = s->View();
  DP t = s->time;
  Increment(t, s->step);
  // The frames at the times of this step and of the 63 that follow it (the
  // compensated increments the steps will perform), built once.
  std::size_t index = s->frame_cache_next;
  if (!(index < s->frame_cache_times.size() &&
        s->frame_cache_times[index] == t.value)) {
    std::vector<double> times;
    DP future = s->time;
    for (int k = 0; k < 64; ++k) {
      Increment(future, s->step);
      times.push_back(future.value);
    }
    PB_RETURN_IF_ERROR(BuildFrames(s, times, stream));
    s->frame_cache_times = times;
    index = 0;
  }
  s->frame_cache_next = index + 1;
  double const* const step_frames =
      s->frames.data + index * 6 * static_cast<std::size_t>(view.n_frames);
  EnsembleLaunch const launch = MakeEnsembleLaunch(s);
  auto const kernel =
      s->multistep.order == 8
          ? (launch.own == 4         ? EnsembleMultistepStepKernel<4, 512, 8>
             : launch.own == 2       ? EnsembleMultistepStepKernel<2, 576, 8>
             : launch.block <= 256   ? EnsembleMultistepStepKernel<1, 256, 8>
                                     : EnsembleMultistepStepKernel<1, 512, 8>)
          : (launch.own == 4         ? EnsembleMultistepStepKernel<4, 512, 12>
             : launch.own == 2       ? EnsembleMultistepStepKernel<2, 576, 12>
             : launch.block <= 256   ? EnsembleMultistepStepKernel<1, 256, 12>
                                     : EnsembleMultistepStepKernel<1, 512, 12>);
  kernel<<<launch.grid, launch.block, launch.shared_bytes, stream>>>(
      view, s->multistep, s->step, step_frames, s->n_systems, s->state.data,
      s->History(), s->oldest);
  PB_LAUNCHED();
  s->oldest = (s->oldest + 1) % s->multistep.order;
  s->time = t;
  return PB_OK;
}

}  // namespace

extern "C" {

pb_status_t pb_nbody_ensemble_create(pb_ephemeris_t* e, int32_t method,
                                     double step, int64_t n_systems,
                                     const double* state, double initial_time,
                                     pb_nbody_ensemble_t** ensemble) {
  if (e == nullptr || ensemble == nullptr || n_systems <= 0 ||
      state == nullptr || !(step > 0)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  *ensemble = nullptr;
  int const n_bodies = static_cast<int>(e->model.bodies.size());
  if (n_bodies > kMaxEnsembleBodies) {
    return Fail(PB_INVALID_ARGUMENT, "too many bodies for an ensemble");
  }
  PB_CUDA(cudaSetDevice(e->device));
  auto* s = new pb_nbody_ensemble;
  s->ephemeris = e;
  s->device = e->device;
  s->method = method;
  s->step = step;
  s->n_systems = n_systems;
  s->n_bodies = n_bodies;
  s->time = {initial_time, 0.0};
  if (MakeMultistepMethod(method, s->multistep)) {
    s->is_multistep = true;
    MakeSrknTableau(PB_BLANES_MOAN_2002_SRKN_14A, s->tableau);
  } else if (!MakeSrknTableau(method, s->tableau)) {
    delete s;
    return Fail(PB_INVALID_ARGUMENT, "unsupported fixed-step method");
  }
  ComputeNodes(s->tableau, s->nodes);
  std::size_t const planes =
      static_cast<std::size_t>(n_bodies) * static_cast<std::size_t>(n_systems);
  pb_status_t status = s->state.Upload(state, 12 * planes, nullptr);
  if (status == PB_OK && s->is_multistep) {
    std::size_t const size = 3 * planes * s->multistep.order;
    status = s->displacement_value.Reserve(size);
    if (status == PB_OK) status = s->displacement_error.Reserve(size);
    if (status == PB_OK) status = s->acceleration.Reserve(size);
  }
  if (status == PB_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) {
    status = Fail(PB_INTERNAL, "upload failed");
  }
  if (status == PB_OK &&
      cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) !=
          cudaSuccess) {
    status = Fail(PB_INTERNAL, "stream creation failed");
  }
  if (status != PB_OK) {
    delete s;
    return status;
  }
  e->dependants.fetch_add(1);
  *ensemble = s;
  return PB_OK;
}

void pb_nbody_ensemble_destroy(pb_nbody_ensemble_t* ensemble) {
  PB_API_RANGE();
  if (ensemble != nullptr) {
    // The ensemble's own copy of the ordinal; the ephemeris outlives its
    // dependants (pb_ephemeris_destroy refuses otherwise).
    cudaSetDevice(ensemble->device);
    if (ensemble->stream != nullptr) {
      cudaStreamSynchronize(ensemble->stream);
      cudaStreamDestroy(ensemble->stream);
    }
    ensemble->ephemeris->dependants.fetch_sub(1);
    delete ensemble;
  }
}

// The append_state callback of the instance: optionally copies the appended
// ODE::State values (6 planes) and its time to the host.
struct EnsembleRecorder {
  int64_t capacity;
  double* times;   // capacity.
  double* states;  // capacity * 6 * n_bodies * n_systems.
  int64_t count;
  // Or, instead of the host arrays: ContinuousTrajectory::Append on the device
  // for every body of every system (Ephemeris::AppendMassiveBodiesState,
  // ephemeris_body.hpp:1071-1116).
  pb_continuous_trajectories* sink = nullptr;
};

static pb_status_t RecordStep(pb_nbody_ensemble* s, EnsembleRecorder* recorder,
                              cudaStream_t stream) {
  if (recorder == nullptr) {
    return PB_OK;
  }
  std::size_t const plane =
      static_cast<std::size_t>(s->n_bodies) * static_cast<std::size_t>(s->n_systems);
  if (recorder->sink != nullptr) {
    ++recorder->count;
    return ContinuousTrajectoriesAppendDevice(
        recorder->sink, s->time.value, s->state.data,
        s->state.data + 6 * plane, stream);
  }
  double* const out = recorder->states + recorder->count * 6 * plane;
  PB_CUDA(cudaMemcpyAsync(out, s->state.data, 3 * plane * sizeof(double),
                          cudaMemcpyDeviceToHost, stream));
  PB_CUDA(cudaMemcpyAsync(out + 3 * plane, s->state.data + 6 * plane,
                          3 * plane * sizeof(double), cudaMemcpyDeviceToHost,
                          stream));
  recorder->times[recorder->count] = s->time.value;
  ++recorder->count;
  return PB_OK;
}

// Integrator::Instance::Solve(t_final); with a recorder, returns as soon as
// the recorder is full (call again to continue).
// The constant bank of the ensemble kernels (c_ensemble): the bodies of one
// ephemeris at a time, shared by the Solves that need the same ones.
static ConstantBank& EnsembleConstantBank() {
  static ConstantBank bank;
  return bank;
}

// Resets pb_nbody_ensemble::constant_tables on scope exit.
struct ConstantTablesScope {
  pb_nbody_ensemble* ensemble;
  ~ConstantTablesScope() { ensemble->constant_tables = false; }
};

static pb_status_t AcquireEnsembleConstants(pb_nbody_ensemble* s,
                                            cudaStream_t const stream,
                                            ConstantBankLease& lease) {
  pb_ephemeris const* const e = s->ephemeris;
  int const n = static_cast<int>(e->model.bodies.size());
  if (n > kEnsembleTableBodies) {
    return PB_OK;  // The pointer path.
  }
  ConstantBank& bank = EnsembleConstantBank();
  pb_status_t const status =
      bank.Acquire(e->serial, 1, [&]() -> pb_status_t {
        static EnsembleConstants constants;
        std::memset(&constants, 0, sizeof(constants));
        for (int k = 0; k < kEnsembleTableSize; ++k) {
          constants.normalization[k] = pb_legendre_normalization_factor[k];
        }
        for (int b = 0; b < n; ++b) {
          DeviceBody const& body = e->model.bodies[b];
          EnsembleBodyTables& tables = constants.bodies[b];
          tables.data = body;
          if (!body.is_oblate) {
            continue;
          }
          tables.n_damping = std::min(body.n_damping, kEnsembleTableDampings);
          for (int d = 0; d < tables.n_damping; ++d) {
            tables.degree_damping[d] =
                e->model.dampings[body.damping_offset + d];
          }
          int const size = std::min((body.degree + 1) * (body.degree + 2) / 2,
                                    kEnsembleTableSize);
          for (int k = 0; k < size; ++k) {
            tables.cos[k] = e->model.cos[body.coefficient_offset + k];
            tables.sin[k] = e->model.sin[body.coefficient_offset + k];
          }
        }
        // No kernel reads the bank now (ConstantBank::Acquire).
        PB_CUDA(cudaMemcpyToSymbolAsync(c_ensemble, &constants,
                                        sizeof(constants), 0,
                                        cudaMemcpyHostToDevice, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        return PB_OK;
      });
  if (status == PB_OK) {
    lease.Adopt(&bank);
    s->constant_tables = true;
  }
  return status;
}

static pb_status_t SolveEnsemble(pb_nbody_ensemble* s, double t_final,
                                 int64_t* n_steps,
                                 EnsembleRecorder* recorder) {
  PB_CUDA(cudaSetDevice(s->device));
  cudaStream_t const stream = s->stream;
  // Every return below follows a synchronisation of the stream: the lease
  // outlives the kernels launched under it.
  ConstantBankLease constants;
  ConstantTablesScope const scope{s};
  PB_RETURN_IF_ERROR(AcquireEnsembleConstants(s, stream, constants));
  int64_t steps = 0;
  double const h = s->step;
  // A full recorder, or a stop request (RETURN_IF_STOPPED after every step,
  // symmetric_linear_multistep_integrator_body.hpp:149): either way Solve
  // returns with the state of the last completed step.
  auto full = [&] {
    return (recorder != nullptr && recorder->count >= recorder->capacity) ||
           s->ephemeris->stop_requested.load() != 0;
  };
  auto result = [&]() -> pb_status_t {
    return s->ephemeris->stop_requested.load() != 0
               ? Fail(PB_CANCELLED, "stop requested")
               : PB_OK;
  };
  if (!s->is_multistep) {
    // SymplecticRungeKuttaNyströmIntegrator::Instance::Solve.
    while (!full() &&
           h <= std::abs((t_final - s->time.value) - s->time.error)) {
      PB_RETURN_IF_ERROR(SrknStep(s, h, stream));
      PB_RETURN_IF_ERROR(RecordStep(s, recorder, stream));
      ++steps;
    }
  } else {
    int const order = s->multistep.order;
    if (s->previous_steps < order) {
      // Starter::Solve, integrators/starter_body.hpp:23-74.
      if (s->previous_steps == 0) {
        PB_RETURN_IF_ERROR(FillStep(s, 0, stream));
        s->previous_steps = 1;
      }
      double const startup_step = h / kStartupStepDivisor;
      if (!s->startup_target_valid) {
        // The target of the startup instance is fixed when Starter::Solve is
        // entered; a recorder-full return resumes the same call.
        s->startup_target = std::min(
            s->time.value + (order - s->previous_steps) * h + h / 2.0, t_final);
        s->startup_target_valid = true;
      }
      double const target = s->startup_target;
      while (!full() &&
             startup_step <=
                 std::abs((target - s->time.value) - s->time.error) &&
             s->previous_steps < order) {
        PB_RETURN_IF_ERROR(SrknStep(s, startup_step, stream));
        if (++s->startup_step_index % kStartupStepDivisor == 0) {
          PB_RETURN_IF_ERROR(FillStep(s, s->previous_steps, stream));
          ++s->previous_steps;
          PB_RETURN_IF_ERROR(RecordStep(s, recorder, stream));
          ++steps;  // instance_->append_state()(state).
        }
      }
      if (full() && s->previous_steps < order) {
        PB_CUDA(cudaStreamSynchronize(stream));
        if (n_steps != nullptr) {
          *n_steps = steps;
        }
        return result();
      }
      s->startup_target_valid = false;
      if (s->previous_steps < order) {
        PB_CUDA(cudaStreamSynchronize(stream));
        if (n_steps != nullptr) {
          *n_steps = steps;
        }
        return result();  // "we'll continue the next time Solve is called".
      }
      s->oldest = 0;
    }
    while (!full() && h <= (t_final - s->time.value) - s->time.error) {
      PB_RETURN_IF_ERROR(MultistepStep(s, stream));
      PB_RETURN_IF_ERROR(RecordStep(s, recorder, stream));
      ++steps;
    }
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  if (n_steps != nullptr) {
    *n_steps = steps;
  }
  return result();
}

pb_status_t pb_nbody_ensemble_solve(pb_nbody_ensemble_t* s, double t_final,
                                    int64_t* n_steps) {
  PB_API_RANGE();
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null ensemble");
  }
  std::lock_guard<std::mutex> const lock(s->mutex);
  return SolveEnsemble(s, t_final, n_steps, nullptr);
}

pb_status_t pb_nbody_ensemble_solve_recording(pb_nbody_ensemble_t* s,
                                              double t_final, int64_t capacity,
                                              double* times, double* states,
                                              int64_t* n_recorded) {
  if (s == nullptr || capacity <= 0 || times == nullptr ||
      states == nullptr || n_recorded == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_API_RANGE();
  std::lock_guard<std::mutex> const lock(s->mutex);
  EnsembleRecorder recorder{capacity, times, states, 0};
  pb_status_t const status = SolveEnsemble(s, t_final, nullptr, &recorder);
  *n_recorded = recorder.count;
  return status;
}

pb_status_t pb_nbody_ensemble_solve_to_trajectories(
    pb_nbody_ensemble_t* s, double t_final,
    pb_continuous_trajectories_t* trajectories, int64_t* n_steps) {
  if (s == nullptr || trajectories == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  long long const plane =
      static_cast<long long>(s->n_bodies) * static_cast<long long>(s->n_systems);
  if (ContinuousTrajectoriesSize(trajectories) != plane) {
    return Fail(PB_INVALID_ARGUMENT,
                "one trajectory per body of every system is needed");
  }
  PB_API_RANGE();
  std::lock_guard<std::mutex> const lock(s->mutex);
  PB_CUDA(cudaSetDevice(s->device));
  if (ContinuousTrajectoriesEmpty(trajectories)) {
    // The initial state, as the Ephemeris constructor appends it
    // (ephemeris_body.hpp:136).
    PB_RETURN_IF_ERROR(ContinuousTrajectoriesAppendDevice(
        trajectories, s->time.value, s->state.data, s->state.data + 6 * plane,
        s->stream));
  }
  EnsembleRecorder recorder{std::numeric_limits<int64_t>::max(), nullptr,
                            nullptr, 0, trajectories};
  pb_status_t const status = SolveEnsemble(s, t_final, n_steps, &recorder);
  if (status != PB_OK) {
    return status;
  }
  return pb_continuous_trajectories_status(trajectories);
}

pb_status_t pb_nbody_ensemble_get_state(pb_nbody_ensemble_t* s, double* state,
                                        double* time) {
  if (s == nullptr || state == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  std::lock_guard<std::mutex> const lock(s->mutex);
  PB_CUDA(cudaSetDevice(s->device));
  std::size_t const size = 12 * static_cast<std::size_t>(s->n_bodies) *
                           static_cast<std::size_t>(s->n_systems);
  PB_CUDA(cudaMemcpy(state, s->state.data, size * sizeof(double),
                     cudaMemcpyDeviceToHost));
  if (time != nullptr) {
    time[0] = s->time.value;
    time[1] = s->time.error;
  }
  return PB_OK;
}

}  // extern "C"
