// C ABI of the B200-native gravitational-flow engine (include/principia_b200.h)
// for the massless-body path: construction, ContinuousTrajectory segments,
// batched accelerations / potential, fixed-step SRKN flow.  The adaptive flow
// and the massive-body flows live in pb_api_adaptive.cu / pb_api_nbody.cu.
//
// Compiled for sm_100a with --fmad=false: IEEE binary64 without contraction,
// like the reference's -ffp-contract=off build.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/principia_b200.h"
#include "pb_contracted.h"
#include "pb_kernels_massless.cuh"
#include "pb_kernels_massless_multistep.cuh"
#include "pb_kernels_massless_specialized.cuh"
#include "pb_runtime.cuh"

using namespace pb;

namespace pb {

// Brings the device copy of the segments up to date: only the segments
// appended since the last upload travel.  Body b's segments live at
// [b * capacity, b * capacity + count_b) of every array; when a body outgrows
// the capacity the arrays are reallocated at twice the size and the old rows
// copied device to device.  The caller holds e->segments_mutex exclusively.
pb_status_t UploadSegments(pb_ephemeris* e) {
  if (!e->segments_dirty) {
    return PB_OK;
  }
  int const n = static_cast<int>(e->trajectories.size());
  cudaStream_t const stream = nullptr;
  if (static_cast<int>(e->uploaded_segments.size()) != n) {
    e->uploaded_segments.assign(n, 0);
  }
  std::size_t longest = 0;
  for (auto const& tr : e->trajectories) {
    longest = std::max(longest, tr.t_max.size());
  }
  if (longest > static_cast<std::size_t>(e->segment_capacity)) {
    // Asynchronous queries may still read the old arrays.
    PB_CUDA(cudaDeviceSynchronize());
    std::size_t const old_capacity = e->segment_capacity;
    std::size_t const capacity =
        std::max<std::size_t>({2 * old_capacity, longest, 64});
    DeviceBuffer<double> t_max, origin, coefficients;
    DeviceBuffer<int32_t> degree;
    PB_RETURN_IF_ERROR(t_max.Reserve(capacity * n));
    PB_RETURN_IF_ERROR(origin.Reserve(capacity * n));
    PB_RETURN_IF_ERROR(degree.Reserve(capacity * n));
    PB_RETURN_IF_ERROR(
        coefficients.Reserve(capacity * n * 3 * kSegmentCoefficients));
    if (old_capacity > 0) {
      auto grow = [&](void* const to, void const* const from,
                      std::size_t const element_bytes) {
        return cudaMemcpy2DAsync(to, capacity * element_bytes, from,
                                 old_capacity * element_bytes,
                                 old_capacity * element_bytes, n,
                                 cudaMemcpyDeviceToDevice, stream);
      };
      PB_CUDA(grow(t_max.data, e->segment_t_max.data, sizeof(double)));
      PB_CUDA(grow(origin.data, e->segment_origin.data, sizeof(double)));
      PB_CUDA(grow(degree.data, e->segment_degree.data, sizeof(int32_t)));
      PB_CUDA(grow(coefficients.data, e->segment_coefficients.data,
                   3 * kSegmentCoefficients * sizeof(double)));
      PB_CUDA(cudaStreamSynchronize(stream));
    }
    e->segment_t_max.Swap(t_max);
    e->segment_origin.Swap(origin);
    e->segment_degree.Swap(degree);
    e->segment_coefficients.Swap(coefficients);
    e->segment_capacity = static_cast<int32_t>(capacity);
  }
  std::size_t const capacity = e->segment_capacity;
  std::vector<int32_t> offset(n), end(n);
  for (int b = 0; b < n; ++b) {
    HostTrajectory const& tr = e->trajectories[b];
    std::size_t const first = e->uploaded_segments[b];
    std::size_t const count = tr.t_max.size() - first;
    offset[b] = static_cast<int32_t>(b * capacity);
    end[b] = static_cast<int32_t>(b * capacity + tr.t_max.size());
    if (count == 0) {
      continue;
    }
    std::size_t const at = b * capacity + first;
    // Pageable host memory: these copies are synchronous with the host.
    PB_CUDA(cudaMemcpyAsync(e->segment_t_max.data + at, tr.t_max.data() + first,
                            count * sizeof(double), cudaMemcpyHostToDevice,
                            stream));
    PB_CUDA(cudaMemcpyAsync(e->segment_origin.data + at,
                            tr.origin.data() + first, count * sizeof(double),
                            cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemcpyAsync(e->segment_degree.data + at,
                            tr.degree.data() + first, count * sizeof(int32_t),
                            cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaMemcpyAsync(
        e->segment_coefficients.data + at * 3 * kSegmentCoefficients,
        tr.coefficients.data() + first * 3 * kSegmentCoefficients,
        count * 3 * kSegmentCoefficients * sizeof(double),
        cudaMemcpyHostToDevice, stream));
    e->uploaded_segments[b] = static_cast<int32_t>(tr.t_max.size());
  }
  // Bodies appended in lock step (Ephemeris::Prolong) share their segment
  // boundaries; kernels then look a time up once for all bodies.  Only the new
  // boundaries are compared; a mismatch is for ever.
  bool same_counts = n > 0;
  for (int b = 1; b < n; ++b) {
    same_counts = same_counts && e->trajectories[b].t_max.size() ==
                                     e->trajectories[0].t_max.size();
  }
  if (same_counts && !e->uniform_broken) {
    std::vector<double> const& reference = e->trajectories[0].t_max;
    for (int b = 1; b < n && !e->uniform_broken; ++b) {
      std::vector<double> const& other = e->trajectories[b].t_max;
      if (!std::equal(reference.begin() + e->uniform_prefix, reference.end(),
                      other.begin() + e->uniform_prefix)) {
        e->uniform_broken = true;
      }
    }
    e->uniform_prefix = reference.size();
  }
  e->uniform_segments = same_counts && !e->uniform_broken;
  // After the data, so that a kernel in flight never sees an end beyond what
  // has arrived.
  PB_RETURN_IF_ERROR(e->segment_offset.Upload(offset, stream));
  PB_RETURN_IF_ERROR(e->segment_end.Upload(end, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  e->segments_dirty = false;
  return PB_OK;
}

pb_status_t SegmentsReadLock::Acquire(pb_ephemeris* const e) {
  for (;;) {
    {
      std::shared_lock<std::shared_mutex> shared(e->segments_mutex);
      if (!e->segments_dirty) {
        lock_ = std::move(shared);
        return PB_OK;
      }
    }
    std::unique_lock<std::shared_mutex> exclusive(e->segments_mutex);
    PB_RETURN_IF_ERROR(UploadSegments(e));
  }
}

double EphemerisTMin(pb_ephemeris const* e) {
  double t = -std::numeric_limits<double>::infinity();
  for (auto const& tr : e->trajectories) {
    if (!tr.has_t_min) {
      return std::numeric_limits<double>::infinity();
    }
    t = std::max(t, tr.t_min);
  }
  return t;
}

double EphemerisTMax(pb_ephemeris const* e) {
  double t = std::numeric_limits<double>::infinity();
  for (auto const& tr : e->trajectories) {
    if (tr.t_max.empty()) {
      return -std::numeric_limits<double>::infinity();
    }
    t = std::min(t, tr.t_max.back());
  }
  return t;
}

// Fills `table` (device, n_times entries) for `times` (host).
pb_status_t BuildStageTable(pb_ephemeris* e, FlowWorkspace* w,
                            double const* times_host, int n_times,
                            double* table, cudaStream_t stream) {
  double lo = std::numeric_limits<double>::infinity();
  double hi = -lo;
  for (int i = 0; i < n_times; ++i) {
    if (times_host[i] != times_host[i]) {
      return Fail(PB_INVALID_ARGUMENT, "NaN time");
    }
    lo = std::min(lo, times_host[i]);
    hi = std::max(hi, times_host[i]);
  }
  // CHECK_LE(t_min, time); CHECK_GE(t_max, time) in
  // ContinuousTrajectory::EvaluatePositionLocked.
  if (lo < EphemerisTMin(e) || hi > EphemerisTMax(e)) {
    return Fail(PB_INVALID_ARGUMENT,
                "time outside [t_min, t_max] of the ephemeris: Prolong first");
  }
  PB_RETURN_IF_ERROR(w->stage_times.Upload(times_host, n_times, stream));
  EphemerisView const view = e->View();
  long long const threads = static_cast<long long>(n_times) * view.n_bodies;
  int const block = 128;
  StageTableKernel<<<static_cast<unsigned>((threads + block - 1) / block),
                     block, 0, stream>>>(view, w->stage_times.data, n_times,
                                         table);
  PB_LAUNCHED();
  return PB_OK;
}

// Defined in pb_api_nbody.cu.
bool MakeMultistepMethod(int32_t method, MultistepMethod& m);

// integrators/symplectic_runge_kutta_nyström_integrator_body.hpp:198-208: the
// nodes c_i are the double-double partial sums of the a_i.
void ComputeNodes(SrknTableau const& tableau, double* c) {
  DP c_i{0.0, 0.0};
  for (int i = 0; i < tableau.stages; ++i) {
    c[i] = c_i.value;
    c_i = Add(c_i, DP{tableau.a[i], 0.0});
  }
}

// The "hot" oblate body of the flow kernels -- the body of degree <=
// kHotDegree whose harmonics of degree >= 3 the most probes of the batch need
// (the Earth for LEO probes); -1 if none -- and the "warm" one: another body
// whose harmonics up to degree kWarmDegree the most probes need (the damped J2
// of the Sun); -1 if none.  Every probe votes (HotBodyVoteKernel), so that an
// unrepresentative first probe does not send the batch down the slow paths.
// The choice only affects speed: every path computes the same bits.
pb_status_t ChooseHotBodies(pb_ephemeris const* e, double const t,
                            long long const n,
                            double const* const state_device,
                            cudaStream_t const stream, int* const hot_out,
                            int* const warm_out) {
  *hot_out = -1;
  *warm_out = -1;
  int const n_oblate = std::min(e->model.n_oblate, 32);
  if (n_oblate == 0 || n <= 0) {
    return PB_OK;
  }
  // Positions of the oblate bodies at t, from the host copy of the segments.
  std::vector<double> positions(3 * n_oblate, 0.0);
  for (int b = 0; b < n_oblate; ++b) {
    HostTrajectory const& tr = e->trajectories[b];
    if (tr.t_max.empty()) {
      // No ephemeris yet for this body: nothing to choose (the flow itself
      // will fail its own precondition).
      return PB_OK;
    }
    std::size_t segment =
        std::lower_bound(tr.t_max.begin(), tr.t_max.end(), t) -
        tr.t_max.begin();
    segment = std::min(segment, tr.t_max.size() - 1);
    double const x = t - tr.origin[segment];
    double const* const c = &tr.coefficients[segment * 3 * kSegmentCoefficients];
    for (int k = 0; k < 3; ++k) {
      positions[3 * b + k] = EstrinEvaluate<1>(c + k * kSegmentCoefficients,
                                               tr.degree[segment], x);
    }
  }
  DeviceBuffer<double> positions_device;
  DeviceBuffer<int32_t> votes_device;
  PB_RETURN_IF_ERROR(positions_device.Upload(positions, stream));
  PB_RETURN_IF_ERROR(votes_device.Reserve(n_oblate * kVoteBins));
  PB_CUDA(cudaMemsetAsync(votes_device.data, 0,
                          n_oblate * kVoteBins * sizeof(int32_t), stream));
  HotBodyVoteKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      e->View(), positions_device.data, n, state_device, votes_device.data);
  PB_LAUNCHED();
  std::vector<int32_t> votes(n_oblate * kVoteBins);
  PB_CUDA(cudaMemcpyAsync(votes.data(), votes_device.data,
                          votes.size() * sizeof(int32_t),
                          cudaMemcpyDeviceToHost, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  auto const count = [&](int const b, int const first, int const last) {
    long long total = 0;
    for (int d = first; d <= last; ++d) {
      total += votes[b * kVoteBins + d];
    }
    return total;
  };
  long long best = 0;
  for (int b = 0; b < n_oblate; ++b) {
    long long const deep = count(b, 3, kVoteBins - 1);
    if (e->model.bodies[b].degree <= kHotDegree && deep > best) {
      best = deep;
      *hot_out = b;
    }
  }
  best = 0;
  for (int b = 0; b < n_oblate; ++b) {
    long long const shallow = count(b, 2, kWarmDegree);
    if (b != *hot_out && shallow > best) {
      best = shallow;
      *warm_out = b;
    }
  }
  return PB_OK;
}

// The constant bank of this translation unit's flow kernels (c_bodies, c_hot),
// and that of the kernels compiled with contraction (pb_contracted.cu).
ConstantBank& FlowConstantBank() {
  static ConstantBank bank;
  return bank;
}
ConstantBank& ContractedConstantBank() {
  static ConstantBank bank;
  return bank;
}

// Makes the constant bank hold the per-body scalars of `e` and the tables of
// its hot and warm bodies, choosing them if they have not been chosen yet (or
// if the previous flow reported that the choice was poor) by a vote of the
// probes in `state_device` (a small kernel and a synchronisation, once per
// choice).  The lease must outlive every kernel launched under it.
pb_status_t AcquireFlowConstants(pb_ephemeris* e, double const t,
                                 long long const n,
                                 double const* const state_device,
                                 double const* const probe_host,
                                 int32_t const arithmetic,
                                 cudaStream_t stream,
                                 ConstantBankLease& lease, int* hot_out,
                                 int* warm_out) {
  int const n_bodies = static_cast<int>(e->model.bodies.size());
  if (n_bodies > kMaxConstantBodies) {
    return Fail(PB_INVALID_ARGUMENT, "too many bodies for the flow kernel");
  }
  int hot;
  int warm;
  {
    std::lock_guard<std::mutex> const lock(e->hot_mutex);
    if (!e->hot_body_valid) {
      PB_RETURN_IF_ERROR(ChooseHotBodies(e, t, n, state_device, stream,
                                         &e->hot_body, &e->warm_body));
      e->hot_body_valid = n > 0;
    }
    hot = e->hot_body;
    warm = e->warm_body;
  }
  *hot_out = hot;
  *warm_out = warm;
  bool const contracted = arithmetic == PB_ARITHMETIC_CONTRACTED;
  ConstantBank& bank =
      contracted ? ContractedConstantBank() : FlowConstantBank();
  // Variant 0 is "no hot body, no warm body".
  pb_status_t const status = bank.Acquire(
      e->serial, (hot + 1) + (kMaxConstantBodies + 1) * (warm + 1),
      [&]() -> pb_status_t {
        static BodyScalars scalars;
        static HotBody hot_body;
        for (int b = 0; b < n_bodies; ++b) {
          DeviceBody const& body = e->model.bodies[b];
          scalars.mu[b] = body.mu;
          scalars.collision_radius[b] = body.collision_radius;
          scalars.outer_threshold_degree_2[b] =
              body.is_oblate && body.n_damping > 2
                  ? e->model.dampings[body.damping_offset + 2].outer_threshold
                  : -std::numeric_limits<double>::infinity();
        }
        std::memset(&hot_body, 0, sizeof(hot_body));
        hot_body.body = hot;
        hot_body.warm_body = warm;
        for (int k = 0; k < kHotSize; ++k) {
          hot_body.normalization[k] = pb_legendre_normalization_factor[k];
        }
        if (warm >= 0) {
          DeviceBody const& body = e->model.bodies[warm];
          hot_body.warm_data = body;
          hot_body.warm_n_damping = std::min(body.n_damping, kWarmDampings);
          for (int d = 0; d < hot_body.warm_n_damping; ++d) {
            hot_body.warm_degree_damping[d] =
                e->model.dampings[body.damping_offset + d];
          }
          int const size =
              std::min((body.degree + 1) * (body.degree + 2) / 2, kWarmSize);
          for (int k = 0; k < size; ++k) {
            hot_body.warm_cos[k] = e->model.cos[body.coefficient_offset + k];
            hot_body.warm_sin[k] = e->model.sin[body.coefficient_offset + k];
          }
        }
        if (hot >= 0) {
          DeviceBody const& body = e->model.bodies[hot];
          int const size = (body.degree + 1) * (body.degree + 2) / 2;
          hot_body.n_damping = body.n_damping;
          for (int k = 0; k < size; ++k) {
            hot_body.cos[k] = e->model.cos[body.coefficient_offset + k];
            hot_body.sin[k] = e->model.sin[body.coefficient_offset + k];
          }
          for (int d = 0; d < body.n_damping; ++d) {
            hot_body.degree_damping[d] =
                e->model.dampings[body.damping_offset + d];
          }
          hot_body.sectoral = body.sectoral;
          hot_body.hot_data = body;
        }
        // No kernel reads the bank now (ConstantBank::Acquire).
        if (contracted) {
          // The contracted kernels read the normalisation factor as 1 (see
          // HotTables): fold it into the coefficients here.
          for (int k = 0; k < kHotSize; ++k) {
            hot_body.cos[k] *= hot_body.normalization[k];
            hot_body.sin[k] *= hot_body.normalization[k];
          }
          for (int k = 0; k < kWarmSize; ++k) {
            hot_body.warm_cos[k] *= hot_body.normalization[k];
            hot_body.warm_sin[k] *= hot_body.normalization[k];
          }
          PB_CUDA(pb_contracted_api::UploadConstants(
              &scalars, sizeof(scalars), &hot_body, sizeof(hot_body), stream));
          return PB_OK;
        }
        PB_CUDA(cudaMemcpyToSymbolAsync(c_bodies, &scalars, sizeof(scalars), 0,
                                        cudaMemcpyHostToDevice, stream));
        PB_CUDA(cudaMemcpyToSymbolAsync(c_hot, &hot_body, sizeof(hot_body), 0,
                                        cudaMemcpyHostToDevice, stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        return PB_OK;
      });
  if (status == PB_OK) {
    lease.Adopt(&bank);
  }
  return status;
}

// What a flow leaves for pb_ephemeris_last_flow_*: its kernel time, and the
// code paths it took.  A flow that met the slow generic harmonics makes the
// next one choose its hot body afresh.
void RecordFlowStatistics(pb_ephemeris* e, double const kernel_ms,
                          std::int64_t const launches, int const hot,
                          int32_t const path_flags) {
  {
    std::lock_guard<std::mutex> const lock(e->statistics_mutex);
    e->last_flow_kernel_ms = kernel_ms;
    e->last_flow_kernel_launches = launches;
    e->last_flow_hot_body = hot >= 0 ? e->model.order[hot] : -1;
    e->last_flow_path_flags = path_flags;
  }
  if ((path_flags & kPathGenericHarmonics) != 0) {
    std::lock_guard<std::mutex> const lock(e->hot_mutex);
    e->hot_body_valid = false;
  }
}

// Slicing of a flow launch across helper streams (PB_FLOW_GROUPS = 1 disables
// it) and steps per sliced launch.
int FlowGroups() {
  static int const value = [] {
    char const* const text = std::getenv("PB_FLOW_GROUPS");
    int const v = text != nullptr ? std::atoi(text) : 4;
    return std::max(1, std::min(v, 16));
  }();
  return value;
}
// The same for the massless multistep flow, whose launches reload the history
// ring: 1 000 steps of 10^5 probes run at 2.04 10^9 probe-steps/s with 10 steps
// per launch, 1.93 10^9 with 4.
int MultistepSubSteps() {
  static int const value = [] {
    char const* const text = std::getenv("PB_MULTISTEP_SUB_STEPS");
    int const v = text != nullptr ? std::atoi(text) : 10;
    return std::max(1, v);
  }();
  return value;
}
int FlowSubSteps() {
  static int const value = [] {
    char const* const text = std::getenv("PB_FLOW_SUB_STEPS");
    // Measured on 10^5 probes x 100 SRKN14A steps, 4 groups: 43.4, 43.3,
    // 43.4, 43.5, 44.0 and 45.8 ms for 2, 3, 4, 5, 10 and 20 steps per launch.
    int const v = text != nullptr ? std::atoi(text) : 4;
    return std::max(1, v);
  }();
  return value;
}

// Resident blocks per SM the flow kernel is compiled for (register budget
// 65536 / (128 * blocks)); PB_SRKN_MIN_BLOCKS overrides it for experiments.
int SrknMinBlocks() {
  static int const value = [] {
    char const* const text = std::getenv("PB_SRKN_MIN_BLOCKS");
    int const v = text != nullptr ? std::atoi(text) : 3;
    return v == 2 || v == 4 ? v : 3;
  }();
  return value;
}

// The warp-specialised kernel ahead of the general one
// (pb_kernels_massless_specialized.cuh): opt-in with PB_SRKN_SPECIALIZED=1; it
// is bit-identical and measured slower (57.3 against 45.6 ms per 100 steps of
// 10^5 probes, profiles/r02w_*).
bool SrknSpecialized() {
  static bool const value = [] {
    char const* const text = std::getenv("PB_SRKN_SPECIALIZED");
    return text != nullptr && std::atoi(text) != 0;
  }();
  return value;
}

// For a launch that starts after `steps_before` steps of a flow sampled every
// `every` steps (sample k, k >= 1, is taken after step k * every and goes to
// slot k - 1): the step of the launch, counted from 1, after which its first
// sample is due, and the slot of that sample.
long long FirstSampleStep(long long const steps_before, long long const every) {
  return every > 0 ? every - steps_before % every : 0;
}
long long FirstSampleSlot(long long const steps_before, long long const every) {
  return every > 0 ? steps_before / every : 0;
}

// Shape of the flow kernel: 0 = 128 threads x SrknMinBlocks() blocks per SM;
// 3 = 384 threads x 1 block (168 registers, 12 warps per SM); 5 = 256 threads x
// 2 blocks (128 registers, 16 warps per SM).  Measured on 10^5 probes: 56.2,
// 50.7 and 48.4 ms per 100 steps (a block barrier per stage did not help; 18,
// 20 and 24 warps per SM at 96 and 80 registers: 69, 66 and 74 ms -- 18 warps
// cannot have more than 96 registers, one scheduler holds 5 of them), so a launch
// that fills every SM takes shape 5, a smaller one shape 3 or 0 (more, smaller
// blocks).  PB_SRKN_VARIANT overrides the choice (tools/flow_variants.sh).
int SrknVariant(long long const n_probes, int const sm_count) {
  static int const forced = [] {
    char const* const text = std::getenv("PB_SRKN_VARIANT");
    int const v = text != nullptr ? std::atoi(text) : -1;
    return v >= 0 && v <= 5 ? v : -1;
  }();
  if (forced >= 0) {
    return forced;
  }
  return n_probes >= 512LL * sm_count ? 5 : n_probes >= 384LL * sm_count ? 3 : 0;
}

// Shape of the massless multistep flow kernel: 128 threads x 3 blocks per SM
// (168 registers).  256 x 2 at 128 registers, the better shape of the SRKN
// kernel, is slower here (10^5 probes, Quinlan1999Order8A: 1.72 against
// 2.04 10^9 probe-steps/s): the history sums add to the live state.
// PB_MULTISTEP_WIDE=1 selects it for experiments.
bool MultistepWide(long long const n_probes, int const sm_count) {
  static int const forced = [] {
    char const* const text = std::getenv("PB_MULTISTEP_WIDE");
    return text != nullptr ? std::atoi(text) : 0;
  }();
  return forced != 0 && n_probes >= 512LL * sm_count;
}

bool MakeSrknTableau(int32_t method, SrknTableau& t) {
  std::memset(&t, 0, sizeof(t));
  if (method == PB_BLANES_MOAN_2002_SRKN_14A) {
    // integrators/methods.hpp:379-419 (ABA, 15 stages, 14 evaluations).
    static double const a[15] = {
        0.037859319840611600,   0.10263563310243500,  -0.025867888266558700,
        0.31424140307144700,    -0.13014445951741500, 0.10641770036954300,
        -0.0087942431285105800, 0.2073050690568954,   -0.0087942431285105800,
        0.10641770036954300,    -0.13014445951741500, 0.31424140307144700,
        -0.025867888266558700,  0.10263563310243500,  0.037859319840611600};
    static double const b[15] = {
        0.0,
        0.091719152624461650,  0.18398317000500600,   -0.056534365832888270,
        0.0049146887747128540, 0.14376112716835800,   0.32856769374680400,
        -0.19641146648645423,  -0.19641146648645423,  0.32856769374680400,
        0.14376112716835800,   0.0049146887747128540, -0.056534365832888270,
        0.18398317000500600,   0.091719152624461650};
    t.stages = 15;
    t.first_stage = 1;
    std::memcpy(t.a, a, sizeof(a));
    std::memcpy(t.b, b, sizeof(b));
    return true;
  }
  if (method == PB_MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL) {
    // integrators/methods.hpp:922-949 (BA, 6 stages).
    static double const a[6] = {0.339839625839110000,  -0.088601336903027329,
                                0.5858564768259621188, -0.603039356536491888,
                                0.3235807965546976394, 0.4423637942197494587};
    static double const b[6] = {0.1193900292875672758,  0.6989273703824752308,
                                -0.1713123582716007754, 0.4012695022513534480,
                                0.0107050818482359840,  -0.0589796254980311632};
    t.stages = 6;
    t.first_stage = 0;
    std::memcpy(t.a, a, sizeof(a));
    std::memcpy(t.b, b, sizeof(b));
    return true;
  }
  return false;
}

}  // namespace pb

extern "C" {

const char* pb_last_error(void) { return LastError().c_str(); }

int64_t pb_kernel_launch_count(void) { return LaunchCount().load(); }

pb_status_t pb_ephemeris_create(int32_t n_bodies, const pb_body_t* bodies,
                                double geopotential_tolerance,
                                int32_t cuda_device,
                                pb_ephemeris_t** ephemeris) {
  PB_API_RANGE();
  if (ephemeris == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null output");
  }
  *ephemeris = nullptr;
  int device_count = 0;
  cudaError_t const count_error = cudaGetDeviceCount(&device_count);
  if (count_error != cudaSuccess || device_count == 0) {
    return Fail(PB_FAILED_PRECONDITION,
                std::string("no CUDA device; principia_b200 has no CPU "
                            "fallback: ") +
                    cudaGetErrorString(count_error));
  }
  if (cuda_device < 0 || cuda_device >= device_count) {
    return Fail(PB_INVALID_ARGUMENT, "bad CUDA device ordinal");
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  auto* e = new pb_ephemeris;
  e->device = cuda_device;
  pb_status_t status =
      BuildHostModel(n_bodies, bodies, geopotential_tolerance, e->model);
  if (status != PB_OK) {
    delete e;
    return Fail(status, "invalid body description");
  }
  e->trajectories.resize(n_bodies);
  cudaStream_t const stream = nullptr;
  std::vector<double> normalization(
      pb_legendre_normalization_factor,
      pb_legendre_normalization_factor +
          kGeopotentialSize * (kGeopotentialSize + 1) / 2);
  // One dummy element keeps the pointers valid for systems without oblate
  // bodies.
  std::vector<Damping> dampings = e->model.dampings;
  std::vector<double> cos = e->model.cos, sin = e->model.sin;
  if (dampings.empty()) {
    dampings.push_back(DefaultDamping());
    cos.push_back(0.0);
    sin.push_back(0.0);
  }
  auto cuda_ok = [](cudaError_t const error) {
    return error == cudaSuccess ? PB_OK
                                : Fail(PB_INTERNAL, cudaGetErrorString(error));
  };
  status = e->bodies.Upload(e->model.bodies, stream);
  if (status == PB_OK) status = e->dampings.Upload(dampings, stream);
  if (status == PB_OK) status = e->cos.Upload(cos, stream);
  if (status == PB_OK) status = e->sin.Upload(sin, stream);
  if (status == PB_OK) status = e->normalization.Upload(normalization, stream);
  if (status == PB_OK) {
    std::vector<int32_t> const internal_of_caller(
        e->model.internal_of_caller.begin(), e->model.internal_of_caller.end());
    status = e->internal_of_caller.Upload(internal_of_caller, stream);
  }
  if (status == PB_OK) {
    status = cuda_ok(cudaMalloc(&e->stop_word_device, sizeof(int32_t)));
  }
  if (status == PB_OK) {
    status = cuda_ok(
        cudaMemsetAsync(e->stop_word_device, 0, sizeof(int32_t), stream));
  }
  if (status == PB_OK) {
    status = cuda_ok(cudaStreamCreateWithFlags(&e->control_stream,
                                               cudaStreamNonBlocking));
  }
  if (status == PB_OK) status = cuda_ok(cudaStreamSynchronize(stream));
  if (status != PB_OK) {
    delete e;
    return status;
  }
  *ephemeris = e;
  return PB_OK;
}

void pb_ephemeris_destroy(pb_ephemeris_t* ephemeris) {
  PB_API_RANGE();
  if (ephemeris == nullptr) {
    return;
  }
  if (ephemeris->dependants.load() != 0) {
    // Instances and ensembles hold a pointer to their ephemeris: destroying it
    // under them would leave them dangling.  The handle leaks instead, and the
    // error is recorded.
    Fail(PB_FAILED_PRECONDITION,
         "pb_ephemeris_destroy: instances or ensembles of this ephemeris are "
         "still alive; destroy them first");
    return;
  }
  cudaSetDevice(ephemeris->device);
  cudaDeviceSynchronize();
  delete ephemeris;
}

pb_status_t pb_ephemeris_request_stop(pb_ephemeris_t* e) {
  PB_API_RANGE();
  if (e == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null ephemeris");
  }
  // No lock: this is called while a flow holds them.
  e->stop_requested.store(1);
  PB_CUDA(cudaSetDevice(e->device));
  static int32_t const one = 1;
  PB_CUDA(cudaMemcpyAsync(e->stop_word_device, &one, sizeof(int32_t),
                          cudaMemcpyHostToDevice, e->control_stream));
  PB_CUDA(cudaStreamSynchronize(e->control_stream));
  return PB_OK;
}

pb_status_t pb_ephemeris_clear_stop(pb_ephemeris_t* e) {
  PB_API_RANGE();
  if (e == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null ephemeris");
  }
  e->stop_requested.store(0);
  PB_CUDA(cudaSetDevice(e->device));
  PB_CUDA(cudaMemsetAsync(e->stop_word_device, 0, sizeof(int32_t),
                          e->control_stream));
  PB_CUDA(cudaStreamSynchronize(e->control_stream));
  return PB_OK;
}

pb_status_t pb_ephemeris_internal_order(const pb_ephemeris_t* e, int32_t* order,
                                        int32_t* n_oblate) {
  if (e == nullptr || order == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  std::copy(e->model.order.begin(), e->model.order.end(), order);
  if (n_oblate != nullptr) {
    *n_oblate = e->model.n_oblate;
  }
  return PB_OK;
}

pb_status_t pb_geopotential_damping(const pb_ephemeris_t* e, int32_t body,
                                    int32_t* n_degrees,
                                    double* inner_thresholds,
                                    double* sectoral_inner_threshold) {
  if (e == nullptr || body < 0 ||
      body >= static_cast<int32_t>(e->model.bodies.size())) {
    return Fail(PB_INVALID_ARGUMENT, "bad body");
  }
  DeviceBody const& b = e->model.bodies[e->model.internal_of_caller[body]];
  if (!b.is_oblate) {
    return Fail(PB_INVALID_ARGUMENT, "body is not oblate");
  }
  *n_degrees = b.n_damping;
  for (int i = 0; i < b.n_damping; ++i) {
    inner_thresholds[i] =
        e->model.dampings[b.damping_offset + i].inner_threshold;
  }
  *sectoral_inner_threshold = b.sectoral.inner_threshold;
  return PB_OK;
}

pb_status_t pb_ephemeris_append_segments(pb_ephemeris_t* e, int32_t body,
                                         int32_t n_segments, double t_min,
                                         const double* t_max,
                                         const double* origin,
                                         const int32_t* degree,
                                         const double* coefficients) {
  PB_API_RANGE();
  if (e == nullptr || body < 0 ||
      body >= static_cast<int32_t>(e->model.bodies.size()) || n_segments < 0 ||
      (n_segments > 0 && (t_max == nullptr || origin == nullptr ||
                          degree == nullptr || coefficients == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  // Validate before mutating anything: a rejected call leaves the trajectory
  // as it was.
  for (int s = 0; s < n_segments; ++s) {
    if (degree[s] < 0 || degree[s] >= kSegmentCoefficients) {
      return Fail(PB_INVALID_ARGUMENT, "segment degree out of range");
    }
  }
  // Waits for the flows and queries in progress.
  std::unique_lock<std::shared_mutex> const lock(e->segments_mutex);
  HostTrajectory& tr = e->trajectories[e->model.internal_of_caller[body]];
  double previous = tr.t_max.empty() ? t_min : tr.t_max.back();
  for (int s = 0; s < n_segments; ++s) {
    if (!(t_max[s] > previous)) {
      return Fail(PB_INVALID_ARGUMENT, "segments must be appended in order");
    }
    previous = t_max[s];
  }
  for (int s = 0; s < n_segments; ++s) {
    if (!tr.has_t_min) {
      tr.has_t_min = true;
      tr.t_min = t_min;
    }
    tr.t_max.push_back(t_max[s]);
    tr.origin.push_back(origin[s]);
    tr.degree.push_back(degree[s]);
    // Host API [18][3] -> device layout [xyz][18]; the slots above the degree
    // are padding: -0.0, the neutral element that EvaluateSegmentPadded relies
    // on (pb_accel.cuh).
    std::size_t const base = tr.coefficients.size();
    tr.coefficients.resize(base + 3 * kSegmentCoefficients, -0.0);
    double const* const in = coefficients + 3 * kSegmentCoefficients * s;
    for (int k = 0; k <= degree[s]; ++k) {
      for (int c = 0; c < 3; ++c) {
        tr.coefficients[base + c * kSegmentCoefficients + k] = in[3 * k + c];
      }
    }
  }
  e->segments_dirty = true;
  return PB_OK;
}

double pb_ephemeris_t_min(const pb_ephemeris_t* e) {
  std::shared_lock<std::shared_mutex> const lock(e->segments_mutex);
  return EphemerisTMin(e);
}
double pb_ephemeris_t_max(const pb_ephemeris_t* e) {
  std::shared_lock<std::shared_mutex> const lock(e->segments_mutex);
  return EphemerisTMax(e);
}

pb_status_t pb_ephemeris_evaluate_all_positions(pb_ephemeris_t* e, double t,
                                                double* positions) {
  PB_API_RANGE();
  if (e == nullptr || positions == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, false, nullptr));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  cudaStream_t const stream = w.stream();
  EphemerisView const view = e->View();
  PB_RETURN_IF_ERROR(w->stage_table.Reserve(view.stage_stride));
  PB_RETURN_IF_ERROR(
      BuildStageTable(e, w.get(), &t, 1, w->stage_table.data, stream));
  std::vector<double> entry(view.stage_stride);
  PB_CUDA(cudaMemcpyAsync(entry.data(), w->stage_table.data,
                          entry.size() * sizeof(double), cudaMemcpyDeviceToHost,
                          stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  for (int b = 0; b < view.n_bodies; ++b) {
    int const caller = e->model.order[b];
    for (int c = 0; c < 3; ++c) {
      positions[3 * caller + c] = entry[3 * b + c];
    }
  }
  return PB_OK;
}

}  // extern "C"

namespace {

// ...ByAllMassiveBodiesOnMasslessBodies on device arrays, asynchronous on
// `stream`; the caller holds the workspace and the segments.
pb_status_t MasslessAccelerations(pb_ephemeris* e, FlowWorkspace* w, double t,
                                  int64_t n, const double* q, double* a,
                                  int32_t* status, cudaStream_t stream) {
  EphemerisView const view = e->View();
  PB_RETURN_IF_ERROR(w->stage_table.Reserve(view.stage_stride));
  PB_RETURN_IF_ERROR(BuildStageTable(e, w, &t, 1, w->stage_table.data, stream));
  if (n > 0) {
    MasslessAccelerationKernel<true>
        <<<static_cast<unsigned>((n + kFlowThreads - 1) / kFlowThreads),
           kFlowThreads, 0, stream>>>(view, w->stage_table.data, n, q, a,
                                      status);
    PB_LAUNCHED();
  }
  return PB_OK;
}

}  // namespace

extern "C" {

pb_status_t pb_compute_gravitational_acceleration_on_massless_bodies_device(
    pb_ephemeris_t* e, double t, int64_t n, const double* q, double* a,
    int32_t* status_device, void* stream_pointer) {
  PB_API_RANGE();
  if (e == nullptr || n < 0 || (n > 0 && (q == nullptr || a == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = static_cast<cudaStream_t>(stream_pointer);
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, true, stream));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  // Asynchronous: the workspace is handed back with an event that the next
  // user waits for.
  return MasslessAccelerations(
      e, w.get(), t, n, q, a,
      status_device != nullptr ? status_device : w->status_word.data, stream);
}

pb_status_t pb_compute_gravitational_acceleration_on_massless_bodies(
    pb_ephemeris_t* e, double t, int64_t n, const double* q, double* a,
    pb_status_t* status) {
  PB_API_RANGE();
  if (e == nullptr || n < 0 || (n > 0 && (q == nullptr || a == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, false, nullptr));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  cudaStream_t const stream = w.stream();
  PB_RETURN_IF_ERROR(w->scratch_a.Upload(q, 3 * n, stream));
  PB_RETURN_IF_ERROR(w->scratch_b.Reserve(3 * n));
  PB_CUDA(cudaMemsetAsync(w->status_word.data, 0, sizeof(int32_t), stream));
  PB_RETURN_IF_ERROR(MasslessAccelerations(e, w.get(), t, n, w->scratch_a.data,
                                           w->scratch_b.data,
                                           w->status_word.data, stream));
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(a, w->scratch_b.data, 3 * n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
  }
  int32_t word = 0;
  PB_CUDA(cudaMemcpyAsync(&word, w->status_word.data, sizeof(int32_t),
                          cudaMemcpyDeviceToHost, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  if (status != nullptr) {
    *status = word;
  }
  return PB_OK;
}

pb_status_t pb_compute_gravitational_potential(pb_ephemeris_t* e, double t,
                                               int64_t n, const double* q,
                                               double* potential) {
  PB_API_RANGE();
  if (e == nullptr || n < 0 ||
      (n > 0 && (q == nullptr || potential == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, false, nullptr));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  cudaStream_t const stream = w.stream();
  EphemerisView const view = e->View();
  PB_RETURN_IF_ERROR(w->stage_table.Reserve(view.stage_stride));
  PB_RETURN_IF_ERROR(
      BuildStageTable(e, w.get(), &t, 1, w->stage_table.data, stream));
  PB_RETURN_IF_ERROR(w->scratch_a.Upload(q, 3 * n, stream));
  PB_RETURN_IF_ERROR(w->scratch_b.Reserve(n));
  if (n > 0) {
    PotentialKernel<<<static_cast<unsigned>((n + kFlowThreads - 1) /
                                            kFlowThreads),
                      kFlowThreads, 0, stream>>>(view, w->stage_table.data, n,
                                                 w->scratch_a.data,
                                                 w->scratch_b.data);
    PB_LAUNCHED();
    PB_CUDA(cudaMemcpyAsync(potential, w->scratch_b.data, n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  return PB_OK;
}

pb_status_t pb_compute_gravitational_acceleration_on_massive_bodies(
    pb_ephemeris_t* e, double t, const double* positions,
    double* accelerations) {
  PB_API_RANGE();
  if (e == nullptr || accelerations == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  WorkspaceLease w;
  PB_RETURN_IF_ERROR(w.Acquire(e, false, nullptr));
  SegmentsReadLock segments;
  PB_RETURN_IF_ERROR(segments.Acquire(e));
  cudaStream_t const stream = w.stream();
  EphemerisView const view = e->View();
  int const n = view.n_bodies;
  PB_RETURN_IF_ERROR(w->stage_table.Reserve(view.stage_stride));
  if (positions == nullptr) {
    PB_RETURN_IF_ERROR(
        BuildStageTable(e, w.get(), &t, 1, w->stage_table.data, stream));
  } else {
    std::vector<double> internal(3 * n);
    for (int b = 0; b < n; ++b) {
      int const caller = e->model.order[b];
      for (int c = 0; c < 3; ++c) {
        internal[3 * b + c] = positions[3 * caller + c];
      }
    }
    PB_CUDA(cudaMemcpyAsync(w->stage_table.data, internal.data(),
                            internal.size() * sizeof(double),
                            cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    FramesKernel<<<(n + 63) / 64, 64, 0, stream>>>(view, t,
                                                  w->stage_table.data);
    PB_LAUNCHED();
  }
  PB_RETURN_IF_ERROR(w->scratch_b.Reserve(3 * n));
  MassiveAccelerationKernel<<<(n + 31) / 32, 32, 0, stream>>>(
      view, w->stage_table.data, w->scratch_b.data);
  PB_LAUNCHED();
  std::vector<double> internal(3 * n);
  PB_CUDA(cudaMemcpyAsync(internal.data(), w->scratch_b.data,
                          internal.size() * sizeof(double),
                          cudaMemcpyDeviceToHost, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  for (int b = 0; b < n; ++b) {
    int const caller = e->model.order[b];
    for (int c = 0; c < 3; ++c) {
      accelerations[3 * caller + c] = internal[3 * b + c];
    }
  }
  return PB_OK;
}

}  // extern "C"

namespace {

// What one fixed-step flow call holds for its duration: a workspace, the
// segments, the constant bank; and what it accumulates.
struct FlowContext {
  pb_ephemeris* e = nullptr;
  // Declared in the order in which they must be released LAST to FIRST: the
  // constant bank and the segments go before the workspace is handed back.
  WorkspaceLease workspace;
  SegmentsReadLock segments;
  ConstantBankLease constants;
  cudaStream_t stream = nullptr;
  int hot = -1;
  int warm = -1;
  int32_t arithmetic = PB_ARITHMETIC_EXACT;
  // RETURN_IF_STOPPED: set when the flow stopped at a chunk boundary because
  // pb_ephemeris_request_stop was called.
  bool cancelled = false;
  bool path_checked = false;
  double kernel_ms = 0;
  std::int64_t launches = 0;
  std::size_t phases = 0;

  FlowWorkspace* w() const { return workspace.get(); }

  // `probe_host`: the first probe's position if the caller has it on the host
  // (3 values at stride n), else null and `state_device` is read.
  pb_status_t Begin(pb_ephemeris* const ephemeris, bool const user_stream_given,
                    cudaStream_t const user_stream) {
    e = ephemeris;
    PB_CUDA(cudaSetDevice(e->device));
    PB_RETURN_IF_ERROR(workspace.Acquire(e, user_stream_given, user_stream));
    stream = workspace.stream();
    PB_RETURN_IF_ERROR(segments.Acquire(e));
    // [0]: right-hand-side status, [1]: path flags.
    PB_CUDA(cudaMemsetAsync(w()->status_word.data, 0, 2 * sizeof(int32_t),
                            stream));
    return PB_OK;
  }

  pb_status_t AcquireConstants(double const t, long long const n,
                               double const* const state_device,
                               double const* const probe_host,
                               int32_t const flow_arithmetic) {
    if (constants.held()) {
      return PB_OK;
    }
    if (flow_arithmetic != PB_ARITHMETIC_EXACT &&
        flow_arithmetic != PB_ARITHMETIC_CONTRACTED) {
      return Fail(PB_INVALID_ARGUMENT, "unknown arithmetic");
    }
    arithmetic = flow_arithmetic;
    return AcquireFlowConstants(e, t, n, state_device, probe_host, arithmetic,
                                stream, constants, &hot, &warm);
  }
  bool contracted() const { return arithmetic == PB_ARITHMETIC_CONTRACTED; }

  EphemerisView View() const { return e->View(); }

  // Once per call: would some probe take the slow generic harmonics?  `stage`
  // is a stage-table entry close to the initial time of the flow.
  pb_status_t CheckPath(double const* const stage, long long const n,
                        double const* const state) {
    if (path_checked || n == 0) {
      return PB_OK;
    }
    path_checked = true;
    FlowPathKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        e->View(), stage, hot, warm, n, state, w()->status_word.data + 1);
    PB_LAUNCHED();
    return PB_OK;
  }

  // Sums the device time of the kernel phases recorded since the last call.
  pb_status_t CollectKernelTime() {
    for (std::size_t i = 0; i < phases; ++i) {
      float elapsed = 0;
      PB_CUDA(cudaEventElapsedTime(&elapsed, w()->timing_events[2 * i],
                                   w()->timing_events[2 * i + 1]));
      kernel_ms += elapsed;
    }
    phases = 0;
    return PB_OK;
  }

  // Reads the status words (the stream is synchronised) and publishes the
  // statistics of the call.  *rhs_status: 0 or PB_OUT_OF_RANGE.
  pb_status_t Finish(pb_status_t* const rhs_status) {
    PB_CUDA(cudaMemcpyAsync(w()->pinned_status.data, w()->status_word.data,
                            2 * sizeof(int32_t), cudaMemcpyDeviceToHost,
                            stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    PB_RETURN_IF_ERROR(CollectKernelTime());
    RecordFlowStatistics(e, kernel_ms, launches, hot,
                         w()->pinned_status.data[1]);
    if (rhs_status != nullptr) {
      // The only error the right-hand side reports is the collision
      // (ephemeris_body.hpp:383-384).
      *rhs_status = w()->pinned_status.data[0] != 0 ? PB_OUT_OF_RANGE : PB_OK;
    }
    return PB_OK;
  }

  pb_status_t EnsureHelpers(int const groups) {
    FlowWorkspace* const ws = w();
    while (static_cast<int>(ws->flow_streams.size()) < groups) {
      cudaStream_t helper;
      PB_CUDA(cudaStreamCreateWithFlags(&helper, cudaStreamNonBlocking));
      ws->flow_streams.push_back(helper);
      cudaEvent_t event;
      PB_CUDA(cudaEventCreateWithFlags(&event, cudaEventDisableTiming));
      ws->flow_events.push_back(event);
    }
    while (ws->timing_events.size() < 2 * (phases + 1)) {
      cudaEvent_t event;
      PB_CUDA(cudaEventCreate(&event));
      ws->timing_events.push_back(event);
    }
    return PB_OK;
  }
};

// The steps of an SRKN flow on device state, at most `max_steps` of them (the
// Starter of the multistep integrators runs groups of 16).  The context has
// been begun and holds the constants.  Does not read the status back
// (FlowContext::Finish does).  On return flow->time and *flow->n_steps /
// n_samples are updated; c.cancelled tells whether the flow stopped early.
pb_status_t SrknSteps(FlowContext& c, const pb_fixed_step_flow_t* flow,
                      long long const max_steps) {
  pb_ephemeris* const e = c.e;
  FlowWorkspace* const w = c.w();
  cudaStream_t const stream = c.stream;
  SrknTableau tableau;
  if (!MakeSrknTableau(flow->method, tableau)) {
    return Fail(PB_INVALID_ARGUMENT, "unsupported fixed-step method");
  }
  if (flow->step == 0 || flow->step != flow->step) {
    return Fail(PB_INVALID_ARGUMENT, "CHECK_NE(Time(), step)");
  }
  EphemerisView const view = c.View();
  double nodes[16];
  ComputeNodes(tableau, nodes);
  int const evaluations = tableau.stages - tableau.first_stage;
  double const h = flow->step;
  double const abs_h = h < 0 ? -h : h;
  DP t{flow->time[0], flow->time[1]};
  double const t_final = flow->t_final;
  bool const sampling = flow->samples != nullptr && flow->sample_every > 0;

  // Steps per launch phase: bounds the stage table (evaluations * stride
  // doubles per step), and is the granularity of cancellation.
  int const chunk_steps = 256;
  std::size_t const table_doubles =
      static_cast<std::size_t>(chunk_steps) * evaluations * view.stage_stride;
  PB_RETURN_IF_ERROR(w->stage_table.Reserve(table_doubles));
  PB_RETURN_IF_ERROR(w->pinned_times.Reserve(
      static_cast<std::size_t>(chunk_steps) * evaluations));
  int sm_count = 0;
  PB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount,
                                 e->device));

  long long steps_done = 0;
  long long samples_done = 0;
  while (steps_done < max_steps &&
         abs_h <= std::abs((t_final - t.value) - t.error)) {
    // The pinned staging buffer is reused across chunks and across calls (the
    // startup of a multistep instance calls this every 16 steps): its previous
    // upload must have completed.  This is also where a stop request is
    // honoured: the state is that of the last completed step
    // (RETURN_IF_STOPPED,
    // symplectic_runge_kutta_nyström_integrator_body.hpp:137).
    PB_CUDA(cudaStreamSynchronize(stream));
    if (e->stop_requested.load() != 0) {
      c.cancelled = true;
      break;
    }
    int steps = 0;
    double* const times = w->pinned_times.data;
    while (steps < chunk_steps && steps_done + steps < max_steps &&
           abs_h <= std::abs((t_final - t.value) - t.error)) {
      for (int st = tableau.first_stage; st < tableau.stages; ++st) {
        times[steps * evaluations + (st - tableau.first_stage)] =
            t.value + (t.error + nodes[st] * h);
      }
      Increment(t, h);
      ++steps;
      if (sampling && (steps_done + steps) % flow->sample_every == 0 &&
          samples_done < flow->max_samples) {
        if (flow->sample_times != nullptr) {
          flow->sample_times[samples_done] = t.value;
        }
        ++samples_done;
      }
    }
    PB_RETURN_IF_ERROR(BuildStageTable(e, w, times, steps * evaluations,
                                       w->stage_table.data, stream));
    PB_RETURN_IF_ERROR(c.CheckPath(w->stage_table.data, flow->n, flow->state));
    if (flow->n > 0) {
      // The probes are cut into `groups` slices and the steps into sub-chunks;
      // slice g runs its sub-chunks in order on helper stream g.  Whole
      // launches of 10^5 probes leave 12 % of the SM slots idle in their last
      // wave (782 CTAs on 444 slots); short launches from several streams keep
      // the slots full, and no kernel ever waits on another one.
      int const variant = SrknVariant(flow->n, sm_count);
      int const threads =
          variant == 5 ? 256 : variant >= 2 ? 384 : kFlowThreads;
      long long const blocks = (flow->n + threads - 1) / threads;
      int const groups = static_cast<int>(
          std::min<long long>(FlowGroups(), std::max<long long>(1, blocks / 64)));
      int const sub_steps = groups > 1 ? FlowSubSteps() : steps;
      PB_RETURN_IF_ERROR(c.EnsureHelpers(groups));
      // One entry per block of every launch of this chunk of steps (the same
      // entries, on the same streams, serve the next chunk).
      PB_RETURN_IF_ERROR(w->block_done.Reserve(static_cast<std::size_t>(
          (blocks + groups) * ((steps + sub_steps - 1) / sub_steps))));
      std::size_t done_offset = 0;
      cudaEvent_t const fork = w->timing_events[2 * c.phases];
      cudaEvent_t const join = w->timing_events[2 * c.phases + 1];
      PB_CUDA(cudaEventRecord(fork, stream));
      long long const every = sampling ? flow->sample_every : 0;
      for (int g = 0; g < groups; ++g) {
        cudaStream_t const helper = groups > 1 ? w->flow_streams[g] : stream;
        if (groups > 1) {
          PB_CUDA(cudaStreamWaitEvent(helper, fork, 0));
        }
        long long const i_begin = blocks * g / groups * threads;
        long long const i_end = std::min<long long>(
            flow->n, blocks * (g + 1) / groups * threads);
        unsigned const grid = static_cast<unsigned>(
            (i_end - i_begin + threads - 1) / threads);
        for (int s0 = 0; s0 < steps; s0 += sub_steps) {
          int const n_sub = std::min(sub_steps, steps - s0);
          double const* const table =
              w->stage_table.data +
              static_cast<std::size_t>(s0) * evaluations * view.stage_stride;
          double const* const stage_times =
              w->stage_times.data + static_cast<std::size_t>(s0) * evaluations;
#define PB_LAUNCH_SRKN(kThreads, kMinBlocks)                                  \
  PB_CUDA((SrknStageCacheAttribute<kThreads, kMinBlocks>()));                 \
  (flow->burns != nullptr ? SrknFlowKernel<kThreads, kMinBlocks, true>        \
                          : SrknFlowKernel<kThreads, kMinBlocks, false>)      \
      <<<grid, kThreads, SrknStageCacheBytes(kThreads, view.stage_stride),    \
         helper>>>(                                                           \
          view, tableau, h, table, stage_times, flow->burns, n_sub, flow->n,  \
          i_begin, i_end, flow->state, steps_done + s0, every,                \
          flow->max_samples, flow->samples, w->status_word.data, skip)
          // The warp-specialised kernel first (pb_kernels_massless_specialized
          // .cuh); the general one then flows the blocks it gave up on.
          bool const specialized = variant == 5 && flow->burns == nullptr &&
                                   c.hot >= 0 && SrknSpecialized();
          int32_t* skip = nullptr;
          if (specialized) {
            skip = w->block_done.data + done_offset;
            done_offset += grid;
          }
          if (specialized && !c.contracted()) {
            static cudaError_t const attribute = cudaFuncSetAttribute(
                SrknFlowSpecializedKernel<0>,
                cudaFuncAttributeMaxDynamicSharedMemorySize,
                static_cast<int>(sizeof(SpecializedShared)));
            PB_CUDA(attribute);
            SrknFlowSpecializedKernel<0>
                <<<grid, kSpecializedThreads, sizeof(SpecializedShared),
                   helper>>>(view, tableau, h, table, n_sub, flow->n, i_begin,
                             i_end, flow->state, every,
                             FirstSampleStep(steps_done + s0, every),
                             FirstSampleSlot(steps_done + s0, every),
                             flow->max_samples, flow->samples,
                             w->status_word.data, skip);
            PB_LAUNCHED();
            ++c.launches;
          }
          if (c.contracted()) {
            pb_contracted_api::SrknLaunch launch;
            launch.view = &view;
            launch.tableau = &tableau;
            launch.h = h;
            launch.stage_table = table;
            launch.stage_times = stage_times;
            launch.burns = flow->burns;
            launch.n_steps = n_sub;
            launch.n = flow->n;
            launch.i_begin = i_begin;
            launch.i_end = i_end;
            launch.state = flow->state;
            launch.steps_before = steps_done + s0;
            launch.sample_every = every;
            launch.max_samples = flow->max_samples;
            launch.samples = flow->samples;
            launch.status = w->status_word.data;
            launch.threads = threads;
            launch.grid = grid;
            launch.stream = helper;
            if (specialized) {
              launch.specialized = true;
              launch.block_done = skip;
              PB_CUDA(pb_contracted_api::LaunchSrkn(launch));
              PB_LAUNCHED();
              ++c.launches;
              launch.specialized = false;
              launch.skip_blocks = skip;
            }
            PB_CUDA(pb_contracted_api::LaunchSrkn(launch));
          } else if (variant == 5) {
            PB_LAUNCH_SRKN(256, 2);
          } else if (variant >= 2) {
            PB_LAUNCH_SRKN(384, 1);
          } else {
            switch (SrknMinBlocks()) {
              case 2: PB_LAUNCH_SRKN(128, 2); break;
              case 4: PB_LAUNCH_SRKN(128, 4); break;
              default: PB_LAUNCH_SRKN(128, 3); break;
            }
          }
#undef PB_LAUNCH_SRKN
          PB_LAUNCHED();
          ++c.launches;
        }
        if (groups > 1) {
          PB_CUDA(cudaEventRecord(w->flow_events[g], helper));
          PB_CUDA(cudaStreamWaitEvent(stream, w->flow_events[g], 0));
        }
      }
      PB_CUDA(cudaEventRecord(join, stream));
      ++c.phases;
    }
    steps_done += steps;
  }
  flow->time[0] = t.value;
  flow->time[1] = t.error;
  if (flow->n_steps != nullptr) {
    *flow->n_steps = steps_done;
  }
  if (flow->n_samples != nullptr) {
    *flow->n_samples = samples_done;
  }
  return PB_OK;
}

}  // namespace

// Ephemeris::NewInstance for massless trajectories (ephemeris.hpp:181-194): the
// integrator instance with its ODE::State on the device and, for the
// multistep integrators, previous_steps_.
struct pb_fixed_step_instance {
  pb_ephemeris* ephemeris = nullptr;
  int device = 0;
  // Counted in ephemeris->dependants (handles created through the C ABI).
  bool registered = false;
  int32_t method = 0;
  bool is_multistep = false;
  MultistepMethod multistep;
  double step = 0;
  long long n = 0;
  DP time{0.0, 0.0};
  double* state = nullptr;  // 12 planes of n: owned_state or the caller's.
  DeviceBuffer<double> owned_state;
  // One Solve at a time per instance.
  std::mutex mutex;
  int32_t arithmetic = PB_ARITHMETIC_EXACT;
  // Tabulated intrinsic accelerations (8 planes of n), or null.
  double const* burns = nullptr;
  DeviceBuffer<double> owned_burns;
  // previous_steps_: a ring of `order` steps.
  DeviceBuffer<double> displacement_value, displacement_error, acceleration;
  int previous_steps = 0;  // previous_steps_.size().
  long long startup_step_index = 0;
  int oldest = 0;  // Ring slot of previous_steps_.front().
  bool startup_target_valid = false;
  double startup_target = 0;
  MasslessHistory History() {
    return {displacement_value.data, displacement_error.data,
            acceleration.data};
  }
};

namespace {

struct InstanceSink {
  long long sample_every;
  long long max_samples;
  double* samples;       // Device.
  double* sample_times;  // Host.
  long long appended = 0;  // States appended by this call.
  long long recorded = 0;
};

pb_status_t PrepareInstance(pb_fixed_step_instance* s) {
  if (s->is_multistep && s->displacement_value.data == nullptr) {
    std::size_t const size =
        3 * static_cast<std::size_t>(s->n) * s->multistep.order;
    PB_RETURN_IF_ERROR(s->displacement_value.Reserve(size));
    PB_RETURN_IF_ERROR(s->displacement_error.Reserve(size));
    PB_RETURN_IF_ERROR(s->acceleration.Reserve(size));
  }
  return PB_OK;
}

// Starter::FillStepFromState at the instance's current state and time.
pb_status_t MasslessFillStep(FlowContext& c, pb_fixed_step_instance* s,
                             int const slot) {
  pb_ephemeris* const e = c.e;
  FlowWorkspace* const w = c.w();
  EphemerisView const view = c.View();
  PB_RETURN_IF_ERROR(w->stage_table.Reserve(view.stage_stride));
  double const t = s->time.value;
  PB_RETURN_IF_ERROR(
      BuildStageTable(e, w, &t, 1, w->stage_table.data, c.stream));
  PB_RETURN_IF_ERROR(c.CheckPath(w->stage_table.data, s->n, s->state));
  if (s->n > 0 && c.contracted()) {
    MasslessHistory const history = s->History();
    pb_contracted_api::FillStepLaunch launch{
        &view, w->stage_table.data, t, s->burns, s->n, s->state, &history,
        slot, w->status_word.data, c.stream};
    PB_CUDA(pb_contracted_api::LaunchFillStep(launch));
    LaunchCount().fetch_add(1);
  } else if (s->n > 0) {
    MasslessFillStepKernel<<<static_cast<unsigned>((s->n + kFlowThreads - 1) /
                                                   kFlowThreads),
                             kFlowThreads, 0, c.stream>>>(
        view, w->stage_table.data, t, s->burns, s->n, s->state, s->History(),
        slot, w->status_word.data);
    PB_LAUNCHED();
  }
  return PB_OK;
}

// append_state of the instance: the sink sees every appended state.
pb_status_t AppendInstanceState(pb_fixed_step_instance* s, InstanceSink* sink,
                                cudaStream_t stream) {
  ++sink->appended;
  if (sink->samples != nullptr && sink->sample_every > 0 &&
      sink->appended % sink->sample_every == 0 &&
      sink->recorded < sink->max_samples) {
    std::size_t const n = static_cast<std::size_t>(s->n);
    double* const out = sink->samples + sink->recorded * 6 * n;
    if (n > 0) {
      PB_CUDA(cudaMemcpyAsync(out, s->state, 3 * n * sizeof(double),
                              cudaMemcpyDeviceToDevice, stream));
      PB_CUDA(cudaMemcpyAsync(out + 3 * n, s->state + 6 * n,
                              3 * n * sizeof(double), cudaMemcpyDeviceToDevice,
                              stream));
    }
    if (sink->sample_times != nullptr) {
      sink->sample_times[sink->recorded] = s->time.value;
    }
    ++sink->recorded;
  }
  return PB_OK;
}

// Integrator::Instance::Solve(t_final) for a massless multistep instance:
// integrators/starter_body.hpp:23-74, then
// symmetric_linear_multistep_integrator_body.hpp:69-149.  The context has been
// begun and holds the constants.
pb_status_t SolveMultistepInstance(FlowContext& c, pb_fixed_step_instance* s,
                                   double const t_final, InstanceSink* sink) {
  pb_ephemeris* const e = c.e;
  FlowWorkspace* const w = c.w();
  cudaStream_t const stream = c.stream;
  PB_RETURN_IF_ERROR(PrepareInstance(s));
  EphemerisView const view = c.View();
  double const h = s->step;
  int const order = s->multistep.order;
  if (s->previous_steps < order) {
    if (s->previous_steps == 0) {
      PB_RETURN_IF_ERROR(MasslessFillStep(c, s, 0));
      s->previous_steps = 1;
    }
    double const startup_step = h / kStartupStepDivisor;
    if (!s->startup_target_valid) {
      s->startup_target = std::min(
          s->time.value + (order - s->previous_steps) * h + h / 2.0, t_final);
      s->startup_target_valid = true;
    }
    double const target = s->startup_target;
    // The startup instance: BlanesMoan2002SRKN14A at step / 16; every 16th
    // step becomes a step of the history and is appended.
    while (s->previous_steps < order &&
           startup_step <=
               std::abs((target - s->time.value) - s->time.error)) {
      long long const wanted =
          kStartupStepDivisor - s->startup_step_index % kStartupStepDivisor;
      double time[2] = {s->time.value, s->time.error};
      int64_t steps_taken = 0;
      pb_fixed_step_flow_t flow{};
      flow.method = PB_BLANES_MOAN_2002_SRKN_14A;
      flow.step = startup_step;
      flow.t_final = target;
      flow.n = s->n;
      flow.state = s->state;
      flow.time = time;
      flow.n_steps = &steps_taken;
      flow.burns = s->burns;
      PB_RETURN_IF_ERROR(SrknSteps(c, &flow, wanted));
      s->time = {time[0], time[1]};
      s->startup_step_index += steps_taken;
      if (c.cancelled) {
        // The startup target stays: the next Solve continues towards it.
        return PB_OK;
      }
      if (steps_taken == wanted) {
        PB_RETURN_IF_ERROR(MasslessFillStep(c, s, s->previous_steps));
        ++s->previous_steps;
        PB_RETURN_IF_ERROR(AppendInstanceState(s, sink, stream));
      } else {
        break;  // The target was reached between two steps of the history.
      }
    }
    s->startup_target_valid = false;
    if (s->previous_steps < order) {
      return PB_OK;  // "we'll continue the next time Solve is called".
    }
    s->oldest = 0;
  }
  // The reference ignores the status of the startup
  // (startup_instance->Solve(...).IgnoreError(), starter_body.hpp:68-73) and of
  // FillStepFromState (symmetric_linear_multistep_integrator_body.hpp:258-261):
  // only the steps below report collisions.
  PB_CUDA(cudaMemsetAsync(w->status_word.data, 0, sizeof(int32_t), stream));
  // Multistep steps, fused `chunk_steps` at a time.
  int const chunk_steps = 256;
  PB_RETURN_IF_ERROR(w->stage_table.Reserve(
      static_cast<std::size_t>(chunk_steps) * view.stage_stride));
  PB_RETURN_IF_ERROR(w->pinned_times.Reserve(chunk_steps));
  int sm_count = 0;
  PB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount,
                                 e->device));
  while (h <= (t_final - s->time.value) - s->time.error) {
    // The pinned staging buffer is reused across chunks; a stop request is
    // honoured here, between chunks (RETURN_IF_STOPPED,
    // symmetric_linear_multistep_integrator_body.hpp:149).
    PB_CUDA(cudaStreamSynchronize(stream));
    if (e->stop_requested.load() != 0) {
      c.cancelled = true;
      break;
    }
    int steps = 0;
    double* const times = w->pinned_times.data;
    long long const appended_before = sink->appended;
    DP t = s->time;
    while (steps < chunk_steps && h <= (t_final - t.value) - t.error) {
      Increment(t, h);
      times[steps] = t.value;
      ++steps;
      // The kernel writes the sample; the host mirrors the bookkeeping.
      ++sink->appended;
      if (sink->samples != nullptr && sink->sample_every > 0 &&
          sink->appended % sink->sample_every == 0 &&
          sink->recorded < sink->max_samples) {
        if (sink->sample_times != nullptr) {
          sink->sample_times[sink->recorded] = t.value;
        }
        ++sink->recorded;
      }
    }
    PB_RETURN_IF_ERROR(
        BuildStageTable(e, w, times, steps, w->stage_table.data, stream));
    PB_RETURN_IF_ERROR(c.CheckPath(w->stage_table.data, s->n, s->state));
    if (s->n > 0) {
      // Probe slices x step sub-chunks on helper streams, as in the SRKN flow
      // (no last-wave tail: 1.91 -> 2.04 10^9 probe-steps/s on 10^5 probes).
      bool const wide = !c.contracted() && MultistepWide(s->n, sm_count);
      int const threads = wide ? 256 : kFlowThreads;
      long long const blocks = (s->n + threads - 1) / threads;
      int const groups = static_cast<int>(
          std::min<long long>(FlowGroups(), std::max<long long>(1, blocks / 64)));
      int const sub_steps = groups > 1 ? MultistepSubSteps() : steps;
      PB_RETURN_IF_ERROR(c.EnsureHelpers(groups));
      cudaEvent_t const fork = w->timing_events[2 * c.phases];
      PB_CUDA(cudaEventRecord(fork, stream));
      bool const sampling = sink->samples != nullptr && sink->sample_every > 0;
      for (int g = 0; g < groups; ++g) {
        cudaStream_t const helper = groups > 1 ? w->flow_streams[g] : stream;
        if (groups > 1) {
          PB_CUDA(cudaStreamWaitEvent(helper, fork, 0));
        }
        long long const i_begin = blocks * g / groups * threads;
        long long const i_end =
            std::min<long long>(s->n, blocks * (g + 1) / groups * threads);
        unsigned const grid = static_cast<unsigned>(
            (i_end - i_begin + threads - 1) / threads);
        for (int s0 = 0; s0 < steps; s0 += sub_steps) {
          int const n_sub = std::min(sub_steps, steps - s0);
#define PB_LAUNCH_MULTISTEP(kOrder, kThreads, kMinBlocks)                     \
  MasslessMultistepFlowKernel<kOrder, kThreads, kMinBlocks>                   \
      <<<grid, kThreads, 0, helper>>>(                                        \
          view, s->multistep, h,                                              \
          w->stage_table.data +                                               \
              static_cast<std::size_t>(s0) * view.stage_stride,               \
          w->stage_times.data + s0, s->burns, n_sub, s->n, i_begin, i_end,    \
          s->state, s->History(), (s->oldest + s0) % order,                   \
          appended_before + s0, sampling ? sink->sample_every : 0,            \
          sink->max_samples, sink->samples, w->status_word.data)
          if (c.contracted()) {
            MasslessHistory const history = s->History();
            pb_contracted_api::MultistepLaunch launch;
            launch.view = &view;
            launch.method = &s->multistep;
            launch.h = h;
            launch.stage_table =
                w->stage_table.data +
                static_cast<std::size_t>(s0) * view.stage_stride;
            launch.stage_times = w->stage_times.data + s0;
            launch.burns = s->burns;
            launch.n_steps = n_sub;
            launch.n = s->n;
            launch.i_begin = i_begin;
            launch.i_end = i_end;
            launch.state = s->state;
            launch.history = &history;
            launch.oldest = (s->oldest + s0) % order;
            launch.steps_before = appended_before + s0;
            launch.sample_every = sampling ? sink->sample_every : 0;
            launch.max_samples = sink->max_samples;
            launch.samples = sink->samples;
            launch.status = w->status_word.data;
            launch.order = order;
            launch.grid = grid;
            launch.stream = helper;
            PB_CUDA(pb_contracted_api::LaunchMultistep(launch));
          } else if (order == 8) {
            if (wide) {
              PB_LAUNCH_MULTISTEP(8, 256, 2);
            } else {
              PB_LAUNCH_MULTISTEP(8, 128, 3);
            }
          } else {
            if (wide) {
              PB_LAUNCH_MULTISTEP(12, 256, 2);
            } else {
              PB_LAUNCH_MULTISTEP(12, 128, 3);
            }
          }
#undef PB_LAUNCH_MULTISTEP
          PB_LAUNCHED();
          ++c.launches;
        }
        if (groups > 1) {
          PB_CUDA(cudaEventRecord(w->flow_events[g], helper));
          PB_CUDA(cudaStreamWaitEvent(stream, w->flow_events[g], 0));
        }
      }
      PB_CUDA(cudaEventRecord(w->timing_events[2 * c.phases + 1], stream));
      ++c.phases;
    }
    s->oldest = (s->oldest + steps) % order;
    s->time = t;
  }
  return PB_OK;
}

// Instance::Solve(t_final) for either family; the outs of `flow` are filled.
// Returns PB_CANCELLED if a stop request cut the flow short (the instance is
// then at its last completed step and can be solved further).
pb_status_t SolveInstance(FlowContext& c, pb_fixed_step_instance* s,
                          const pb_fixed_step_flow_t* flow) {
  PB_RETURN_IF_ERROR(c.AcquireConstants(s->time.value, s->n, s->state,
                                        nullptr, s->arithmetic));
  if (!s->is_multistep) {
    pb_fixed_step_flow_t srkn = *flow;
    srkn.method = s->method;
    srkn.step = s->step;
    srkn.n = s->n;
    srkn.state = s->state;
    srkn.burns = s->burns;
    srkn.status = nullptr;
    double time[2] = {s->time.value, s->time.error};
    srkn.time = time;
    PB_RETURN_IF_ERROR(
        SrknSteps(c, &srkn, std::numeric_limits<long long>::max()));
    s->time = {time[0], time[1]};
    PB_RETURN_IF_ERROR(c.Finish(flow->status));
    return c.cancelled ? Fail(PB_CANCELLED, "stop requested") : PB_OK;
  }
  InstanceSink sink{flow->sample_every, flow->max_samples, flow->samples,
                    flow->sample_times};
  PB_RETURN_IF_ERROR(SolveMultistepInstance(c, s, flow->t_final, &sink));
  PB_RETURN_IF_ERROR(c.Finish(flow->status));
  if (flow->n_steps != nullptr) {
    *flow->n_steps = sink.appended;
  }
  if (flow->n_samples != nullptr) {
    *flow->n_samples = sink.recorded;
  }
  return c.cancelled ? Fail(PB_CANCELLED, "stop requested") : PB_OK;
}

// NewInstance + FlowWithFixedStep in one call for a multistep method: the
// instance (and the history it would keep) lives for this call only.
pb_status_t MultistepFlowDevice(FlowContext& c,
                                const pb_fixed_step_flow_t* flow) {
  if (!(flow->step > 0)) {
    return Fail(PB_INVALID_ARGUMENT,
                "the multistep massless flow integrates forward");
  }
  pb_fixed_step_instance instance;
  instance.ephemeris = c.e;
  instance.device = c.e->device;
  instance.method = flow->method;
  instance.is_multistep = true;
  MakeMultistepMethod(flow->method, instance.multistep);
  instance.step = flow->step;
  instance.n = flow->n;
  instance.time = {flow->time[0], flow->time[1]};
  instance.state = flow->state;
  instance.burns = flow->burns;
  instance.arithmetic = flow->arithmetic;
  pb_status_t const status = SolveInstance(c, &instance, flow);
  flow->time[0] = instance.time.value;
  flow->time[1] = instance.time.error;
  return status;
}

}  // namespace

extern "C" {

pb_status_t pb_fixed_step_instance_create(pb_ephemeris_t* e, int32_t method,
                                          double step, int64_t n,
                                          const double* state,
                                          const double* time,
                                          pb_fixed_step_instance_t** instance) {
  if (e == nullptr || instance == nullptr || n < 0 || time == nullptr ||
      (n > 0 && state == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  *instance = nullptr;
  if (step == 0 || step != step) {
    return Fail(PB_INVALID_ARGUMENT, "CHECK_NE(Time(), step)");
  }
  PB_CUDA(cudaSetDevice(e->device));
  auto* s = new pb_fixed_step_instance;
  s->ephemeris = e;
  s->device = e->device;
  s->method = method;
  s->step = step;
  s->n = n;
  s->time = {time[0], time[1]};
  SrknTableau tableau;
  if (MakeMultistepMethod(method, s->multistep)) {
    s->is_multistep = true;
    if (!(step > 0)) {
      delete s;
      return Fail(PB_INVALID_ARGUMENT,
                  "the multistep massless flow integrates forward");
    }
  } else if (!MakeSrknTableau(method, tableau)) {
    delete s;
    return Fail(PB_INVALID_ARGUMENT, "unsupported fixed-step method");
  }
  pb_status_t status =
      s->owned_state.Upload(state, 12 * static_cast<std::size_t>(n), nullptr);
  if (status == PB_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) {
    status = Fail(PB_INTERNAL, "upload failed");
  }
  if (status != PB_OK) {
    delete s;
    return status;
  }
  s->state = s->owned_state.data;
  s->registered = true;
  e->dependants.fetch_add(1);
  *instance = s;
  return PB_OK;
}

void pb_fixed_step_instance_destroy(pb_fixed_step_instance_t* instance) {
  PB_API_RANGE();
  if (instance != nullptr) {
    // The device ordinal is the instance's own copy; the ephemeris cannot have
    // been destroyed under a registered instance (pb_ephemeris_destroy
    // refuses).
    cudaSetDevice(instance->device);
    if (instance->registered) {
      instance->ephemeris->dependants.fetch_sub(1);
    }
    delete instance;
  }
}

pb_status_t pb_fixed_step_instance_set_intrinsic_accelerations(
    pb_fixed_step_instance_t* s, const double* burns) {
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null instance");
  }
  PB_CUDA(cudaSetDevice(s->device));
  if (burns == nullptr) {
    s->burns = nullptr;
    return PB_OK;
  }
  PB_RETURN_IF_ERROR(s->owned_burns.Upload(
      burns, PB_BURN_PLANES * static_cast<std::size_t>(s->n), nullptr));
  PB_CUDA(cudaStreamSynchronize(nullptr));
  s->burns = s->owned_burns.data;
  return PB_OK;
}

pb_status_t pb_fixed_step_instance_set_arithmetic(
    pb_fixed_step_instance_t* s, int32_t arithmetic) {
  if (s == nullptr || (arithmetic != PB_ARITHMETIC_EXACT &&
                       arithmetic != PB_ARITHMETIC_CONTRACTED)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  std::lock_guard<std::mutex> const instance_lock(s->mutex);
  s->arithmetic = arithmetic;
  return PB_OK;
}

pb_status_t pb_fixed_step_instance_flow(pb_fixed_step_instance_t* s,
                                        double t_final, int64_t sample_every,
                                        int64_t max_samples, double* samples,
                                        double* sample_times,
                                        int64_t* n_steps, int64_t* n_samples,
                                        pb_status_t* flow_status) {
  PB_API_RANGE();
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null instance");
  }
  // An instance is not a concurrent object (one Solve at a time, as in the
  // reference); distinct instances flow concurrently.
  std::lock_guard<std::mutex> const instance_lock(s->mutex);
  FlowContext c;
  PB_RETURN_IF_ERROR(c.Begin(s->ephemeris, false, nullptr));
  FlowWorkspace* const w = c.w();
  std::size_t const n = static_cast<std::size_t>(s->n);
  bool const sampling =
      samples != nullptr && sample_every > 0 && max_samples > 0;
  pb_fixed_step_flow_t flow{};
  flow.t_final = t_final;
  flow.sample_every = sampling ? sample_every : 0;
  flow.max_samples = max_samples;
  flow.sample_times = sample_times;
  int64_t recorded = 0;
  flow.n_samples = &recorded;
  flow.n_steps = n_steps;
  flow.status = flow_status;
  if (sampling) {
    PB_RETURN_IF_ERROR(
        w->scratch_c.Reserve(static_cast<std::size_t>(max_samples) * 6 * n));
    flow.samples = w->scratch_c.data;
  }
  pb_status_t const status = SolveInstance(c, s, &flow);
  if (status != PB_OK && status != PB_CANCELLED) {
    return status;
  }
  if (sampling && recorded > 0 && n > 0) {
    PB_CUDA(cudaMemcpyAsync(
        samples, w->scratch_c.data,
        static_cast<std::size_t>(recorded) * 6 * n * sizeof(double),
        cudaMemcpyDeviceToHost, c.stream));
  }
  PB_CUDA(cudaStreamSynchronize(c.stream));
  if (n_samples != nullptr) {
    *n_samples = recorded;
  }
  return status;
}

pb_status_t pb_fixed_step_instance_flow_to_downsampler(
    pb_fixed_step_instance_t* s, double t_final,
    pb_downsampler_t* downsampler, int64_t* n_steps,
    pb_status_t* flow_status) {
  PB_API_RANGE();
  if (s == nullptr || downsampler == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  std::lock_guard<std::mutex> const instance_lock(s->mutex);
  FlowContext c;
  PB_RETURN_IF_ERROR(c.Begin(s->ephemeris, false, nullptr));
  FlowWorkspace* const w = c.w();
  cudaStream_t const stream = c.stream;
  std::size_t const n = static_cast<std::size_t>(s->n);
  // States per chunk: at most 64 MiB of samples on the device.
  long long const chunk = std::max<long long>(
      1, std::min<long long>(256, (64LL << 20) / (48 * std::max<long long>(
                                                           1, s->n))));
  // Two spare slots: the time bound below admits one more state when the
  // instance sits between two steps of a multistep startup.
  long long const slots = chunk + 2;
  PB_RETURN_IF_ERROR(
      w->scratch_c.Reserve(static_cast<std::size_t>(slots) * 6 * n + 1));
  std::vector<double> sample_times(slots);
  int64_t total_steps = 0;
  pb_status_t overall = PB_OK;
  pb_status_t result = PB_OK;
  double const abs_step = std::abs(s->step);
  for (;;) {
    // Instance::Solve(t) for the next `chunk` steps at most; during the startup
    // of a multistep instance only every 16th step of the starter is appended,
    // hence the bound on the time rather than on the steps.
    double const remaining = (t_final - s->time.value) - s->time.error;
    if (!(abs_step <= std::abs(remaining))) {
      break;
    }
    double t_chunk = s->time.value + (s->step > 0 ? 1 : -1) *
                                         (static_cast<double>(chunk) + 0.5) *
                                         abs_step;
    if (s->step > 0 ? t_chunk > t_final : t_chunk < t_final) {
      t_chunk = t_final;
    }
    pb_fixed_step_flow_t flow{};
    flow.t_final = t_chunk;
    flow.sample_every = 1;
    flow.max_samples = slots;
    flow.samples = w->scratch_c.data;
    flow.sample_times = sample_times.data();
    int64_t steps = 0, recorded = 0;
    pb_status_t status = PB_OK;
    flow.n_steps = &steps;
    flow.n_samples = &recorded;
    flow.status = &status;
    DP const before = s->time;
    result = SolveInstance(c, s, &flow);
    if (result != PB_OK && result != PB_CANCELLED) {
      return result;
    }
    if (recorded != steps) {
      return Fail(PB_INTERNAL, "more states appended than the chunk holds");
    }
    PB_RETURN_IF_ERROR(pb_downsampler_append_device(
        downsampler, recorded, sample_times.data(), w->scratch_c.data,
        stream));
    total_steps += steps;
    if (overall == PB_OK) {
      overall = status;
    }
    if (result == PB_CANCELLED ||
        (s->time.value == before.value && s->time.error == before.error)) {
      break;  // Stopped, or no progress: t_final is closer than one step.
    }
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  if (n_steps != nullptr) {
    *n_steps = total_steps;
  }
  if (flow_status != nullptr) {
    *flow_status = overall;
  }
  return result;
}

// The checkpoint of an instance: a header of kCheckpointHeader doubles, the
// 12 planes of the state, the 8 planes of the burns if any, and for a multistep
// instance the three history rings (3 n order doubles each).
namespace {
constexpr int kCheckpointHeader = 16;
constexpr double kCheckpointMagic = 20261019.0;

std::size_t CheckpointDoubles(pb_fixed_step_instance const* s) {
  std::size_t const n = static_cast<std::size_t>(s->n);
  std::size_t size = kCheckpointHeader + 12 * n;
  if (s->burns != nullptr) {
    size += PB_BURN_PLANES * n;
  }
  if (s->is_multistep && s->displacement_value.data != nullptr) {
    size += 3 * (3 * n * s->multistep.order);
  }
  return size;
}
}  // namespace

int64_t pb_fixed_step_instance_checkpoint_size(
    const pb_fixed_step_instance_t* s) {
  return s == nullptr ? 0 : static_cast<int64_t>(CheckpointDoubles(s));
}

pb_status_t pb_fixed_step_instance_write_checkpoint(
    pb_fixed_step_instance_t* s, double* checkpoint) {
  PB_API_RANGE();
  if (s == nullptr || checkpoint == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  std::lock_guard<std::mutex> const instance_lock(s->mutex);
  PB_CUDA(cudaSetDevice(s->device));
  std::size_t const n = static_cast<std::size_t>(s->n);
  bool const has_history =
      s->is_multistep && s->displacement_value.data != nullptr;
  double* out = checkpoint;
  double const header[kCheckpointHeader] = {
      kCheckpointMagic,
      static_cast<double>(s->method),
      s->step,
      static_cast<double>(s->n),
      s->time.value,
      s->time.error,
      static_cast<double>(s->previous_steps),
      static_cast<double>(s->startup_step_index),
      static_cast<double>(s->oldest),
      s->startup_target_valid ? 1.0 : 0.0,
      s->startup_target,
      s->burns != nullptr ? 1.0 : 0.0,
      has_history ? 1.0 : 0.0,
      0.0, 0.0, 0.0};
  std::memcpy(out, header, sizeof(header));
  out += kCheckpointHeader;
  auto download = [&](double const* device, std::size_t const count) {
    if (count == 0) {
      return cudaSuccess;
    }
    cudaError_t const error = cudaMemcpy(out, device, count * sizeof(double),
                                         cudaMemcpyDeviceToHost);
    out += count;
    return error;
  };
  PB_CUDA(download(s->state, 12 * n));
  if (s->burns != nullptr) {
    PB_CUDA(download(s->burns, PB_BURN_PLANES * n));
  }
  if (has_history) {
    std::size_t const ring = 3 * n * s->multistep.order;
    PB_CUDA(download(s->displacement_value.data, ring));
    PB_CUDA(download(s->displacement_error.data, ring));
    PB_CUDA(download(s->acceleration.data, ring));
  }
  return PB_OK;
}

pb_status_t pb_fixed_step_instance_read_checkpoint(
    pb_ephemeris_t* e, const double* checkpoint, int64_t size,
    pb_fixed_step_instance_t** instance) {
  if (e == nullptr || checkpoint == nullptr || instance == nullptr ||
      size < kCheckpointHeader || checkpoint[0] != kCheckpointMagic) {
    return Fail(PB_INVALID_ARGUMENT, "not a checkpoint of an instance");
  }
  *instance = nullptr;
  double const time[2] = {checkpoint[4], checkpoint[5]};
  int64_t const n = static_cast<int64_t>(checkpoint[3]);
  bool const has_burns = checkpoint[11] != 0.0;
  bool const has_history = checkpoint[12] != 0.0;
  if (n < 0 || size < kCheckpointHeader + 12 * n) {
    return Fail(PB_INVALID_ARGUMENT, "truncated checkpoint");
  }
  double const* in = checkpoint + kCheckpointHeader;
  pb_fixed_step_instance_t* s = nullptr;
  PB_RETURN_IF_ERROR(pb_fixed_step_instance_create(
      e, static_cast<int32_t>(checkpoint[1]), checkpoint[2], n, in, time, &s));
  in += 12 * n;
  std::size_t const count = static_cast<std::size_t>(n);
  std::size_t expected = kCheckpointHeader + 12 * count +
                         (has_burns ? PB_BURN_PLANES * count : 0);
  if (has_history) {
    expected += 3 * (3 * count * s->multistep.order);
  }
  pb_status_t status = PB_OK;
  if (static_cast<std::size_t>(size) != expected ||
      (has_history && !s->is_multistep)) {
    status = Fail(PB_INVALID_ARGUMENT, "inconsistent checkpoint");
  }
  if (status == PB_OK && has_burns) {
    status = pb_fixed_step_instance_set_intrinsic_accelerations(s, in);
    in += PB_BURN_PLANES * count;
  }
  if (status == PB_OK && has_history) {
    std::size_t const ring = 3 * count * s->multistep.order;
    status = s->displacement_value.Upload(in, ring, nullptr);
    if (status == PB_OK) {
      status = s->displacement_error.Upload(in + ring, ring, nullptr);
    }
    if (status == PB_OK) {
      status = s->acceleration.Upload(in + 2 * ring, ring, nullptr);
    }
    if (status == PB_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) {
      status = Fail(PB_INTERNAL, "upload failed");
    }
  }
  if (status != PB_OK) {
    pb_fixed_step_instance_destroy(s);
    return status;
  }
  s->previous_steps = static_cast<int>(checkpoint[6]);
  s->startup_step_index = static_cast<long long>(checkpoint[7]);
  s->oldest = static_cast<int>(checkpoint[8]);
  s->startup_target_valid = checkpoint[9] != 0.0;
  s->startup_target = checkpoint[10];
  *instance = s;
  return PB_OK;
}

pb_status_t pb_fixed_step_instance_get_state(pb_fixed_step_instance_t* s,
                                             double* state, double* time) {
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null instance");
  }
  std::lock_guard<std::mutex> const instance_lock(s->mutex);
  PB_CUDA(cudaSetDevice(s->device));
  if (s->n > 0 && state != nullptr) {
    PB_CUDA(cudaMemcpy(state, s->state,
                       12 * static_cast<std::size_t>(s->n) * sizeof(double),
                       cudaMemcpyDeviceToHost));
  }
  if (time != nullptr) {
    time[0] = s->time.value;
    time[1] = s->time.error;
  }
  return PB_OK;
}

}  // extern "C"

namespace {

pb_status_t ValidateFixedStepFlow(pb_ephemeris_t* e,
                                  const pb_fixed_step_flow_t* flow) {
  if (e == nullptr || flow == nullptr || flow->n < 0 || flow->time == nullptr ||
      (flow->n > 0 && flow->state == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  if (flow->step == 0 || flow->step != flow->step) {
    return Fail(PB_INVALID_ARGUMENT, "CHECK_NE(Time(), step)");
  }
  if (flow->time[0] != flow->time[0] || flow->t_final != flow->t_final) {
    return Fail(PB_INVALID_ARGUMENT, "NaN time");
  }
  return PB_OK;
}

// NewInstance + FlowWithFixedStep on device state.  `probe_host`: see
// AcquireFlowConstants.
pb_status_t FixedStepFlowDevice(FlowContext& c,
                                const pb_fixed_step_flow_t* flow,
                                double const* const probe_host) {
  PB_RETURN_IF_ERROR(c.AcquireConstants(flow->time[0], flow->n, flow->state,
                                        probe_host, flow->arithmetic));
  MultistepMethod multistep;
  if (MakeMultistepMethod(flow->method, multistep)) {
    return MultistepFlowDevice(c, flow);
  }
  pb_fixed_step_flow_t srkn = *flow;
  srkn.status = nullptr;
  PB_RETURN_IF_ERROR(
      SrknSteps(c, &srkn, std::numeric_limits<long long>::max()));
  PB_RETURN_IF_ERROR(c.Finish(flow->status));
  return c.cancelled ? Fail(PB_CANCELLED, "stop requested") : PB_OK;
}

}  // namespace

extern "C" {

pb_status_t pb_flow_with_fixed_step_device(pb_ephemeris_t* e,
                                           const pb_fixed_step_flow_t* flow,
                                           void* stream_pointer) {
  PB_API_RANGE();
  PB_RETURN_IF_ERROR(ValidateFixedStepFlow(e, flow));
  FlowContext c;
  PB_RETURN_IF_ERROR(
      c.Begin(e, true, static_cast<cudaStream_t>(stream_pointer)));
  return FixedStepFlowDevice(c, flow, nullptr);
}

pb_status_t pb_flow_with_fixed_step(pb_ephemeris_t* e,
                                    const pb_fixed_step_flow_t* flow) {
  PB_API_RANGE();
  PB_RETURN_IF_ERROR(ValidateFixedStepFlow(e, flow));
  FlowContext c;
  PB_RETURN_IF_ERROR(c.Begin(e, false, nullptr));
  FlowWorkspace* const w = c.w();
  cudaStream_t const stream = c.stream;
  std::size_t const n = static_cast<std::size_t>(flow->n);
  PB_RETURN_IF_ERROR(w->scratch_a.Upload(flow->state, 12 * n, stream));
  pb_fixed_step_flow_t device_flow = *flow;
  device_flow.state = w->scratch_a.data;
  if (flow->burns != nullptr) {
    PB_RETURN_IF_ERROR(
        w->burns_scratch.Upload(flow->burns, PB_BURN_PLANES * n, stream));
    device_flow.burns = w->burns_scratch.data;
  }
  bool const sampling = flow->samples != nullptr && flow->sample_every > 0 &&
                        flow->max_samples > 0;
  if (sampling) {
    PB_RETURN_IF_ERROR(w->scratch_c.Reserve(
        static_cast<std::size_t>(flow->max_samples) * 6 * n));
    device_flow.samples = w->scratch_c.data;
  } else {
    device_flow.samples = nullptr;
  }
  int64_t n_samples = 0;
  device_flow.n_samples = &n_samples;
  pb_status_t const status = FixedStepFlowDevice(c, &device_flow, flow->state);
  if (status != PB_OK && status != PB_CANCELLED) {
    return status;
  }
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(flow->state, w->scratch_a.data,
                            12 * n * sizeof(double), cudaMemcpyDeviceToHost,
                            stream));
    if (sampling && n_samples > 0) {
      PB_CUDA(cudaMemcpyAsync(flow->samples, w->scratch_c.data,
                              static_cast<std::size_t>(n_samples) * 6 * n *
                                  sizeof(double),
                              cudaMemcpyDeviceToHost, stream));
    }
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  if (flow->n_samples != nullptr) {
    *flow->n_samples = n_samples;
  }
  return status;
}

double pb_ephemeris_last_flow_kernel_ms(const pb_ephemeris_t* e,
                                        int64_t* launches) {
  std::lock_guard<std::mutex> const lock(e->statistics_mutex);
  if (launches != nullptr) {
    *launches = e->last_flow_kernel_launches;
  }
  return e->last_flow_kernel_ms;
}

pb_status_t pb_ephemeris_last_flow_path(const pb_ephemeris_t* e,
                                        int32_t* hot_body,
                                        int32_t* generic_harmonics) {
  if (e == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null ephemeris");
  }
  std::lock_guard<std::mutex> const lock(e->statistics_mutex);
  if (hot_body != nullptr) {
    *hot_body = e->last_flow_hot_body;
  }
  if (generic_harmonics != nullptr) {
    *generic_harmonics =
        (e->last_flow_path_flags & kPathGenericHarmonics) != 0 ? 1 : 0;
  }
  return PB_OK;
}

pb_status_t pb_measure_fp64_peak(int32_t cuda_device, double* flops_per_second,
                                 double* ms) {
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0) {
    return Fail(PB_FAILED_PRECONDITION, "no CUDA device");
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  cudaDeviceProp properties;
  PB_CUDA(cudaGetDeviceProperties(&properties, cuda_device));
  double* out = nullptr;
  PB_CUDA(cudaMalloc(&out, sizeof(double)));
  int const iterations = 1 << 16;
  int const block = 256;
  int const grid = properties.multiProcessorCount * 8;
  cudaEvent_t start, stop;
  PB_CUDA(cudaEventCreate(&start));
  PB_CUDA(cudaEventCreate(&stop));
  float best = std::numeric_limits<float>::infinity();
  for (int repeat = 0; repeat < 4; ++repeat) {
    PB_CUDA(cudaEventRecord(start));
    Fp64PeakKernel<<<grid, block>>>(out, iterations);
    PB_LAUNCHED();
    PB_CUDA(cudaEventRecord(stop));
    PB_CUDA(cudaEventSynchronize(stop));
    float elapsed = 0;
    PB_CUDA(cudaEventElapsedTime(&elapsed, start, stop));
    if (repeat > 0) {
      best = std::min(best, elapsed);
    }
  }
  PB_CUDA(cudaEventDestroy(start));
  PB_CUDA(cudaEventDestroy(stop));
  PB_CUDA(cudaFree(out));
  double const flops = 2.0 * 8.0 * iterations * static_cast<double>(block) * grid;
  if (flops_per_second != nullptr) {
    *flops_per_second = flops / (best * 1e-3);
  }
  if (ms != nullptr) {
    *ms = best;
  }
  return PB_OK;
}

namespace {
// A stream and a mapped result of its own per calling thread: nothing here
// waits on, or is waited on by, the flows in progress.
struct SmClockProbe {
  cudaStream_t stream = nullptr;
  double* result = nullptr;
  int device = -1;
};
pb_status_t AcquireSmClockProbe(int32_t const cuda_device,
                                SmClockProbe*& out) {
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0) {
    return Fail(PB_FAILED_PRECONDITION, "no CUDA device");
  }
  PB_CUDA(cudaSetDevice(cuda_device));
  thread_local SmClockProbe probe;
  if (probe.device != cuda_device) {
    if (probe.stream != nullptr) {
      cudaStreamDestroy(probe.stream);
      cudaFreeHost(probe.result);
      probe = SmClockProbe();
    }
    int least = 0, greatest = 0;
    PB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    PB_CUDA(cudaStreamCreateWithPriority(&probe.stream, cudaStreamNonBlocking,
                                         greatest));
    PB_CUDA(cudaMallocHost(&probe.result, sizeof(double)));
    *probe.result = 0.0;
    probe.device = cuda_device;
  }
  out = &probe;
  return PB_OK;
}
}  // namespace

pb_status_t pb_sm_clock_probe_start(int32_t cuda_device,
                                    double delay_microseconds,
                                    double microseconds) {
  if (!(microseconds > 0) || microseconds > 1e5 ||
      !(delay_microseconds >= 0) || delay_microseconds > 1e5) {
    return Fail(PB_INVALID_ARGUMENT, "pb_sm_clock_probe_start: bad duration");
  }
  SmClockProbe* probe = nullptr;
  PB_RETURN_IF_ERROR(AcquireSmClockProbe(cuda_device, probe));
  SmClockKernel<<<1, 32, 0, probe->stream>>>(
      static_cast<unsigned long long>(delay_microseconds * 1e3),
      static_cast<unsigned long long>(microseconds * 1e3), probe->result);
  PB_LAUNCHED();
  return PB_OK;
}

pb_status_t pb_sm_clock_probe_read(int32_t cuda_device, double* mhz) {
  if (mhz == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "pb_sm_clock_probe_read: null result");
  }
  SmClockProbe* probe = nullptr;
  PB_RETURN_IF_ERROR(AcquireSmClockProbe(cuda_device, probe));
  PB_CUDA(cudaStreamSynchronize(probe->stream));
  *mhz = *probe->result;
  return PB_OK;
}

pb_status_t pb_measure_sm_clock(int32_t cuda_device, double microseconds,
                                double* mhz) {
  PB_RETURN_IF_ERROR(pb_sm_clock_probe_start(cuda_device, 0.0, microseconds));
  return pb_sm_clock_probe_read(cuda_device, mhz);
}

}  // extern "C"
