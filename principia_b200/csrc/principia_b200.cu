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
#include "pb_kernels_massless.cuh"
#include "pb_kernels_massless_multistep.cuh"
#include "pb_runtime.cuh"

using namespace pb;

namespace pb {

pb_status_t UploadSegments(pb_ephemeris* e, cudaStream_t stream) {
  if (!e->segments_dirty) {
    return PB_OK;
  }
  int const n = static_cast<int>(e->trajectories.size());
  std::vector<int32_t> offset(n + 1, 0);
  std::vector<double> t_max, origin, coefficients;
  std::vector<int32_t> degree;
  for (int b = 0; b < n; ++b) {
    HostTrajectory const& tr = e->trajectories[b];
    offset[b] = static_cast<int32_t>(t_max.size());
    t_max.insert(t_max.end(), tr.t_max.begin(), tr.t_max.end());
    origin.insert(origin.end(), tr.origin.begin(), tr.origin.end());
    degree.insert(degree.end(), tr.degree.begin(), tr.degree.end());
    // Host [segment][18][3] -> device [segment][3][18].
    std::size_t const n_segments = tr.t_max.size();
    std::size_t const base = coefficients.size();
    coefficients.resize(base + n_segments * 3 * kSegmentCoefficients);
    for (std::size_t segment = 0; segment < n_segments; ++segment) {
      double const* const in =
          tr.coefficients.data() + segment * 3 * kSegmentCoefficients;
      double* const out =
          coefficients.data() + base + segment * 3 * kSegmentCoefficients;
      for (int k = 0; k < kSegmentCoefficients; ++k) {
        for (int c = 0; c < 3; ++c) {
          out[c * kSegmentCoefficients + k] = in[3 * k + c];
        }
      }
    }
  }
  offset[n] = static_cast<int32_t>(t_max.size());
  // Bodies appended in lock step (Ephemeris::Prolong) share their segment
  // boundaries; kernels then look a time up once for all bodies.
  e->uniform_segments = n > 0;
  for (int b = 1; b < n; ++b) {
    if (e->trajectories[b].t_max != e->trajectories[0].t_max) {
      e->uniform_segments = false;
    }
  }
  // The copies below are from pageable memory and therefore synchronous with
  // respect to the host, so the temporaries may die at the end of this scope.
  PB_RETURN_IF_ERROR(e->segment_offset.Upload(offset, stream));
  PB_RETURN_IF_ERROR(e->segment_t_max.Upload(t_max, stream));
  PB_RETURN_IF_ERROR(e->segment_origin.Upload(origin, stream));
  PB_RETURN_IF_ERROR(e->segment_degree.Upload(degree, stream));
  PB_RETURN_IF_ERROR(e->segment_coefficients.Upload(coefficients, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  e->segments_dirty = false;
  return PB_OK;
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
pb_status_t BuildStageTable(pb_ephemeris* e, double const* times_host,
                            int n_times, double* table, cudaStream_t stream) {
  PB_RETURN_IF_ERROR(UploadSegments(e, stream));
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
  PB_RETURN_IF_ERROR(e->stage_times.Upload(times_host, n_times, stream));
  EphemerisView const view = e->View();
  long long const threads = static_cast<long long>(n_times) * view.n_bodies;
  int const block = 128;
  StageTableKernel<<<static_cast<unsigned>((threads + block - 1) / block),
                     block, 0, stream>>>(view, e->stage_times.data, n_times,
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

// Uploads the constant-bank tables of the flow kernel: per-body scalars and the
// tables of the "hot" oblate body -- the body of degree <= kHotDegree with the
// highest limiting degree at the distance of the first probe (the Earth for
// LEO probes).  The choice only affects speed: both paths compute the same
// bits.
pb_status_t UploadFlowConstants(pb_ephemeris* e, double const t,
                                long long const n, double const* state_device,
                                cudaStream_t stream) {
  int const n_bodies = static_cast<int>(e->model.bodies.size());
  if (n_bodies > kMaxConstantBodies) {
    return Fail(PB_INVALID_ARGUMENT, "too many bodies for the flow kernel");
  }
  static std::uint64_t uploaded_for = 0;
  static int uploaded_hot = -2;
  // Position of the first probe.
  double probe[3] = {0.0, 0.0, 0.0};
  if (n > 0) {
    for (int c = 0; c < 3; ++c) {
      PB_CUDA(cudaMemcpyAsync(&probe[c], state_device + c * n, sizeof(double),
                              cudaMemcpyDeviceToHost, stream));
    }
    PB_CUDA(cudaStreamSynchronize(stream));
  }
  int hot = -1;
  int best_degree = 2;
  for (int b = 0; b < e->model.n_oblate; ++b) {
    DeviceBody const& body = e->model.bodies[b];
    if (body.degree > kHotDegree) {
      continue;
    }
    HostTrajectory const& tr = e->trajectories[b];
    if (tr.t_max.empty()) {
      continue;
    }
    std::size_t segment =
        std::lower_bound(tr.t_max.begin(), tr.t_max.end(), t) -
        tr.t_max.begin();
    segment = std::min(segment, tr.t_max.size() - 1);
    double const x = t - tr.origin[segment];
    double const* const c = &tr.coefficients[segment * 3 * kSegmentCoefficients];
    double d2 = 0;
    for (int k = 0; k < 3; ++k) {
      double const p = EstrinEvaluate<3>(c + k, tr.degree[segment], x);
      d2 += (p - probe[k]) * (p - probe[k]);
    }
    int const limiting = LimitingDegree(
        e->model.dampings.data() + body.damping_offset, body.n_damping,
        std::sqrt(d2));
    if (limiting - 1 > best_degree) {
      best_degree = limiting - 1;
      hot = b;
    }
  }
  if (uploaded_for == e->serial && uploaded_hot == hot) {
    return PB_OK;
  }
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
  if (hot >= 0) {
    DeviceBody const& body = e->model.bodies[hot];
    int const size = (body.degree + 1) * (body.degree + 2) / 2;
    hot_body.n_damping = body.n_damping;
    for (int k = 0; k < size; ++k) {
      hot_body.cos[k] = e->model.cos[body.coefficient_offset + k];
      hot_body.sin[k] = e->model.sin[body.coefficient_offset + k];
    }
    for (int k = 0; k < kHotSize; ++k) {
      hot_body.normalization[k] = pb_legendre_normalization_factor[k];
    }
    for (int d = 0; d < body.n_damping; ++d) {
      hot_body.degree_damping[d] = e->model.dampings[body.damping_offset + d];
    }
    hot_body.sectoral = body.sectoral;
  }
  // The previous launch may still read the symbols on this stream: the copies
  // are stream-ordered.
  PB_CUDA(cudaMemcpyToSymbolAsync(c_bodies, &scalars, sizeof(scalars), 0,
                                  cudaMemcpyHostToDevice, stream));
  PB_CUDA(cudaMemcpyToSymbolAsync(c_hot, &hot_body, sizeof(hot_body), 0,
                                  cudaMemcpyHostToDevice, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  uploaded_for = e->serial;
  uploaded_hot = hot;
  return PB_OK;
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
int FlowSubSteps() {
  static int const value = [] {
    char const* const text = std::getenv("PB_FLOW_SUB_STEPS");
    int const v = text != nullptr ? std::atoi(text) : 10;
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

// Shape of the flow kernel: 0 = 128 threads x SrknMinBlocks() blocks per SM;
// 3 = 384 threads x 1 block (168 registers, 12 warps per SM); 5 = 256 threads x
// 2 blocks (128 registers, 16 warps per SM).  Measured on 10^5 probes: 56.2,
// 50.7 and 48.4 ms per 100 steps (a block barrier per stage did not help; 18,
// 20 and 24 warps per SM at 96 and 80 registers: 69, 66 and 74 ms), so a launch
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
  status = e->bodies.Upload(e->model.bodies, stream);
  if (status == PB_OK) status = e->dampings.Upload(dampings, stream);
  if (status == PB_OK) status = e->cos.Upload(cos, stream);
  if (status == PB_OK) status = e->sin.Upload(sin, stream);
  if (status == PB_OK) status = e->normalization.Upload(normalization, stream);
  if (status == PB_OK) status = e->status_word.Reserve(4);
  if (status == PB_OK) status = e->pinned_status.Reserve(4);
  if (status == PB_OK) {
    cudaError_t const sync_error = cudaStreamSynchronize(stream);
    if (sync_error != cudaSuccess) {
      status = Fail(PB_INTERNAL, cudaGetErrorString(sync_error));
    }
  }
  if (status != PB_OK) {
    delete e;
    return status;
  }
  *ephemeris = e;
  return PB_OK;
}

void pb_ephemeris_destroy(pb_ephemeris_t* ephemeris) {
  if (ephemeris != nullptr) {
    cudaSetDevice(ephemeris->device);
    delete ephemeris;
  }
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
  if (e == nullptr || body < 0 ||
      body >= static_cast<int32_t>(e->model.bodies.size()) || n_segments < 0) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  HostTrajectory& tr = e->trajectories[e->model.internal_of_caller[body]];
  for (int s = 0; s < n_segments; ++s) {
    if (degree[s] < 0 || degree[s] >= kSegmentCoefficients) {
      return Fail(PB_INVALID_ARGUMENT, "segment degree out of range");
    }
    double const previous = tr.t_max.empty() ? t_min : tr.t_max.back();
    if (!(t_max[s] > previous)) {
      return Fail(PB_INVALID_ARGUMENT, "segments must be appended in order");
    }
    if (!tr.has_t_min) {
      tr.has_t_min = true;
      tr.t_min = t_min;
    }
    tr.t_max.push_back(t_max[s]);
    tr.origin.push_back(origin[s]);
    tr.degree.push_back(degree[s]);
    tr.coefficients.insert(tr.coefficients.end(),
                           coefficients + 3 * kSegmentCoefficients * s,
                           coefficients + 3 * kSegmentCoefficients * (s + 1));
    // The slots above the degree are padding: -0.0, the neutral element that
    // EvaluateSegmentPadded relies on (pb_accel.cuh).
    double* const slots =
        tr.coefficients.data() + tr.coefficients.size() -
        3 * kSegmentCoefficients;
    for (int k = 3 * (degree[s] + 1); k < 3 * kSegmentCoefficients; ++k) {
      slots[k] = -0.0;
    }
  }
  e->segments_dirty = true;
  return PB_OK;
}

double pb_ephemeris_t_min(const pb_ephemeris_t* e) { return EphemerisTMin(e); }
double pb_ephemeris_t_max(const pb_ephemeris_t* e) { return EphemerisTMax(e); }

pb_status_t pb_ephemeris_evaluate_all_positions(pb_ephemeris_t* e, double t,
                                                double* positions) {
  if (e == nullptr || positions == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = nullptr;
  EphemerisView const view = e->View();
  PB_RETURN_IF_ERROR(e->stage_table.Reserve(view.stage_stride));
  PB_RETURN_IF_ERROR(BuildStageTable(e, &t, 1, e->stage_table.data, stream));
  std::vector<double> entry(view.stage_stride);
  PB_CUDA(cudaMemcpyAsync(entry.data(), e->stage_table.data,
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

pb_status_t pb_compute_gravitational_acceleration_on_massless_bodies_device(
    pb_ephemeris_t* e, double t, int64_t n, const double* q, double* a,
    int32_t* status_device, void* stream_pointer) {
  if (e == nullptr || n < 0 || (n > 0 && (q == nullptr || a == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = static_cast<cudaStream_t>(stream_pointer);
  EphemerisView const view = e->View();
  PB_RETURN_IF_ERROR(e->stage_table.Reserve(view.stage_stride));
  PB_RETURN_IF_ERROR(BuildStageTable(e, &t, 1, e->stage_table.data, stream));
  int32_t* const status =
      status_device != nullptr ? status_device : e->status_word.data;
  if (n > 0) {
    MasslessAccelerationKernel<true>
        <<<static_cast<unsigned>((n + kFlowThreads - 1) / kFlowThreads),
           kFlowThreads, 0, stream>>>(view, e->stage_table.data, n, q, a,
                                      status);
    PB_LAUNCHED();
  }
  return PB_OK;
}

pb_status_t pb_compute_gravitational_acceleration_on_massless_bodies(
    pb_ephemeris_t* e, double t, int64_t n, const double* q, double* a,
    pb_status_t* status) {
  if (e == nullptr || n < 0 || (n > 0 && (q == nullptr || a == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = nullptr;
  PB_RETURN_IF_ERROR(e->scratch_a.Upload(q, 3 * n, stream));
  PB_RETURN_IF_ERROR(e->scratch_b.Reserve(3 * n));
  PB_CUDA(cudaMemsetAsync(e->status_word.data, 0, sizeof(int32_t), stream));
  PB_RETURN_IF_ERROR(
      pb_compute_gravitational_acceleration_on_massless_bodies_device(
          e, t, n, e->scratch_a.data, e->scratch_b.data, e->status_word.data,
          stream));
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(a, e->scratch_b.data, 3 * n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
  }
  int32_t word = 0;
  PB_CUDA(cudaMemcpyAsync(&word, e->status_word.data, sizeof(int32_t),
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
  if (e == nullptr || n < 0 ||
      (n > 0 && (q == nullptr || potential == nullptr))) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = nullptr;
  EphemerisView const view = e->View();
  PB_RETURN_IF_ERROR(e->stage_table.Reserve(view.stage_stride));
  PB_RETURN_IF_ERROR(BuildStageTable(e, &t, 1, e->stage_table.data, stream));
  PB_RETURN_IF_ERROR(e->scratch_a.Upload(q, 3 * n, stream));
  PB_RETURN_IF_ERROR(e->scratch_b.Reserve(n));
  if (n > 0) {
    PotentialKernel<<<static_cast<unsigned>((n + kFlowThreads - 1) /
                                            kFlowThreads),
                      kFlowThreads, 0, stream>>>(view, e->stage_table.data, n,
                                                 e->scratch_a.data,
                                                 e->scratch_b.data);
    PB_LAUNCHED();
    PB_CUDA(cudaMemcpyAsync(potential, e->scratch_b.data, n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  return PB_OK;
}

pb_status_t pb_compute_gravitational_acceleration_on_massive_bodies(
    pb_ephemeris_t* e, double t, const double* positions,
    double* accelerations) {
  if (e == nullptr || accelerations == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = nullptr;
  EphemerisView const view = e->View();
  int const n = view.n_bodies;
  PB_RETURN_IF_ERROR(e->stage_table.Reserve(view.stage_stride));
  if (positions == nullptr) {
    PB_RETURN_IF_ERROR(BuildStageTable(e, &t, 1, e->stage_table.data, stream));
  } else {
    std::vector<double> internal(3 * n);
    for (int b = 0; b < n; ++b) {
      int const caller = e->model.order[b];
      for (int c = 0; c < 3; ++c) {
        internal[3 * b + c] = positions[3 * caller + c];
      }
    }
    PB_CUDA(cudaMemcpyAsync(e->stage_table.data, internal.data(),
                            internal.size() * sizeof(double),
                            cudaMemcpyHostToDevice, stream));
    PB_CUDA(cudaStreamSynchronize(stream));
    FramesKernel<<<(n + 63) / 64, 64, 0, stream>>>(view, t,
                                                  e->stage_table.data);
    PB_LAUNCHED();
  }
  PB_RETURN_IF_ERROR(e->scratch_b.Reserve(3 * n));
  MassiveAccelerationKernel<<<(n + 31) / 32, 32, 0, stream>>>(
      view, e->stage_table.data, e->scratch_b.data);
  PB_LAUNCHED();
  std::vector<double> internal(3 * n);
  PB_CUDA(cudaMemcpyAsync(internal.data(), e->scratch_b.data,
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

pb_status_t MultistepFlowDevice(pb_ephemeris_t* e,
                                const pb_fixed_step_flow_t* flow,
                                cudaStream_t stream);

// pb_flow_with_fixed_step_device for the SRKN methods, at most `max_steps`
// steps (the Starter of the multistep integrators runs it in groups of 16).
pb_status_t SrknFlowDevice(pb_ephemeris_t* e, const pb_fixed_step_flow_t* flow,
                           void* stream_pointer, long long const max_steps) {
  SrknTableau tableau;
  if (!MakeSrknTableau(flow->method, tableau)) {
    return Fail(PB_INVALID_ARGUMENT, "unsupported fixed-step method");
  }
  if (flow->step == 0 || flow->step != flow->step) {
    return Fail(PB_INVALID_ARGUMENT, "CHECK_NE(Time(), step)");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = static_cast<cudaStream_t>(stream_pointer);
  EphemerisView const view = e->View();
  double c[16];
  ComputeNodes(tableau, c);
  int const evaluations = tableau.stages - tableau.first_stage;
  double const h = flow->step;
  double const abs_h = h < 0 ? -h : h;
  DP t{flow->time[0], flow->time[1]};
  double const t_final = flow->t_final;
  bool const sampling = flow->samples != nullptr && flow->sample_every > 0;

  // Steps per launch: bounds the stage table (evaluations * stride doubles per
  // step).
  int const chunk_steps = 256;
  std::size_t const table_doubles =
      static_cast<std::size_t>(chunk_steps) * evaluations * view.stage_stride;
  PB_RETURN_IF_ERROR(e->stage_table.Reserve(table_doubles));
  PB_RETURN_IF_ERROR(
      e->pinned_times.Reserve(static_cast<std::size_t>(chunk_steps) *
                              evaluations));
  PB_CUDA(cudaMemsetAsync(e->status_word.data, 0, sizeof(int32_t), stream));
  PB_RETURN_IF_ERROR(UploadSegments(e, stream));
  PB_RETURN_IF_ERROR(UploadFlowConstants(e, t.value, flow->n, flow->state, stream));

  long long steps_done = 0;
  long long samples_done = 0;
  std::size_t launches = 0;
  std::size_t phases = 0;
  while (steps_done < max_steps &&
         abs_h <= std::abs((t_final - t.value) - t.error)) {
    int steps = 0;
    double* const times = e->pinned_times.data;
    // The pinned staging buffer is reused across chunks.
    PB_CUDA(cudaStreamSynchronize(stream));
    while (steps < chunk_steps && steps_done + steps < max_steps &&
           abs_h <= std::abs((t_final - t.value) - t.error)) {
      for (int st = tableau.first_stage; st < tableau.stages; ++st) {
        times[steps * evaluations + (st - tableau.first_stage)] =
            t.value + (t.error + c[st] * h);
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
    PB_RETURN_IF_ERROR(BuildStageTable(e, times, steps * evaluations,
                                       e->stage_table.data, stream));
    if (flow->n > 0) {
      while (e->timing_events.size() < 2 * (phases + 1)) {
        cudaEvent_t event;
        PB_CUDA(cudaEventCreate(&event));
        e->timing_events.push_back(event);
      }
      // The probes are cut into `groups` slices and the steps into sub-chunks;
      // slice g runs its sub-chunks in order on helper stream g.  Whole
      // launches of 10^5 probes leave 12 % of the SM slots idle in their last
      // wave (782 CTAs on 444 slots); short launches from several streams keep
      // the slots full, and no kernel ever waits on another one.
      int sm_count = 0;
      PB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount,
                                     e->device));
      int const variant = SrknVariant(flow->n, sm_count);
      int const threads =
          variant == 5 ? 256 : variant >= 2 ? 384 : kFlowThreads;
      long long const blocks = (flow->n + threads - 1) / threads;
      int const groups = static_cast<int>(
          std::min<long long>(FlowGroups(), std::max<long long>(1, blocks / 64)));
      int const sub_steps = groups > 1 ? FlowSubSteps() : steps;
      while (static_cast<int>(e->flow_streams.size()) < groups) {
        cudaStream_t helper;
        PB_CUDA(cudaStreamCreateWithFlags(&helper, cudaStreamNonBlocking));
        e->flow_streams.push_back(helper);
        cudaEvent_t event;
        PB_CUDA(cudaEventCreateWithFlags(&event, cudaEventDisableTiming));
        e->flow_events.push_back(event);
      }
      cudaEvent_t const fork = e->timing_events[2 * phases];
      cudaEvent_t const join = e->timing_events[2 * phases + 1];
      PB_CUDA(cudaEventRecord(fork, stream));
      long long const every = sampling ? flow->sample_every : 0;
      for (int g = 0; g < groups; ++g) {
        cudaStream_t const helper = groups > 1 ? e->flow_streams[g] : stream;
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
              e->stage_table.data +
              static_cast<std::size_t>(s0) * evaluations * view.stage_stride;
          double const* const times =
              e->stage_times.data + static_cast<std::size_t>(s0) * evaluations;
#define PB_LAUNCH_SRKN(kThreads, kMinBlocks)                                  \
  (flow->burns != nullptr ? SrknFlowKernel<kThreads, kMinBlocks, true>        \
                          : SrknFlowKernel<kThreads, kMinBlocks, false>)      \
      <<<grid, kThreads, 0, helper>>>(                                        \
          view, tableau, h, table, times, flow->burns, n_sub, flow->n,        \
          i_begin, i_end, flow->state, steps_done + s0, every,                \
          flow->max_samples, flow->samples, e->status_word.data)
          if (variant == 5) {
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
          ++launches;
        }
        if (groups > 1) {
          PB_CUDA(cudaEventRecord(e->flow_events[g], helper));
          PB_CUDA(cudaStreamWaitEvent(stream, e->flow_events[g], 0));
        }
      }
      PB_CUDA(cudaEventRecord(join, stream));
      ++phases;
    }
    steps_done += steps;
  }
  PB_CUDA(cudaMemcpyAsync(e->pinned_status.data, e->status_word.data,
                          sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  e->last_flow_kernel_ms = 0;
  e->last_flow_kernel_launches = static_cast<std::int64_t>(launches);
  // Device time of the flow-kernel phases (fork to join on the caller's
  // stream; the kernels of one phase overlap each other).
  for (std::size_t i = 0; i < phases; ++i) {
    float elapsed = 0;
    PB_CUDA(cudaEventElapsedTime(&elapsed, e->timing_events[2 * i],
                                 e->timing_events[2 * i + 1]));
    e->last_flow_kernel_ms += elapsed;
  }
  flow->time[0] = t.value;
  flow->time[1] = t.error;
  if (flow->n_steps != nullptr) {
    *flow->n_steps = steps_done;
  }
  if (flow->n_samples != nullptr) {
    *flow->n_samples = samples_done;
  }
  if (flow->status != nullptr) {
    // The only error the right-hand side reports is the collision
    // (ephemeris_body.hpp:383-384).
    *flow->status = e->pinned_status.data[0] != 0 ? PB_OUT_OF_RANGE : PB_OK;
  }
  return PB_OK;
}

}  // namespace

// Ephemeris::NewInstance for massless trajectories (ephemeris.hpp:181-194): the
// integrator instance with its ODE::State on the device and, for the
// multistep integrators, previous_steps_.
struct pb_fixed_step_instance {
  pb_ephemeris* ephemeris = nullptr;
  int32_t method = 0;
  bool is_multistep = false;
  MultistepMethod multistep;
  double step = 0;
  long long n = 0;
  DP time{0.0, 0.0};
  double* state = nullptr;  // 12 planes of n: owned_state or the caller's.
  DeviceBuffer<double> owned_state;
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
pb_status_t MasslessFillStep(pb_fixed_step_instance* s, int const slot,
                             cudaStream_t stream) {
  pb_ephemeris* const e = s->ephemeris;
  EphemerisView const view = e->View();
  PB_RETURN_IF_ERROR(e->stage_table.Reserve(view.stage_stride));
  double const t = s->time.value;
  PB_RETURN_IF_ERROR(BuildStageTable(e, &t, 1, e->stage_table.data, stream));
  if (s->n > 0) {
    MasslessFillStepKernel<<<static_cast<unsigned>((s->n + kFlowThreads - 1) /
                                                   kFlowThreads),
                             kFlowThreads, 0, stream>>>(
        view, e->stage_table.data, t, s->burns, s->n, s->state, s->History(),
        slot, e->status_word.data);
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
// symmetric_linear_multistep_integrator_body.hpp:69-149.
pb_status_t SolveMultistepInstance(pb_fixed_step_instance* s,
                                   double const t_final, InstanceSink* sink,
                                   cudaStream_t stream) {
  pb_ephemeris* const e = s->ephemeris;
  PB_RETURN_IF_ERROR(PrepareInstance(s));
  PB_RETURN_IF_ERROR(UploadSegments(e, stream));
  PB_RETURN_IF_ERROR(
      UploadFlowConstants(e, s->time.value, s->n, s->state, stream));
  EphemerisView const view = e->View();
  double const h = s->step;
  int const order = s->multistep.order;
  if (s->previous_steps < order) {
    if (s->previous_steps == 0) {
      PB_RETURN_IF_ERROR(MasslessFillStep(s, 0, stream));
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
      pb_status_t flow_status = PB_OK;
      pb_fixed_step_flow_t flow{};
      flow.method = PB_BLANES_MOAN_2002_SRKN_14A;
      flow.step = startup_step;
      flow.t_final = target;
      flow.n = s->n;
      flow.state = s->state;
      flow.time = time;
      flow.n_steps = &steps_taken;
      flow.status = &flow_status;
      flow.burns = s->burns;
      PB_RETURN_IF_ERROR(SrknFlowDevice(e, &flow, stream, wanted));
      s->time = {time[0], time[1]};
      s->startup_step_index += steps_taken;
      if (steps_taken == wanted) {
        PB_RETURN_IF_ERROR(MasslessFillStep(s, s->previous_steps, stream));
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
  PB_CUDA(cudaMemsetAsync(e->status_word.data, 0, sizeof(int32_t), stream));
  // Multistep steps, fused `chunk_steps` at a time.
  int const chunk_steps = 256;
  PB_RETURN_IF_ERROR(e->stage_table.Reserve(
      static_cast<std::size_t>(chunk_steps) * view.stage_stride));
  PB_RETURN_IF_ERROR(e->pinned_times.Reserve(chunk_steps));
  std::size_t phases = 0;
  e->last_flow_kernel_launches = 0;
  while (h <= (t_final - s->time.value) - s->time.error) {
    int steps = 0;
    double* const times = e->pinned_times.data;
    PB_CUDA(cudaStreamSynchronize(stream));
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
        BuildStageTable(e, times, steps, e->stage_table.data, stream));
    if (s->n > 0) {
      while (e->timing_events.size() < 2 * (phases + 1)) {
        cudaEvent_t event;
        PB_CUDA(cudaEventCreate(&event));
        e->timing_events.push_back(event);
      }
      // Probe slices x step sub-chunks on helper streams, as in the SRKN flow
      // (no last-wave tail: 1.91 -> 2.04 10^9 probe-steps/s on 10^5 probes).
      int sm_count = 0;
      PB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount,
                                     e->device));
      bool const wide = MultistepWide(s->n, sm_count);
      int const threads = wide ? 256 : kFlowThreads;
      long long const blocks = (s->n + threads - 1) / threads;
      int const groups = static_cast<int>(
          std::min<long long>(FlowGroups(), std::max<long long>(1, blocks / 64)));
      int const sub_steps = groups > 1 ? FlowSubSteps() : steps;
      while (static_cast<int>(e->flow_streams.size()) < groups) {
        cudaStream_t helper;
        PB_CUDA(cudaStreamCreateWithFlags(&helper, cudaStreamNonBlocking));
        e->flow_streams.push_back(helper);
        cudaEvent_t event;
        PB_CUDA(cudaEventCreateWithFlags(&event, cudaEventDisableTiming));
        e->flow_events.push_back(event);
      }
      cudaEvent_t const fork = e->timing_events[2 * phases];
      PB_CUDA(cudaEventRecord(fork, stream));
      bool const sampling = sink->samples != nullptr && sink->sample_every > 0;
      for (int g = 0; g < groups; ++g) {
        cudaStream_t const helper = groups > 1 ? e->flow_streams[g] : stream;
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
          e->stage_table.data +                                               \
              static_cast<std::size_t>(s0) * view.stage_stride,               \
          e->stage_times.data + s0, s->burns, n_sub, s->n, i_begin, i_end,    \
          s->state, s->History(), (s->oldest + s0) % order,                   \
          appended_before + s0, sampling ? sink->sample_every : 0,            \
          sink->max_samples, sink->samples, e->status_word.data)
          if (order == 8) {
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
          ++e->last_flow_kernel_launches;
        }
        if (groups > 1) {
          PB_CUDA(cudaEventRecord(e->flow_events[g], helper));
          PB_CUDA(cudaStreamWaitEvent(stream, e->flow_events[g], 0));
        }
      }
      PB_CUDA(cudaEventRecord(e->timing_events[2 * phases + 1], stream));
      ++phases;
    }
    s->oldest = (s->oldest + steps) % order;
    s->time = t;
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  e->last_flow_kernel_ms = 0;
  for (std::size_t i = 0; i < phases; ++i) {
    float elapsed = 0;
    PB_CUDA(cudaEventElapsedTime(&elapsed, e->timing_events[2 * i],
                                 e->timing_events[2 * i + 1]));
    e->last_flow_kernel_ms += elapsed;
  }
  return PB_OK;
}

// Instance::Solve(t_final) for either family; the outs of `flow` are filled.
pb_status_t SolveInstance(pb_fixed_step_instance* s,
                          const pb_fixed_step_flow_t* flow,
                          cudaStream_t stream) {
  pb_ephemeris* const e = s->ephemeris;
  if (!s->is_multistep) {
    pb_fixed_step_flow_t srkn = *flow;
    srkn.method = s->method;
    srkn.step = s->step;
    srkn.n = s->n;
    srkn.state = s->state;
    srkn.burns = s->burns;
    double time[2] = {s->time.value, s->time.error};
    srkn.time = time;
    PB_RETURN_IF_ERROR(SrknFlowDevice(e, &srkn, stream,
                                      std::numeric_limits<long long>::max()));
    s->time = {time[0], time[1]};
    return PB_OK;
  }
  PB_CUDA(cudaMemsetAsync(e->status_word.data, 0, sizeof(int32_t), stream));
  InstanceSink sink{flow->sample_every, flow->max_samples, flow->samples,
                    flow->sample_times};
  PB_RETURN_IF_ERROR(SolveMultistepInstance(s, flow->t_final, &sink, stream));
  PB_CUDA(cudaMemcpyAsync(e->pinned_status.data, e->status_word.data,
                          sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  PB_CUDA(cudaStreamSynchronize(stream));
  if (flow->n_steps != nullptr) {
    *flow->n_steps = sink.appended;
  }
  if (flow->n_samples != nullptr) {
    *flow->n_samples = sink.recorded;
  }
  if (flow->status != nullptr) {
    *flow->status = e->pinned_status.data[0] != 0 ? PB_OUT_OF_RANGE : PB_OK;
  }
  return PB_OK;
}

pb_status_t MultistepFlowDevice(pb_ephemeris_t* e,
                                const pb_fixed_step_flow_t* flow,
                                cudaStream_t stream) {
  if (!(flow->step > 0)) {
    return Fail(PB_INVALID_ARGUMENT,
                "the multistep massless flow integrates forward");
  }
  PB_CUDA(cudaSetDevice(e->device));
  pb_fixed_step_instance instance;
  instance.ephemeris = e;
  instance.method = flow->method;
  instance.is_multistep = true;
  MakeMultistepMethod(flow->method, instance.multistep);
  instance.step = flow->step;
  instance.n = flow->n;
  instance.time = {flow->time[0], flow->time[1]};
  instance.state = flow->state;
  instance.burns = flow->burns;
  PB_RETURN_IF_ERROR(e->pinned_status.Reserve(1));
  PB_RETURN_IF_ERROR(e->status_word.Reserve(1));
  PB_RETURN_IF_ERROR(SolveInstance(&instance, flow, stream));
  flow->time[0] = instance.time.value;
  flow->time[1] = instance.time.error;
  return PB_OK;
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
  *instance = s;
  return PB_OK;
}

void pb_fixed_step_instance_destroy(pb_fixed_step_instance_t* instance) {
  if (instance != nullptr) {
    cudaSetDevice(instance->ephemeris->device);
    delete instance;
  }
}

pb_status_t pb_fixed_step_instance_set_intrinsic_accelerations(
    pb_fixed_step_instance_t* s, const double* burns) {
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null instance");
  }
  PB_CUDA(cudaSetDevice(s->ephemeris->device));
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

pb_status_t pb_fixed_step_instance_flow(pb_fixed_step_instance_t* s,
                                        double t_final, int64_t sample_every,
                                        int64_t max_samples, double* samples,
                                        double* sample_times,
                                        int64_t* n_steps, int64_t* n_samples,
                                        pb_status_t* flow_status) {
  std::lock_guard<std::recursive_mutex> const flow_lock(FlowMutex());
  if (s == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null instance");
  }
  pb_ephemeris* const e = s->ephemeris;
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = nullptr;
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
        e->scratch_c.Reserve(static_cast<std::size_t>(max_samples) * 6 * n));
    flow.samples = e->scratch_c.data;
  }
  PB_RETURN_IF_ERROR(e->pinned_status.Reserve(1));
  PB_RETURN_IF_ERROR(e->status_word.Reserve(1));
  PB_RETURN_IF_ERROR(SolveInstance(s, &flow, stream));
  if (sampling && recorded > 0 && n > 0) {
    PB_CUDA(cudaMemcpyAsync(
        samples, e->scratch_c.data,
        static_cast<std::size_t>(recorded) * 6 * n * sizeof(double),
        cudaMemcpyDeviceToHost, stream));
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  if (n_samples != nullptr) {
    *n_samples = recorded;
  }
  return PB_OK;
}

pb_status_t pb_fixed_step_instance_flow_to_downsampler(
    pb_fixed_step_instance_t* s, double t_final,
    pb_downsampler_t* downsampler, int64_t* n_steps,
    pb_status_t* flow_status) {
  std::lock_guard<std::recursive_mutex> const flow_lock(FlowMutex());
  if (s == nullptr || downsampler == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  pb_ephemeris* const e = s->ephemeris;
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = nullptr;
  std::size_t const n = static_cast<std::size_t>(s->n);
  // States per chunk: at most 64 MiB of samples on the device.
  long long const chunk = std::max<long long>(
      1, std::min<long long>(256, (64LL << 20) / (48 * std::max<long long>(
                                                           1, s->n))));
  // Two spare slots: the time bound below admits one more state when the
  // instance sits between two steps of a multistep startup.
  long long const slots = chunk + 2;
  PB_RETURN_IF_ERROR(
      e->scratch_c.Reserve(static_cast<std::size_t>(slots) * 6 * n + 1));
  PB_RETURN_IF_ERROR(e->pinned_status.Reserve(1));
  PB_RETURN_IF_ERROR(e->status_word.Reserve(1));
  std::vector<double> sample_times(slots);
  int64_t total_steps = 0;
  pb_status_t overall = PB_OK;
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
    flow.samples = e->scratch_c.data;
    flow.sample_times = sample_times.data();
    int64_t steps = 0, recorded = 0;
    pb_status_t status = PB_OK;
    flow.n_steps = &steps;
    flow.n_samples = &recorded;
    flow.status = &status;
    DP const before = s->time;
    PB_RETURN_IF_ERROR(SolveInstance(s, &flow, stream));
    if (recorded != steps) {
      return Fail(PB_INTERNAL, "more states appended than the chunk holds");
    }
    PB_RETURN_IF_ERROR(pb_downsampler_append_device(
        downsampler, recorded, sample_times.data(), e->scratch_c.data,
        stream));
    total_steps += steps;
    if (overall == PB_OK) {
      overall = status;
    }
    if (s->time.value == before.value && s->time.error == before.error) {
      break;  // No progress: t_final is closer than one step.
    }
  }
  if (n_steps != nullptr) {
    *n_steps = total_steps;
  }
  if (flow_status != nullptr) {
    *flow_status = overall;
  }
  return PB_OK;
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
  if (s == nullptr || checkpoint == nullptr) {
    return Fail(PB_INVALID_ARGUMENT, "null argument");
  }
  std::lock_guard<std::recursive_mutex> const flow_lock(FlowMutex());
  PB_CUDA(cudaSetDevice(s->ephemeris->device));
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
  PB_CUDA(cudaSetDevice(s->ephemeris->device));
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

pb_status_t pb_flow_with_fixed_step_device(pb_ephemeris_t* e,
                                           const pb_fixed_step_flow_t* flow,
                                           void* stream_pointer) {
  std::lock_guard<std::recursive_mutex> const flow_lock(FlowMutex());
  if (e == nullptr || flow == nullptr || flow->n < 0 || flow->time == nullptr ||
      (flow->n > 0 && flow->state == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  if (flow->step == 0 || flow->step != flow->step) {
    return Fail(PB_INVALID_ARGUMENT, "CHECK_NE(Time(), step)");
  }
  MultistepMethod multistep;
  if (MakeMultistepMethod(flow->method, multistep)) {
    // NewInstance + FlowWithFixedStep in one call: the instance (and the
    // history it would keep) lives for this call only.
    return MultistepFlowDevice(e, flow,
                               static_cast<cudaStream_t>(stream_pointer));
  }
  return SrknFlowDevice(e, flow, stream_pointer,
                        std::numeric_limits<long long>::max());
}

pb_status_t pb_flow_with_fixed_step(pb_ephemeris_t* e,
                                    const pb_fixed_step_flow_t* flow) {
  std::lock_guard<std::recursive_mutex> const flow_lock(FlowMutex());
  if (e == nullptr || flow == nullptr || flow->n < 0 ||
      (flow->n > 0 && flow->state == nullptr)) {
    return Fail(PB_INVALID_ARGUMENT, "bad argument");
  }
  PB_CUDA(cudaSetDevice(e->device));
  cudaStream_t const stream = nullptr;
  std::size_t const n = static_cast<std::size_t>(flow->n);
  PB_RETURN_IF_ERROR(e->scratch_a.Upload(flow->state, 12 * n, stream));
  pb_fixed_step_flow_t device_flow = *flow;
  device_flow.state = e->scratch_a.data;
  if (flow->burns != nullptr) {
    PB_RETURN_IF_ERROR(
        e->burns_scratch.Upload(flow->burns, PB_BURN_PLANES * n, stream));
    device_flow.burns = e->burns_scratch.data;
  }
  bool const sampling = flow->samples != nullptr && flow->sample_every > 0 &&
                        flow->max_samples > 0;
  if (sampling) {
    PB_RETURN_IF_ERROR(e->scratch_c.Reserve(
        static_cast<std::size_t>(flow->max_samples) * 6 * n));
    device_flow.samples = e->scratch_c.data;
  } else {
    device_flow.samples = nullptr;
  }
  int64_t n_samples = 0;
  device_flow.n_samples = &n_samples;
  PB_RETURN_IF_ERROR(pb_flow_with_fixed_step_device(e, &device_flow, stream));
  if (n > 0) {
    PB_CUDA(cudaMemcpyAsync(flow->state, e->scratch_a.data,
                            12 * n * sizeof(double), cudaMemcpyDeviceToHost,
                            stream));
    if (sampling && n_samples > 0) {
      PB_CUDA(cudaMemcpyAsync(flow->samples, e->scratch_c.data,
                              static_cast<std::size_t>(n_samples) * 6 * n *
                                  sizeof(double),
                              cudaMemcpyDeviceToHost, stream));
    }
  }
  PB_CUDA(cudaStreamSynchronize(stream));
  if (flow->n_samples != nullptr) {
    *flow->n_samples = n_samples;
  }
  return PB_OK;
}

double pb_ephemeris_last_flow_kernel_ms(const pb_ephemeris_t* e,
                                        int64_t* launches) {
  if (launches != nullptr) {
    *launches = e->last_flow_kernel_launches;
  }
  return e->last_flow_kernel_ms;
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

}  // extern "C"
