// Adaptive massless flow: EmbeddedExplicitRungeKuttaNyströmIntegrator::
// Instance::Solve (integrators/embedded_explicit_runge_kutta_nyström_
// integrator_body.hpp:44-258) with DormandالمكاوىPrince1986RKN434FM
// (integrators/methods.hpp:574-597), driven like
// Ephemeris::FlowODEWithAdaptiveStep (physics/ephemeris_body.hpp:1624-1717):
// one independent trajectory per lane; step size, error control, rejection
// loop and FSAL state all in registers.  The lanes are persistent: a lane that
// finishes its trajectory takes the next one from a queue ordered by decreasing
// estimated cost, so a warp never idles behind its longest trajectory.
#pragma once

#include <cuda_runtime.h>

#include "pb_accel.cuh"
#include "pb_downsampling.cuh"
#include "pb_kernels_jerk.cuh"
#include "pb_model.h"

namespace pb {

constexpr int kAdaptiveThreads = 128;

// Bodies per iteration of the right-hand side's loop (see
// FastAccelerationAtTime).
#if !defined(PB_ADAPTIVE_PAIRS)
#define PB_ADAPTIVE_PAIRS 0
#endif
// Whether the deep body's tesseral harmonics take the unrolled evaluation.
#if !defined(PB_ADAPTIVE_DEEP_UNROLLED)
#define PB_ADAPTIVE_DEEP_UNROLLED 1
#endif

static __constant__ BodyScalars c_adaptive_bodies;

// The oblate bodies (the first of the internal order) with what their
// harmonics up to degree kAdaptiveTableDegree need, in the constant bank: the
// damped J2 of the Sun at every evaluation, the Earth beyond 23 000 km.  Through
// the pointers of the EphemerisView the same evaluation is a chain of
// dependent global loads (the body record, the threshold search, the
// coefficients) that two warps per scheduler do not hide.
constexpr int kAdaptiveTableBodies = 16;
constexpr int kAdaptiveTableDegree = 3;
constexpr int kAdaptiveTableSize =
    (kAdaptiveTableDegree + 1) * (kAdaptiveTableDegree + 2) / 2;
// The thresholds up to degree kAdaptiveTableDegree + 1 tell whether the degree
// exceeds kAdaptiveTableDegree.
constexpr int kAdaptiveTableDampings = kAdaptiveTableDegree + 2;

struct AdaptiveBodyTables {
  DeviceBody data;
  Damping degree_damping[kAdaptiveTableDampings];
  int32_t n_damping;  // min(n_damping, kAdaptiveTableDampings).
  int32_t pad;
  double cos[kAdaptiveTableSize];
  double sin[kAdaptiveTableSize];
};

struct AdaptiveConstants {
  int32_t n_tables;  // Bodies [0, n_tables) have an entry.
  int32_t pad;
  double normalization[kAdaptiveTableSize];
  AdaptiveBodyTables bodies[kAdaptiveTableBodies];
};

static __constant__ AdaptiveConstants c_adaptive_tables;

// The *deep* body: the one oblate body of degree 4 .. kMaxUnrolledDegree that
// most trajectories of the call are closest to (the Earth of a batch of Earth
// orbiters), with all its thresholds and coefficients in the constant bank.
// The rolled evaluation walks the tables with run-time indices, which the
// constant bank serves (LDC) without a round trip through the L1 that the
// evaluation's own thread-local scratch keeps evicting.
constexpr int kAdaptiveDeepSize =
    (kMaxUnrolledDegree + 1) * (kMaxUnrolledDegree + 2) / 2;

struct AdaptiveDeepBody {
  int32_t body;       // Internal index < kAdaptiveTableBodies, or -1.
  int32_t n_damping;  // The body's own.
  Damping degree_damping[kMaxUnrolledDegree + 1];
  double cos[kAdaptiveDeepSize];
  double sin[kAdaptiveDeepSize];
  double normalization[kAdaptiveDeepSize];
};

static __constant__ AdaptiveDeepBody c_adaptive_deep;

struct AdaptiveDeepTables {
  __device__ __forceinline__ double Cos(int const k) const {
    return c_adaptive_deep.cos[k];
  }
  __device__ __forceinline__ double Sin(int const k) const {
    return c_adaptive_deep.sin[k];
  }
  __device__ __forceinline__ double Normalization(int const k) const {
    return c_adaptive_deep.normalization[k];
  }
  __device__ __forceinline__ Damping const& DegreeDamping(int const n) const {
    return c_adaptive_deep.degree_damping[n];
  }
  __device__ __forceinline__ Damping const& SectoralDamping() const {
    return c_adaptive_tables.bodies[c_adaptive_deep.body].data.sectoral;
  }
};

struct AdaptiveTables {
  int b;
  __device__ __forceinline__ double Cos(int const k) const {
    return c_adaptive_tables.bodies[b].cos[k];
  }
  __device__ __forceinline__ double Sin(int const k) const {
    return c_adaptive_tables.bodies[b].sin[k];
  }
  __device__ __forceinline__ double Normalization(int const k) const {
    return c_adaptive_tables.normalization[k];
  }
  __device__ __forceinline__ Damping const& DegreeDamping(int const n) const {
    return c_adaptive_tables.bodies[b].degree_damping[n];
  }
  __device__ __forceinline__ Damping const& SectoralDamping() const {
    return c_adaptive_tables.bodies[b].data.sectoral;
  }
};

// Right-hand side at a per-probe time: every massive body's polynomial is
// evaluated here (the stage-table trick of the fixed-step kernel needs a time
// shared by all probes).  `hint` caches the segment index relative to the
// body's first segment; all bodies of an ephemeris are appended in lock step,
// so one hint serves all of them and is re-validated per body.
template<bool kUnrolled>
__device__ __forceinline__ int32_t AccelerationOnMasslessBodyAtTime(
    EphemerisView const& e, double const t, Vec3 const& q, Vec3& acceleration,
    int& hint) {
  acceleration = {0.0, 0.0, 0.0};
  int32_t error = 0;
  for (int b1 = 0; b1 < e.n_bodies; ++b1) {
    DeviceBody const& body1 = e.bodies[b1];
    double const mu1 = body1.mu;
    int const begin = e.segment_offset[b1];
    int const end = e.segment_end[b1];
    int segment = begin + hint;
    // First segment with t <= t_max (continuous_trajectory_body.hpp:860-885).
    if (!(segment < end && t <= e.segment_t_max[segment] &&
          (segment == begin || e.segment_t_max[segment - 1] < t))) {
      segment = FindSegment(e.segment_t_max, begin, end, t);
      hint = segment - begin;
    }
    Vec3 const position1 = EvaluateSegment(e, segment, t);
    Vec3 const dq = position1 - q;
    double const dq2 = Norm2(dq);
    double const dq_norm = sqrt(dq2);
    error |= dq_norm > body1.collision_radius ? 0 : 11;
    double const one_over_dq3 = dq_norm / (dq2 * dq2);
    double const mu1_over_dq3 = mu1 * one_over_dq3;
    acceleration += dq * mu1_over_dq3;
    if (b1 < e.n_oblate) {
      GeopotentialBody const g = MakeGeopotentialBody(e, b1);
      FrameSource const frame{nullptr, t};
      Vec3 const effect = GeneralSphericalHarmonicsAcceleration<kUnrolled>(
          g, frame, -dq, dq_norm, dq2, one_over_dq3);
      acceleration += mu1 * effect;
    }
  }
  return error;
}

struct AdaptiveParameters {
  double t_requested;  // The caller's t.
  double t_final;      // min(t, t_max of the ephemeris).
  double t_min;        // Of the ephemeris.
  double length_integration_tolerance;
  double speed_integration_tolerance;
  long long max_steps;
  long long step_log_capacity;
};

// std::max / std::min of the reference, NaN behaviour included.
__device__ __forceinline__ double StdMax(double const a, double const b) {
  return (a < b) ? b : a;
}
__device__ __forceinline__ double StdMin(double const a, double const b) {
  return (b < a) ? b : a;
}

// The reference formulation with the IEEE operators, out of line: taken when
// a distance leaves the range of the branch-free square root / division.
__device__ __noinline__ int32_t SlowAccelerationAtTime(EphemerisView const& e,
                                                       double const t,
                                                       Vec3 const& q,
                                                       Vec3& acceleration) {
  int hint = 0;
  return AccelerationOnMasslessBodyAtTime<false>(e, t, q, acceleration, hint);
}

// The code of the right-hand side is kept SMALL (run-time loops, one tree for
// all polynomial degrees): the lanes of the resident warps sit at different
// bodies, degrees and step-control branches, and a multi-hundred-kB unrolled
// body (the fixed-step kernel's choice, where all warps run the same code)
// measured 80 % of the issue slots lost to instruction fetch here.
__device__ __noinline__ Vec3 AdaptiveHarmonics(EphemerisView const& e,
                                               int const b1, double const t,
                                               Vec3 const& dq,
                                               double const dq_norm,
                                               double const dq2,
                                               double const one_over_dq3) {
  if (b1 < c_adaptive_tables.n_tables && dq_norm == dq_norm) {
    AdaptiveBodyTables const& tables = c_adaptive_tables.bodies[b1];
    bool const deep = b1 == c_adaptive_deep.body;
    // LimitingDegree over the first thresholds is min(LimitingDegree,
    // kAdaptiveTableDampings): exact whenever it says degree <= 3.  The deep
    // body has all its thresholds here.
    int const max_degree =
        (deep ? LimitingDegree(c_adaptive_deep.degree_damping,
                               c_adaptive_deep.n_damping, dq_norm)
              : LimitingDegree(tables.degree_damping, tables.n_damping,
                               dq_norm)) -
        1;
    if (deep || max_degree <= kAdaptiveTableDegree) {
      // The operations of GeneralSphericalHarmonicsAccelerationRolled (the
      // unrolled and the rolled evaluation of a degree have the same ones).
      if (max_degree == 1) {
        return {0.0, 0.0, 0.0};
      }
      DeviceBody const& body = tables.data;
      bool const is_zonal =
          body.is_zonal || dq_norm > body.sectoral.outer_threshold;
      Vec3 x_hat, y_hat;
      if (is_zonal) {
        x_hat = body.equatorial;
        y_hat = body.biequatorial;
      } else {
        SurfaceFrameAxes(body, t, x_hat, y_hat);
      }
      GeopotentialSetup s;
      InitializeSetup(body, x_hat, y_hat, -dq, dq_norm, dq2, one_over_dq3, s);
#if PB_ADAPTIVE_DEEP_UNROLLED
      // The deep body: the unrolled evaluation of the full degree, every lane
      // leaving at its own degree (the lanes of a warp near the body walk ONE
      // body of code whatever their degrees).  Coefficients are constant-bank
      // operands, the Legendre values live in registers.
      if (deep) {
        FlaggedScratch<kMaxUnrolledDegree> w;
        Vec3 const acceleration =
            is_zonal
                ? GeopotentialUnrolledAcceleration<kMaxUnrolledDegree, true,
                                                   true>(AdaptiveDeepTables{},
                                                         s, w, max_degree)
                : GeopotentialUnrolledAcceleration<kMaxUnrolledDegree, false,
                                                   true>(AdaptiveDeepTables{},
                                                         s, w, max_degree);
        if (!w.Failed()) {
          return acceleration;
        }
      }
#endif
      if (max_degree > kAdaptiveTableDegree) {
        return GeopotentialRolledAcceleration<kMaxUnrolledDegree + 1>(
            AdaptiveDeepTables{}, max_degree, is_zonal, s);
      }
      AdaptiveTables const g{b1};
      return is_zonal
                 ? GeopotentialUnrolledDispatch<true, AdaptiveTables,
                                                kAdaptiveTableDegree>(
                       g, max_degree, s)
                 : GeopotentialUnrolledDispatch<false, AdaptiveTables,
                                                kAdaptiveTableDegree>(
                       g, max_degree, s);
    }
  }
  GeopotentialBody const g = MakeGeopotentialBody(e, b1);
  return GeneralSphericalHarmonicsAccelerationRolled(g, t, -dq, dq_norm, dq2,
                                                     one_over_dq3);
}

// AccelerationOnMasslessBodyAtTime with the padded Estrin tree, the
// branch-free square root and division (pb_math.cuh) and the early exit for
// oblate bodies beyond their degree-2 outer threshold: the same arithmetic,
// body by body in the same order.  `hint` is the index of the segment of the
// previous evaluation (relative to the body's first segment).
template<bool kTwoPhase>
__device__ __forceinline__ int32_t FastAccelerationAtTime(
    EphemerisView const& e, double const t, Vec3 const& q, Vec3& acceleration,
    int& hint) {
  double ax = 0.0, ay = 0.0, az = 0.0;
  int32_t error = 0;
  bool in_range = true;
  int const n_bodies = e.n_bodies;
  int const n_oblate = e.n_oblate;
  bool const uniform = e.uniform_segments != 0;
  if (uniform) {
    // One lookup for all bodies: first segment with t <= t_max
    // (continuous_trajectory_body.hpp:860-885).
    int const end = __ldg(e.segment_end);
    if (!(hint < end && t <= __ldg(e.segment_t_max + hint) &&
          (hint == 0 || __ldg(e.segment_t_max + hint - 1) < t))) {
      hint = FindSegment(e.segment_t_max, 0, end, t);
    }
  }
  auto const position_of = [&](int const b1) {
    // The segments of body b1 start at b1 * segment_capacity: no load between
    // the hint and the addresses of the coefficients.
    int const begin = b1 * e.segment_capacity;
    int segment = begin + hint;
    if (!uniform) {
      int const end = __ldg(e.segment_end + b1);
      if (!(segment < end && t <= __ldg(e.segment_t_max + segment) &&
            (segment == begin || __ldg(e.segment_t_max + segment - 1) < t))) {
        segment = FindSegment(e.segment_t_max, begin, end, t);
        hint = segment - begin;
      }
    }
    // (Prefetching the next body's coefficients into the L1 while this one is
    // evaluated changed nothing.  Serving them from a window of the last 3 to
    // 8 segments copied to shared memory -- the lanes of a warp sit in half a
    // dozen different segments, ncu attributes 18 % of the warp time to these
    // loads -- was 3 to 9 % SLOWER, the more so the larger the window: what it
    // takes from the L1 is missed by the thread-local scratch of the
    // harmonics.  Skipping the loads and selects of the leaves that are
    // padding in every segment of a body -- a per-body bound passed with the
    // kernel's parameters -- was 14 % slower: the loads then wait for the
    // bound.)
    return EvaluateSegmentPadded(e, segment, t);
  };
  // kTwoPhase: the polynomials of all the bodies first, in a loop of their own
  // whose coefficient loads (27 16-byte loads per body) the compiler pipelines
  // across bodies, then the forces from a thread-local table of positions.
  // Measured SLOWER than the fused loop on the 125 000-probe workload
  // (5.05 against 5.56 10^8 evaluations/s: the table goes through the same
  // L1 as the scratch of the harmonics) and not instantiated; kept as the
  // record of the experiment.
  double body_positions[kTwoPhase ? 3 * kMaxConstantBodies : 3];
  if constexpr (kTwoPhase) {
#pragma unroll 2
    for (int b1 = 0; b1 < n_bodies; ++b1) {
      Vec3 const p = position_of(b1);
      body_positions[3 * b1] = p.x;
      body_positions[3 * b1 + 1] = p.y;
      body_positions[3 * b1 + 2] = p.z;
    }
  }
  // The geometry of a body -- its polynomials, the distance, 1 / distance^3 --
  // does not depend on the running sum.  PB_ADAPTIVE_PAIRS evaluates it for
  // two spherical bodies at a time, so that the coefficient loads and the
  // square root and division chains of the second overlap those of the first:
  // measured no faster (9.6-9.9 against 10.0 10^8 evaluations/s; with the
  // oblate bodies paired as well, 4 % slower) and off.  The terms are added
  // body by body in the reference's order either way.
  struct Geometry {
    double dx, dy, dz, dq2, dq_norm, one_over_dq3;
  };
  auto const geometry_of = [&](int const b1) {
    Vec3 position1;
    if constexpr (kTwoPhase) {
      position1 = {body_positions[3 * b1], body_positions[3 * b1 + 1],
                   body_positions[3 * b1 + 2]};
    } else {
      position1 = position_of(b1);
    }
    Geometry g;
    g.dx = position1.x - q.x;
    g.dy = position1.y - q.y;
    g.dz = position1.z - q.z;
    g.dq2 = g.dx * g.dx + g.dy * g.dy + g.dz * g.dz;
    in_range = in_range && InFastRange(g.dq2);
    g.dq_norm = FastSqrt(g.dq2);
    g.one_over_dq3 = FastDivide(g.dq_norm, g.dq2 * g.dq2);
    return g;
  };
  auto const add_point_mass = [&](int const b1, Geometry const& g) {
    double const mu1 = c_adaptive_bodies.mu[b1];
    error |= g.dq_norm > c_adaptive_bodies.collision_radius[b1] ? 0 : 11;
    double const mu1_over_dq3 = mu1 * g.one_over_dq3;
    ax += g.dx * mu1_over_dq3;
    ay += g.dy * mu1_over_dq3;
    az += g.dz * mu1_over_dq3;
    return mu1;
  };
  // The oblate bodies come first in the internal order; the spherical ones
  // have a loop of their own without the call and the reconvergence point of
  // the harmonics (ncu, r04a: 4 % of the warp time sat on the instruction
  // after that point and 2 % on the reload of the loop bound after the call,
  // for all 34 bodies).
  int b1 = 0;
#pragma unroll 1
  for (; b1 < n_oblate; ++b1) {
    Geometry const g = geometry_of(b1);
    double const mu1 = add_point_mass(b1, g);
    // A NaN distance fails the comparison and takes the full path.
    if (!(g.dq_norm >= c_adaptive_bodies.outer_threshold_degree_2[b1])) {
      Vec3 const effect =
          AdaptiveHarmonics(e, b1, t, Vec3{g.dx, g.dy, g.dz}, g.dq_norm, g.dq2,
                            g.one_over_dq3);
      ax += mu1 * effect.x;
      ay += mu1 * effect.y;
      az += mu1 * effect.z;
    } else {
      // The reference adds mu1 * (zero vector) here.
      double const zero = mu1 * 0.0;
      ax += zero;
      ay += zero;
      az += zero;
    }
  }
#if PB_ADAPTIVE_PAIRS
#pragma unroll 1
  for (; b1 + 1 < n_bodies; b1 += 2) {
    Geometry const first = geometry_of(b1);
    Geometry const second = geometry_of(b1 + 1);
    add_point_mass(b1, first);
    add_point_mass(b1 + 1, second);
  }
#endif
#pragma unroll 1
  for (; b1 < n_bodies; ++b1) {
    add_point_mass(b1, geometry_of(b1));
  }
  if (!in_range) {
    return SlowAccelerationAtTime(e, t, q, acceleration);
  }
  acceleration = {ax, ay, az};
  return error;
}

// append_state into the trajectory sink, out of line (a few hundred
// instructions that run once per accepted step).
__device__ __noinline__ void AppendToSink(DownsamplerView const& sink,
                                          long long const i, double const t,
                                          Vec3 const& q, Vec3 const& v) {
  DownsamplerAppend(sink, i, t, q, v);
}

// Scheduling key of a trajectory: an estimate of its number of steps, from the
// osculating orbit around the body with the strongest tide at the initial
// state.  With a 4th-order error estimate the step scales like
// (tolerance / a)^(1/4) / n (n the mean motion), so the cost over a span is
// proportional to span * n * a^(1/4).  Only the order of execution depends on
// it, never a result.  votes[b]: how many trajectories have body b as their
// strongest tide (the host chooses the deep body from them).
static __global__ void AdaptiveCostKernel(EphemerisView const e,
                                          double const t_final,
                                          long long const n,
                                          double const* __restrict__ state,
                                          double const* __restrict__ time,
                                          float* __restrict__ cost,
                                          int* __restrict__ index,
                                          int32_t* __restrict__ votes) {
  __shared__ int32_t block_votes[kMaxConstantBodies];
  for (int b = threadIdx.x; b < kMaxConstantBodies; b += blockDim.x) {
    block_votes[b] = 0;
  }
  __syncthreads();
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    double const t = time[i];
    Vec3 const q{state[i], state[n + i], state[2 * n + i]};
    Vec3 const v{state[6 * n + i], state[7 * n + i], state[8 * n + i]};
    double best_tide = -1.0;
    double best_mu = 0.0;
    int best_body = 0;
    Vec3 best_dq{0.0, 0.0, 0.0}, best_dv{0.0, 0.0, 0.0};
    for (int b = 0; b < e.n_bodies; ++b) {
      int const segment = FindSegment(e.segment_t_max, e.segment_offset[b],
                                      e.segment_end[b], t);
      Vec3 const p0 = EvaluateSegment(e, segment, t);
      Vec3 const dq = q - p0;
      double const r2 = Norm2(dq);
      double const mu = e.bodies[b].mu;
      double const tide = mu / (r2 * sqrt(r2));
      if (tide > best_tide) {
        Vec3 const p1 = EvaluateSegment(e, segment, t + 1.0);
        best_tide = tide;
        best_mu = mu;
        best_body = b;
        best_dq = dq;
        best_dv = v - (p1 - p0);
      }
    }
    double const r = sqrt(Norm2(best_dq));
    double const energy = 0.5 * Norm2(best_dv) - best_mu / r;
    double a = r;
    if (energy < 0.0) {
      a = -0.5 * best_mu / energy;
    }
    double const mean_motion = sqrt(best_mu / (a * a * a));
    double const steps = fabs(t_final - t) * mean_motion * sqrt(sqrt(a));
    cost[i] = steps == steps ? static_cast<float>(steps) : 0.0f;
    index[i] = static_cast<int>(i);
    if (best_body < kMaxConstantBodies) {
      atomicAdd(block_votes + best_body, 1);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kMaxConstantBodies; b += blockDim.x) {
    if (block_votes[b] != 0) {
      atomicAdd(votes + b, block_votes[b]);
    }
  }
}

// The inertial direction of a burn given in the Frenet frame of the trajectory
// in the body-centred non-rotating frame of body `centre` (internal index):
// Manœuvre::ComputeFrenetFrame, ksp_plugin/manœuvre_body.hpp:383-393, with
// ReferenceFrame::FrenetFrame, physics/reference_frame_body.hpp:38-54 --
// tangent = Normalize(velocity), normal = Normalize(acceleration
// .OrthogonalizationAgainst(velocity)), binormal = Wedge(tangent, normal) --
// on the velocity and the geometric acceleration relative to the centre
// (RigidReferenceFrame::GeometricAcceleration, rigid_reference_frame_body.hpp
// :59-78,364-398: the gravitational acceleration at the point, `gravitational`,
// minus the acceleration of the centre,
// Ephemeris::ComputeGravitationalAccelerationOnMassiveBody,
// ephemeris_body.hpp:1249-1315).  Formed directly in the inertial axes (the
// trihedron is covariant under the rotation to the centre's celestial frame
// that the reference applies and undoes); the oracle does the same operations
// in the same order (oracle/ephemeris.hpp FrenetDirection).  A cold path: it
// evaluates every body's polynomial and the full force on the centre.
__device__ __noinline__ Vec3 FrenetDirection(EphemerisView const& e,
                                             double const t, int const centre,
                                             Vec3 const& direction,
                                             Vec3 const& v,
                                             Vec3 const& gravitational) {
  // One stage-table entry in thread-local memory: positions, then frames.
  double entry[3 * kMaxConstantBodies + 6 * 8];
  int const n = e.n_bodies;
  Vec3 centre_velocity{0.0, 0.0, 0.0};
  for (int b = 0; b < n; ++b) {
    int const segment = FindSegment(e.segment_t_max, e.segment_offset[b],
                                    e.segment_end[b], t);
    Vec3 const p = EvaluateSegment(e, segment, t);
    entry[3 * b] = p.x;
    entry[3 * b + 1] = p.y;
    entry[3 * b + 2] = p.z;
    if (b == centre) {
      centre_velocity = EvaluateSegmentDerivative(e, segment, t);
    }
    int const slot = e.bodies[b].frame_slot;
    if (slot >= 0 && slot < 8) {
      Vec3 x_hat, y_hat;
      SurfaceFrameAxes(e.bodies[b], t, x_hat, y_hat);
      double* const frame = entry + 3 * n + 6 * slot;
      frame[0] = x_hat.x;
      frame[1] = x_hat.y;
      frame[2] = x_hat.z;
      frame[3] = y_hat.x;
      frame[4] = y_hat.y;
      frame[5] = y_hat.z;
    }
  }
  // ComputeGravitationalAccelerationOnMassiveBody(centre, t): the other bodies
  // in increasing internal order (MassiveAccelerationKernel).
  double const* const frames = entry + 3 * n;
  Vec3 const q1{entry[3 * centre], entry[3 * centre + 1],
                entry[3 * centre + 2]};
  Vec3 centre_acceleration{0.0, 0.0, 0.0};
  for (int b2 = 0; b2 < n; ++b2) {
    if (b2 == centre) {
      continue;
    }
    Vec3 const q2{entry[3 * b2], entry[3 * b2 + 1], entry[3 * b2 + 2]};
    Vec3 unused{0.0, 0.0, 0.0};
    MassivePairInteraction<false>(e, centre, b2, frames, q1, q2,
                                  centre_acceleration, unused);
  }
  Vec3 const velocity = v - centre_velocity;
  Vec3 const acceleration = gravitational - centre_acceleration;
  Vec3 const tangent = velocity / sqrt(Norm2(velocity));
  Vec3 const normal_acceleration =
      acceleration - Dot(acceleration, tangent) * tangent;
  Vec3 const normal = normal_acceleration / sqrt(Norm2(normal_acceleration));
  Vec3 const binormal = Cross(tangent, normal);
  return direction.x * tangent + direction.y * normal + direction.z * binormal;
}

// Manœuvre::ComputeIntrinsicAcceleration with the direction resolved by the
// caller: for a Frenet burn the direction is only needed while the engine
// fires.
__device__ __forceinline__ Vec3 GeneralizedBurnIntrinsicAcceleration(
    EphemerisView const& e, double const* const burns,
    int32_t const* const frenet_centres, long long const n, long long const i,
    double const t, Vec3 const& v, Vec3 const& gravitational) {
  int const caller = frenet_centres != nullptr ? frenet_centres[i] : -1;
  if (caller < 0) {
    return BurnIntrinsicAcceleration(burns, n, i, t);
  }
  double const initial_time = burns[6 * n + i];
  double const final_time = burns[7 * n + i];
  if (t >= initial_time && t <= final_time) {
    Vec3 const frenet{burns[i], burns[n + i], burns[2 * n + i]};
    Vec3 const direction = FrenetDirection(e, t, e.internal_of_caller[caller],
                                           frenet, v, gravitational);
    double const thrust = burns[3 * n + i];
    double const initial_mass = burns[4 * n + i];
    double const mass_flow = burns[5 * n + i];
    return direction * thrust / (initial_mass - (t - initial_time) * mass_flow);
  } else {
    return {0.0, 0.0, 0.0};
  }
}

// The embedded methods, integrators/methods.hpp.  a is strictly lower
// triangular, (i, j) -> i (i - 1) / 2 + j.
// DormandالمكاوىPrince1986RKN434FM, :574-597: 4 stages, first same as last.
#define PB_METHOD_TABLE(name, size, ...)                              \
  __device__ __forceinline__ static double name(int const index) {    \
    constexpr double table[size] = {__VA_ARGS__};                     \
    return table[index];                                              \
  }

struct DormandElMikkawyPrince1986RKN434FM {
  static constexpr int kStages = 4;
  static constexpr bool kFirstSameAsLast = true;
  PB_METHOD_TABLE(c, 4, 0.0, 1.0 / 4.0, 7.0 / 10.0, 1.0)
  PB_METHOD_TABLE(a, 6, 1.0 / 32.0,
                  7.0 / 1000.0, 119.0 / 500.0,
                  1.0 / 14.0, 8.0 / 27.0, 25.0 / 189.0)
  // Not a generalized method: no velocity stages.
  PB_METHOD_TABLE(a_prime, 6, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
  PB_METHOD_TABLE(b_hat, 4, 1.0 / 14.0, 8.0 / 27.0, 25.0 / 189.0, 0.0)
  PB_METHOD_TABLE(b_hat_prime, 4, 1.0 / 14.0, 32.0 / 81.0, 250.0 / 567.0,
                  5.0 / 54.0)
  PB_METHOD_TABLE(b, 4, -7.0 / 150.0, 67.0 / 150.0, 3.0 / 20.0, -1.0 / 20.0)
  PB_METHOD_TABLE(b_prime, 4, 13.0 / 21.0, -20.0 / 27.0, 275.0 / 189.0,
                  -1.0 / 3.0)
};

// Fine1987RKNG34, :543-570: an embedded explicit *generalized* method (the
// right-hand side may depend on the velocity), the integrator of the
// generalized FlowWithAdaptiveStep (ephemeris_body.hpp:446-478; the plugin's
// burn integrator, ksp_plugin/integrators.cpp:54-63).  The right-hand side of
// this library never depends on the velocity, so the velocity stages (built
// with a') feed nothing and the generalized Solve
// (embedded_explicit_generalized_runge_kutta_nyström_integrator_body.hpp:
// 44-300) performs exactly the operations below.  5 stages, no FSAL;
// b = b_hat - delta as the reference forms them.
struct Fine1987RKNG34 {
  static constexpr int kStages = 5;
  static constexpr bool kFirstSameAsLast = false;
  PB_METHOD_TABLE(c, 5, 0.0, 2 / 9.0, 1 / 3.0, 3 / 4.0, 1.0)
  PB_METHOD_TABLE(a, 10, 2 / 81.0,
                  1 / 36.0, 1 / 36.0,
                  9 / 128.0, 0.0, 27 / 128.0,
                  11 / 60.0, -3 / 20.0, 9 / 25.0, 8 / 75.0)
  // The velocity stages of the generalized method (used when the right-hand
  // side depends on the velocity).
  PB_METHOD_TABLE(a_prime, 10, 2 / 9.0,
                  1 / 12.0, 1 / 4.0,
                  69 / 128.0, -234 / 128.0, 135 / 64.0,
                  -17 / 12.0, 27 / 4.0, -27 / 5.0, 16 / 15.0)
  PB_METHOD_TABLE(b_hat, 5, 19 / 180.0, 0.0, 63 / 200.0, 16 / 225.0,
                  1 / 120.0)
  PB_METHOD_TABLE(b_hat_prime, 5, 1 / 9.0, 0.0, 9 / 20.0, 16 / 45.0, 1 / 12.0)
  PB_METHOD_TABLE(b, 5, 19 / 180.0 - 25 / 1116.0, 0.0 - 0.0,
                  63 / 200.0 - (-63 / 1240.0), 16 / 225.0 - 64 / 1395.0,
                  1 / 120.0 - (-13 / 744.0))
  PB_METHOD_TABLE(b_prime, 5, 1 / 9.0 - 2 / 125.0, 0.0 - 0.0,
                  9 / 20.0 - (-27 / 625.0), 16 / 45.0 - 32 / 625.0,
                  1 / 12.0 - (-3 / 125.0))
};

#undef PB_METHOD_TABLE

// state: 12 planes of n; time: 2 planes of n (value, error).  `order` lists
// the trajectories in the order in which the lanes take them; `queue` is the
// number of trajectories taken so far (zeroed by the host).
// kGeneralized: the generalized Solve (embedded_explicit_generalized_runge_
// kutta_nyström_integrator_body.hpp:44-300): the velocity stages a' are formed
// and the intrinsic acceleration sees them (Frenet burns).
template<typename Method, int kMinBlocks, bool kGeneralized = false,
         bool kTwoPhase = false>
__global__ void __launch_bounds__(kAdaptiveThreads, kMinBlocks)
ErknFlowKernel(EphemerisView const e, AdaptiveParameters const parameters,
               long long const n, int const* __restrict__ order,
               unsigned long long* __restrict__ queue,
               int32_t const* __restrict__ stop_word,
               double* __restrict__ state, double* __restrict__ time,
               int32_t* __restrict__ status_out,
               long long* __restrict__ accepted_out,
               long long* __restrict__ rejected_out,
               long long* __restrict__ evaluations_out,
               double* __restrict__ step_log,
               double* __restrict__ state_log, DownsamplerView const sink,
               double const* __restrict__ burns,
               int32_t const* __restrict__ frenet_centres) {
  constexpr int kStages = Method::kStages;
  constexpr double safety_factor = 0.9;
  double const t_final = parameters.t_final;

  // The trajectory of this lane.
  bool busy = false;
  long long i = 0;
  DP t{0.0, 0.0};
  DP q[3], v[3];
  double h = 0.0;
  double tolerance_to_error_ratio = 0.0;
  // The stage accelerations; every index is a compile-time constant after
  // unrolling, so that they stay in registers.
  Vec3 g[kStages];
#pragma unroll
  for (int j = 0; j < kStages; ++j) {
    g[j] = {0.0, 0.0, 0.0};
  }
  bool first_iteration = true;
  int first_stage = 0;
  int hint = 0;
  long long accepted = 0, rejected = 0, evaluations = 0;
  int32_t status = PB_OK, step_status = PB_OK, final_status = PB_OK;

  for (;;) {
    bool finished = false;
    // Whether the exit goes through the epilogue of FlowODEWithAdaptiveStep.
    bool flowed = true;
    if (!busy) {
      unsigned long long const ticket = atomicAdd(queue, 1ULL);
      if (ticket >= static_cast<unsigned long long>(n)) {
        break;
      }
      i = order != nullptr ? order[ticket] : static_cast<long long>(ticket);
      t = {time[i], time[n + i]};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        q[k].value = state[(0 + k) * n + i];
        q[k].error = state[(3 + k) * n + i];
        v[k].value = state[(6 + k) * n + i];
        v[k].error = state[(9 + k) * n + i];
      }
      accepted = rejected = evaluations = 0;
      status = step_status = final_status = PB_OK;
      first_iteration = true;
      first_stage = 0;
      hint = 0;
      tolerance_to_error_ratio = 0.0;
      busy = true;
      // FlowODEWithAdaptiveStep, ephemeris_body.hpp:1634-1637.
      if (*reinterpret_cast<int32_t const volatile*>(stop_word) != 0) {
        // RETURN_IF_STOPPED before the first step: the trajectory is left
        // where it was.
        final_status = PB_CANCELLED;
        finished = true;
        flowed = false;
      } else if (t.value == parameters.t_requested) {
        finished = true;
        flowed = false;
      } else if (!(t.value >= parameters.t_min)) {
        // CHECK_LE(t_min(), time) in
        // ContinuousTrajectory::EvaluatePositionLocked (also a NaN time).
        final_status = PB_INVALID_ARGUMENT;
        finished = true;
        flowed = false;
      } else {
        h = t_final - t.value;  // first_step.
        if (!(h > 0.0)) {
          // CHECK_GT(first_step, 0): "Flow back to the future".
          final_status = PB_INVALID_ARGUMENT;
          finished = true;
          flowed = false;
        }
      }
    }
    if (!finished && !first_iteration) {
      step_status = PB_OK;
      // Root<lower_order + 1> = x^(1/4).
      h *= safety_factor * CrRoot4(tolerance_to_error_ratio);
      if (t.value + (t.error + h) == t.value) {
        final_status = PB_FAILED_PRECONDITION;  // VanishingStepSize.
        finished = true;
      }
    }
    if (!finished) {
      first_iteration = false;
      // last_step_is_exact.
      double const time_to_end = (t_final - t.value) - t.error;
      bool const at_end = h >= time_to_end;
      if (at_end) {
        h = time_to_end;
      }
      double const h2 = h * h;
#pragma unroll 1
      for (int s = first_stage; s < kStages; ++s) {
        double const c_s = Method::c(s);
        double const t_stage = (at_end && c_s == 1.0)
                                   ? t_final
                                   : t.value + (t.error + c_s * h);
        // sum_j a(s, j) g_j over j < s, in increasing j.
        Vec3 sum{0.0, 0.0, 0.0};
        int const row = s * (s - 1) / 2;
#pragma unroll
        for (int j = 0; j < kStages - 1; ++j) {
          if (j < s) {
            sum += Method::a(row + j) * g[j];
          }
        }
        Vec3 const q_stage{q[0].value + h * c_s * v[0].value + h2 * sum.x,
                           q[1].value + h * c_s * v[1].value + h2 * sum.y,
                           q[2].value + h * c_s * v[2].value + h2 * sum.z};
        Vec3 acceleration;
        int32_t const rhs_error =
            FastAccelerationAtTime<kTwoPhase>(e, t_stage, q_stage, acceleration,
                                              hint);
        if constexpr (kGeneralized) {
          if (burns != nullptr) {
            // :186-201 of the generalized integrator: the velocity stage, then
            // ephemeris_body.hpp:462-470.
            Vec3 sum_prime{0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < kStages - 1; ++j) {
              if (j < s) {
                sum_prime += Method::a_prime(row + j) * g[j];
              }
            }
            Vec3 const v_stage{v[0].value + h * sum_prime.x,
                               v[1].value + h * sum_prime.y,
                               v[2].value + h * sum_prime.z};
            acceleration += GeneralizedBurnIntrinsicAcceleration(
                e, burns, frenet_centres, n, i, t_stage, v_stage, acceleration);
          }
        } else if (burns != nullptr) {
          // ephemeris_body.hpp:430-432.
          acceleration += BurnIntrinsicAcceleration(burns, n, i, t_stage);
        }
#pragma unroll
        for (int j = 0; j < kStages; ++j) {
          if (j == s) {
            g[j] = acceleration;
          }
        }
        ++evaluations;
        if (step_status == PB_OK && rhs_error != 0) {
          step_status = PB_OUT_OF_RANGE;
        }
      }
      double position_error[3], velocity_error[3];
      double dqh[3], dvh[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double sum_b_hat = 0.0, sum_b = 0.0, sum_b_hat_prime = 0.0,
               sum_b_prime = 0.0;
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
          double const gs = k == 0 ? g[s].x : (k == 1 ? g[s].y : g[s].z);
          sum_b_hat += Method::b_hat(s) * gs;
          sum_b += Method::b(s) * gs;
          sum_b_hat_prime += Method::b_hat_prime(s) * gs;
          sum_b_prime += Method::b_prime(s) * gs;
        }
        dqh[k] = h * v[k].value + h2 * sum_b_hat;
        double const dq_low = h * v[k].value + h2 * sum_b;
        dvh[k] = h * sum_b_hat_prime;
        double const dv_low = h * sum_b_prime;
        position_error[k] = dq_low - dqh[k];
        velocity_error[k] = dv_low - dvh[k];
      }
      // Ephemeris::ToleranceToErrorRatio, ephemeris_body.hpp:1698-1717.
      Vec3 const pe{position_error[0], position_error[1], position_error[2]};
      Vec3 const ve{velocity_error[0], velocity_error[1], velocity_error[2]};
      double const max_length_error = StdMax(0.0, sqrt(Norm2(pe)));
      double const max_speed_error = StdMax(0.0, sqrt(Norm2(ve)));
      tolerance_to_error_ratio =
          StdMin(parameters.length_integration_tolerance / max_length_error,
                 parameters.speed_integration_tolerance / max_speed_error);
      if (tolerance_to_error_ratio < 1.0) {
        ++rejected;
      } else {
        if (status == PB_OK) {
          status = step_status;
        }
        if constexpr (Method::kFirstSameAsLast) {
          // swap(g.front(), g.back()).
          Vec3 const front = g[0];
          g[0] = g[kStages - 1];
          g[kStages - 1] = front;
          first_stage = 1;
        }
        Increment(t, h);
        Increment(q[0], dqh[0]);
        Increment(q[1], dqh[1]);
        Increment(q[2], dqh[2]);
        Increment(v[0], dvh[0]);
        Increment(v[1], dvh[1]);
        Increment(v[2], dvh[2]);
        // append_state.
        if (accepted < parameters.step_log_capacity) {
          if (step_log != nullptr) {
            step_log[accepted * n + i] = t.value;
          }
          if (state_log != nullptr) {
            double* const out = state_log + accepted * 6 * n;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              out[k * n + i] = q[k].value;
              out[(3 + k) * n + i] = v[k].value;
            }
          }
        }
        if (sink.n > 0) {
          AppendToSink(sink, i, t.value, Vec3{q[0].value, q[1].value, q[2].value},
                       Vec3{v[0].value, v[1].value, v[2].value});
        }
        ++accepted;
        if (at_end) {
          finished = true;
        } else if (accepted == parameters.max_steps) {
          final_status = PB_RESOURCE_EXHAUSTED;  // ReachedMaximalStepCount.
          finished = true;
        } else if (*reinterpret_cast<int32_t const volatile*>(stop_word) !=
                   0) {
          // RETURN_IF_STOPPED after the step,
          // embedded_explicit_runge_kutta_nyström_integrator_body.hpp:243: the
          // state is that of the last accepted step.
          final_status = PB_CANCELLED;
          finished = true;
          flowed = false;
        }
      }
    }
    if (finished) {
      if (flowed) {
        if (final_status == PB_OK) {
          final_status = status;
        }
        // ephemeris_body.hpp:1676-1695: collisions are swallowed; not reaching
        // t is DeadlineExceeded.
        if (final_status == PB_OUT_OF_RANGE) {
          final_status = PB_OK;
        }
        if (final_status == PB_OK && t_final != parameters.t_requested) {
          final_status = PB_DEADLINE_EXCEEDED;
        }
      }
      // The last appended state (the trajectory's back()).
      time[i] = t.value;
      time[n + i] = t.error;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        state[(0 + k) * n + i] = q[k].value;
        state[(3 + k) * n + i] = q[k].error;
        state[(6 + k) * n + i] = v[k].value;
        state[(9 + k) * n + i] = v[k].error;
      }
      if (status_out != nullptr) {
        status_out[i] = final_status;
      }
      if (accepted_out != nullptr) {
        accepted_out[i] = accepted;
      }
      if (rejected_out != nullptr) {
        rejected_out[i] = rejected;
      }
      if (evaluations_out != nullptr) {
        evaluations_out[i] = evaluations;
      }
      busy = false;
    }
  }
}

}  // namespace pb
