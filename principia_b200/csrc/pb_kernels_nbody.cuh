// Massive-body flows: ensembles of small independent systems (a warp per body,
// a lane per system, the reference's accumulation order preserved) and the
// large-N all-pairs system (tiled, one thread per target).
#pragma once

#include <cuda_runtime.h>

#include "pb_accel.cuh"
#include "pb_model.h"

namespace pb {

constexpr int kMaxEnsembleBodies = 64;

// Thread layout of the ensemble kernels: a block integrates 32 systems; warp w
// owns the bodies w, w + W, ... (W warps) of all of them and lane l is system
// l of the block.  The positions of a stage are exchanged through shared
// memory, [body][xyz][lane].  Within a warp the body is uniform, so the
// branches on oblateness and pair order are uniform and every load of the
// [plane][body][system] arrays is one contiguous 256-byte line.
//
// Every thread evaluates ALL the interactions of its body (each unordered pair
// is computed by both of its bodies, from the same operands, hence with the
// same bits) and accumulates them in the order in which the reference updates
// that body: Ephemeris::ComputeGravitationalAccelerationBetweenAllMassiveBodies
// (physics/ephemeris_body.hpp:1503-1552) visits the pairs b1 < b2 in
// lexicographic order, so body b sees j = 0, 1, ... b - 1 (as b2), then
// b + 1, ... (as b1), and within a pair the point-mass term, the harmonics of
// b1, the harmonics of b2 (:1342-1410).  Twice the arithmetic of the
// third-law formulation, no scatter, n times the parallelism.
constexpr int kEnsembleLanes = 32;
// Bodies per warp: 1 up to 16 bodies, 4 up to kMaxEnsembleBodies, so that a
// block has at most 16 warps; the kernels are compiled for 8 warps (the
// TRAPPIST-1 shape) and for 16.  (ptxas settles on 128 registers, two blocks of
// 8 warps per SM; forcing one block with up to 255 registers was measured 1.6x
// slower on the SRKN startup and no faster on the multistep step.)

struct EnsembleShape {
  int warps;  // W.
  int own;    // Bodies per warp.
};

// A single system (the cooperative form of the right-hand side: the lanes of a
// warp share out the partners of its body) is latency-bound, so its bodies are
// spread over as many warps as a block holds: 2 per warp above 16 bodies (17
// warps for the 34 bodies of the Sol system) instead of 4.
inline EnsembleShape MakeEnsembleShape(int const n_bodies,
                                       long long const n_systems = 0) {
  EnsembleShape shape;
  shape.own = n_bodies <= 16 ? 1 : n_systems == 1 && n_bodies <= 36 ? 2 : 4;
  shape.warps = (n_bodies + shape.own - 1) / shape.own;
  return shape;
}

// The bodies of a small system (TRAPPIST-1: 8) in the constant bank: the body
// records and, for the oblate ones, what their harmonics up to degree
// kEnsembleTableDegree need (the damped J2 of every pair of an ensemble member
// is evaluated through these).  Through the pointers of the EphemerisView every
// pair costs a chain of dependent global loads -- the body record, the
// threshold search, the coefficients -- that the arithmetic does not hide.
constexpr int kEnsembleTableBodies = 16;
constexpr int kEnsembleTableDegree = 3;
constexpr int kEnsembleTableSize =
    (kEnsembleTableDegree + 1) * (kEnsembleTableDegree + 2) / 2;
// The thresholds up to degree kEnsembleTableDegree + 1 tell whether the
// degree exceeds kEnsembleTableDegree.
constexpr int kEnsembleTableDampings = kEnsembleTableDegree + 2;

struct EnsembleBodyTables {
  DeviceBody data;
  Damping degree_damping[kEnsembleTableDampings];
  int32_t n_damping;  // min(n_damping, kEnsembleTableDampings).
  int32_t pad;
  double cos[kEnsembleTableSize];
  double sin[kEnsembleTableSize];
};

struct EnsembleConstants {
  double normalization[kEnsembleTableSize];
  EnsembleBodyTables bodies[kEnsembleTableBodies];
};

static __constant__ EnsembleConstants c_ensemble;

struct EnsembleTables {
  int b;
  __device__ __forceinline__ double Cos(int const k) const {
    return c_ensemble.bodies[b].cos[k];
  }
  __device__ __forceinline__ double Sin(int const k) const {
    return c_ensemble.bodies[b].sin[k];
  }
  __device__ __forceinline__ double Normalization(int const k) const {
    return c_ensemble.normalization[k];
  }
  __device__ __forceinline__ Damping const& DegreeDamping(int const n) const {
    return c_ensemble.bodies[b].degree_damping[n];
  }
  __device__ __forceinline__ Damping const& SectoralDamping() const {
    return c_ensemble.bodies[b].data.sectoral;
  }
};

// GeneralSphericalHarmonicsAcceleration<true, 3> of body `source` with the
// operands of the constant bank: the same operations, hence the same bits.
__device__ __forceinline__ Vec3 EnsembleHarmonics(
    EphemerisView const& e, int const source, double const* const frames,
    Vec3 const& r, double const r_norm, double const r2,
    double const one_over_r3) {
  EnsembleBodyTables const& tables = c_ensemble.bodies[source];
  if (r_norm == r_norm) {
    // LimitingDegree over the first thresholds is min(LimitingDegree,
    // kEnsembleTableDampings): exact whenever it says degree <= 3.
    int const max_degree =
        LimitingDegree(tables.degree_damping, tables.n_damping, r_norm) - 1;
    if (max_degree <= kEnsembleTableDegree) {
      if (max_degree == 1) {
        return {0.0, 0.0, 0.0};
      }
      DeviceBody const& body = tables.data;
      bool const is_zonal =
          body.is_zonal || r_norm > body.sectoral.outer_threshold;
      Vec3 x_hat, y_hat;
      if (is_zonal) {
        x_hat = body.equatorial;
        y_hat = body.biequatorial;
      } else {
        double const* const axes = frames + 6 * body.frame_slot;
        x_hat = {axes[0], axes[1], axes[2]};
        y_hat = {axes[3], axes[4], axes[5]};
      }
      GeopotentialSetup s;
      InitializeSetup(body, x_hat, y_hat, r, r_norm, r2, one_over_r3, s);
      EnsembleTables const g{source};
      return is_zonal
                 ? GeopotentialUnrolledDispatch<true, EnsembleTables,
                                                kEnsembleTableDegree>(
                       g, max_degree, s)
                 : GeopotentialUnrolledDispatch<false, EnsembleTables,
                                                kEnsembleTableDegree>(
                       g, max_degree, s);
    }
  }
  // A NaN distance, or harmonics above degree 3: the pointer path.
  DeviceBody const& body = e.bodies[source];
  GeopotentialBody const g = MakeGeopotentialBody(e, source);
  FrameSource const frame{
      body.frame_slot >= 0 ? frames + 6 * body.frame_slot : nullptr, 0.0};
  return GeneralSphericalHarmonicsAcceleration<true, 3>(g, frame, r, r_norm, r2,
                                                        one_over_r3);
}

// Acceleration on body b of this lane's system.  A template so that every
// kernel shape gets a copy compiled for its own register budget.
template<int kMaxThreads>
static __device__ __noinline__ Vec3 AccelerationOnBodyOfSystem(
    EphemerisView const& e, double const* const frames,
    double const* const positions, int const b, int const lane) {
  int const n = e.n_bodies;
  bool const tables = e.ensemble_tables != 0;
  DeviceBody const& body_b = tables ? c_ensemble.bodies[b].data : e.bodies[b];
  Vec3 const q_b{positions[(3 * b + 0) * kEnsembleLanes + lane],
                 positions[(3 * b + 1) * kEnsembleLanes + lane],
                 positions[(3 * b + 2) * kEnsembleLanes + lane]};
  Vec3 a{0.0, 0.0, 0.0};
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
    if (j == b) {
      continue;
    }
    DeviceBody const& body_j =
        tables ? c_ensemble.bodies[j].data : e.bodies[j];
    Vec3 const q_j{positions[(3 * j + 0) * kEnsembleLanes + lane],
                   positions[(3 * j + 1) * kEnsembleLanes + lane],
                   positions[(3 * j + 2) * kEnsembleLanes + lane]};
    // The pair is (b1, b2) = (j, b) if j < b, (b, j) otherwise.
    bool const b_is_b2 = j < b;
    Vec3 const dq = b_is_b2 ? q_j - q_b : q_b - q_j;  // q_b1 - q_b2.
    double const dq2 = Norm2(dq);
    // The correctly rounded branch-free square root and division
    // (pb_math.cuh) inside their range, the IEEE operators outside.
    bool const in_range = InFastRange(dq2);
    double const dq_norm = in_range ? FastSqrt(dq2) : sqrt(dq2);
    double const one_over_dq3 = in_range ? FastDivide(dq_norm, dq2 * dq2)
                                         : dq_norm / (dq2 * dq2);
    double const mu_j = body_j.mu;
    // a2 += dq * (mu1 / dq^3); a1 -= dq * (mu2 / dq^3): for body b the other
    // body's mu either way.
    double const mu_j_over_dq3 = mu_j * one_over_dq3;
    Vec3 const point_mass = dq * mu_j_over_dq3;
    a = b_is_b2 ? a + point_mass : a - point_mass;
    // Harmonics of b1, then of b2.
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      bool const source_is_b = (which == 0) != b_is_b2;
      int const source = source_is_b ? b : j;
      DeviceBody const& source_body = source_is_b ? body_b : body_j;
      if (!source_body.is_oblate) {
        continue;
      }
      // b1's field at b2: r = -dq; b2's field at b1: r = dq.
      Vec3 effect;
      if (tables) {
        effect = EnsembleHarmonics(e, source, frames, which == 0 ? -dq : dq,
                                   dq_norm, dq2, one_over_dq3);
      } else {
        GeopotentialBody const g = MakeGeopotentialBody(e, source);
        FrameSource const frame{source_body.frame_slot >= 0
                                    ? frames + 6 * source_body.frame_slot
                                    : nullptr,
                                0.0};
        effect = GeneralSphericalHarmonicsAcceleration<true, 3>(
            g, frame, which == 0 ? -dq : dq, dq_norm, dq2, one_over_dq3);
      }
      // b1 oblate: a1 -= mu2 effect, a2 += mu1 effect;
      // b2 oblate: a1 += mu2 effect, a2 -= mu1 effect.
      Vec3 const scaled = mu_j * effect;
      bool const subtract = (which == 0) != b_is_b2;
      a = subtract ? a - scaled : a + scaled;
    }
  }
  return a;
}

// The same acceleration when the block integrates ONE system (the Sol
// ephemeris of Ephemeris::Prolong: 34 bodies, a single lane of every warp would
// work): the lanes of the warp that owns body b take the partners
// j = lane, lane + 32, compute the addends of their pair -- the point-mass
// term, the harmonics of b1, the harmonics of b2, each with the sign it has in
// the sequential form -- and the warp then adds them to the acceleration in the
// reference's order (j ascending, the three addends of a pair in order), one
// shuffle per addend.  a - x is a + (-x) bit for bit, and an addend that the
// sequential form skips (a body that is not oblate) is skipped here too, so
// the result is the sequential one; the latency is that of one pair plus the
// ordered sum instead of n - 1 pairs in a row.  All 32 lanes hold the same
// system (the callers clamp the system index), every lane returns the sum.
template<int kMaxThreads>
static __device__ __noinline__ Vec3 CooperativeAccelerationOnBody(
    EphemerisView const& e, double const* const frames,
    double const* const positions, int const b, int const lane) {
  constexpr int kSets = kMaxEnsembleBodies / kEnsembleLanes;
  int const n = e.n_bodies;
  DeviceBody const& body_b = e.bodies[b];
  Vec3 const q_b{positions[(3 * b + 0) * kEnsembleLanes + lane],
                 positions[(3 * b + 1) * kEnsembleLanes + lane],
                 positions[(3 * b + 2) * kEnsembleLanes + lane]};
  Vec3 addend[kSets][3];
  unsigned present[kSets];
#pragma unroll
  for (int set = 0; set < kSets; ++set) {
    present[set] = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      addend[set][k] = {0.0, 0.0, 0.0};
    }
    int const j = lane + set * kEnsembleLanes;
    if (j >= n || j == b) {
      continue;
    }
    DeviceBody const& body_j = e.bodies[j];
    Vec3 const q_j{positions[(3 * j + 0) * kEnsembleLanes + lane],
                   positions[(3 * j + 1) * kEnsembleLanes + lane],
                   positions[(3 * j + 2) * kEnsembleLanes + lane]};
    bool const b_is_b2 = j < b;
    Vec3 const dq = b_is_b2 ? q_j - q_b : q_b - q_j;  // q_b1 - q_b2.
    double const dq2 = Norm2(dq);
    bool const in_range = InFastRange(dq2);
    double const dq_norm = in_range ? FastSqrt(dq2) : sqrt(dq2);
    double const one_over_dq3 = in_range ? FastDivide(dq_norm, dq2 * dq2)
                                         : dq_norm / (dq2 * dq2);
    double const mu_j = body_j.mu;
    double const mu_j_over_dq3 = mu_j * one_over_dq3;
    Vec3 const point_mass = dq * mu_j_over_dq3;
    addend[set][0] = b_is_b2 ? point_mass : -point_mass;
    present[set] = 1u;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      bool const source_is_b = (which == 0) != b_is_b2;
      int const source = source_is_b ? b : j;
      DeviceBody const& source_body = source_is_b ? body_b : body_j;
      if (!source_body.is_oblate) {
        continue;
      }
      GeopotentialBody const g = MakeGeopotentialBody(e, source);
      FrameSource const frame{source_body.frame_slot >= 0
                                  ? frames + 6 * source_body.frame_slot
                                  : nullptr,
                              0.0};
      Vec3 const effect = GeneralSphericalHarmonicsAcceleration<true, 3>(
          g, frame, which == 0 ? -dq : dq, dq_norm, dq2, one_over_dq3);
      Vec3 const scaled = mu_j * effect;
      bool const subtract = (which == 0) != b_is_b2;
      addend[set][1 + which] = subtract ? -scaled : scaled;
      present[set] |= 2u << which;
    }
  }
  __syncwarp();
  Vec3 a{0.0, 0.0, 0.0};
#pragma unroll
  for (int set = 0; set < kSets; ++set) {
    int const count = n - set * kEnsembleLanes < kEnsembleLanes
                          ? n - set * kEnsembleLanes
                          : kEnsembleLanes;
#pragma unroll 1
    for (int source = 0; source < count; ++source) {
      // The mask is the same in every lane, so the branches are uniform and
      // an absent addend costs no shuffle.
      unsigned const mask = __shfl_sync(0xffffffffu, present[set], source);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (mask & (1u << k)) {
          Vec3 const x{__shfl_sync(0xffffffffu, addend[set][k].x, source),
                       __shfl_sync(0xffffffffu, addend[set][k].y, source),
                       __shfl_sync(0xffffffffu, addend[set][k].z, source)};
          a = a + x;
        }
      }
    }
  }
  return a;
}

// A block that integrates a single system takes the cooperative form.
template<int kMaxThreads>
__device__ __forceinline__ Vec3 AccelerationOnBody(
    EphemerisView const& e, double const* const frames,
    double const* const positions, int const b, int const lane,
    long long const n_systems) {
  return n_systems == 1
             ? CooperativeAccelerationOnBody<kMaxThreads>(e, frames, positions,
                                                          b, lane)
             : AccelerationOnBodyOfSystem<kMaxThreads>(e, frames, positions, b,
                                                       lane);
}

// Index into an ensemble plane set: [plane][body][system].
__device__ __forceinline__ long long EnsembleIndex(int const plane,
                                                   int const body,
                                                   int const n_bodies,
                                                   long long const n_systems,
                                                   long long const system) {
  return (static_cast<long long>(plane) * n_bodies + body) * n_systems + system;
}

// One step of SymplecticRungeKuttaNyströmIntegrator::Instance::Solve
// (symplectic_runge_kutta_nyström_integrator_body.hpp:97-135) for every
// system; the ODE::State (12 planes) is advanced in place.  `frames` holds the
// surface-frame axes at each evaluation time of the step (6 n_frames doubles
// per evaluation).  Dynamic shared memory: 3 n_bodies kEnsembleLanes doubles.
template<int kOwn, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads)
EnsembleSrknStepKernel(EphemerisView const e,
                                       SrknTableau const tableau,
                                       double const h,
                                       double const* __restrict__ frames,
                                       long long const n_systems,
                                       double* __restrict__ state) {
  extern __shared__ double positions[];
  int const lane = threadIdx.x % kEnsembleLanes;
  int const warp = threadIdx.x / kEnsembleLanes;
  int const warps = blockDim.x / kEnsembleLanes;
  int const n = e.n_bodies;
  long long const system =
      static_cast<long long>(blockIdx.x) * kEnsembleLanes + lane;
  bool const active = system < n_systems;
  // A lane beyond the end shadows the last system and stores nothing.
  long long const s = active ? system : n_systems - 1;
  Vec3 q[kOwn], v[kOwn], dq[kOwn], dv[kOwn];
#pragma unroll
  for (int o = 0; o < kOwn; ++o) {
    int const b = warp + o * warps;
    dq[o] = {0.0, 0.0, 0.0};
    dv[o] = {0.0, 0.0, 0.0};
    if (b < n) {
      q[o] = {state[EnsembleIndex(0, b, n, n_systems, s)],
              state[EnsembleIndex(1, b, n, n_systems, s)],
              state[EnsembleIndex(2, b, n, n_systems, s)]};
      v[o] = {state[EnsembleIndex(6, b, n, n_systems, s)],
              state[EnsembleIndex(7, b, n, n_systems, s)],
              state[EnsembleIndex(8, b, n, n_systems, s)]};
      if (tableau.first_stage == 1) {
        double const ha = h * tableau.a[0];
        dq[o].x += ha * (v[o].x + dv[o].x);
        dq[o].y += ha * (v[o].y + dv[o].y);
        dq[o].z += ha * (v[o].z + dv[o].z);
      }
    }
  }
  int const frame_stride = 6 * e.n_frames;
  for (int st = tableau.first_stage; st < tableau.stages; ++st) {
#pragma unroll
    for (int o = 0; o < kOwn; ++o) {
      int const b = warp + o * warps;
      if (b < n) {
        positions[(3 * b + 0) * kEnsembleLanes + lane] = q[o].x + dq[o].x;
        positions[(3 * b + 1) * kEnsembleLanes + lane] = q[o].y + dq[o].y;
        positions[(3 * b + 2) * kEnsembleLanes + lane] = q[o].z + dq[o].z;
      }
    }
    __syncthreads();
    double const hb = h * tableau.b[st];
    double const ha = h * tableau.a[st];
#pragma unroll
    for (int o = 0; o < kOwn; ++o) {
      int const b = warp + o * warps;
      if (b < n) {
        Vec3 const g = AccelerationOnBody<kMaxThreads>(
            e, frames + (st - tableau.first_stage) * frame_stride, positions, b,
            lane, n_systems);
        dv[o].x += hb * g.x;
        dv[o].y += hb * g.y;
        dv[o].z += hb * g.z;
        dq[o].x += ha * (v[o].x + dv[o].x);
        dq[o].y += ha * (v[o].y + dv[o].y);
        dq[o].z += ha * (v[o].z + dv[o].z);
      }
    }
    __syncthreads();
  }
  if (!active) {
    return;
  }
#pragma unroll
  for (int o = 0; o < kOwn; ++o) {
    int const b = warp + o * warps;
    if (b < n) {
      double const increments[6] = {dq[o].x, dq[o].y, dq[o].z,
                                    dv[o].x, dv[o].y, dv[o].z};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        long long const qi = EnsembleIndex(c, b, n, n_systems, s);
        long long const qe = EnsembleIndex(3 + c, b, n, n_systems, s);
        DP qd{state[qi], state[qe]};
        Increment(qd, increments[c]);
        state[qi] = qd.value;
        state[qe] = qd.error;
        long long const vi = EnsembleIndex(6 + c, b, n, n_systems, s);
        long long const ve = EnsembleIndex(9 + c, b, n, n_systems, s);
        DP vd{state[vi], state[ve]};
        Increment(vd, increments[3 + c]);
        state[vi] = vd.value;
        state[ve] = vd.error;
      }
    }
  }
}

// History of the symmetric linear multistep integrator, a ring of `order`
// steps: displacement value/error planes and acceleration planes, each
// [slot][component][body][system].
struct MultistepHistory {
  double* displacement_value;
  double* displacement_error;
  double* acceleration;
};

__device__ __forceinline__ long long HistoryIndex(int const slot,
                                                  int const component,
                                                  int const body,
                                                  int const n_bodies,
                                                  long long const n_systems,
                                                  long long const system) {
  return ((static_cast<long long>(slot) * 3 + component) * n_bodies + body) *
             n_systems +
         system;
}

// Starter::FillStepFromState, symmetric_linear_multistep_integrator_body.hpp:
// 246-263: displacement = position - DoublePrecision<Position>() (a
// renormalisation), acceleration = RHS at the state's positions.
template<int kOwn, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads)
EnsembleFillStepKernel(EphemerisView const e,
                                       double const* __restrict__ frames,
                                       long long const n_systems,
                                       double const* __restrict__ state,
                                       MultistepHistory const history,
                                       int const slot) {
  extern __shared__ double positions[];
  int const lane = threadIdx.x % kEnsembleLanes;
  int const warp = threadIdx.x / kEnsembleLanes;
  int const warps = blockDim.x / kEnsembleLanes;
  int const n = e.n_bodies;
  long long const system =
      static_cast<long long>(blockIdx.x) * kEnsembleLanes + lane;
  bool const active = system < n_systems;
  long long const s = active ? system : n_systems - 1;
#pragma unroll
  for (int o = 0; o < kOwn; ++o) {
    int const b = warp + o * warps;
    if (b < n) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        DP const p{state[EnsembleIndex(c, b, n, n_systems, s)],
                   state[EnsembleIndex(3 + c, b, n, n_systems, s)]};
        DP const displacement = Subtract(p, DP{0.0, 0.0});
        if (active) {
          long long const hi = HistoryIndex(slot, c, b, n, n_systems, s);
          history.displacement_value[hi] = displacement.value;
          history.displacement_error[hi] = displacement.error;
        }
        positions[(3 * b + c) * kEnsembleLanes + lane] = p.value;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int o = 0; o < kOwn; ++o) {
    int const b = warp + o * warps;
    if (b < n) {
      Vec3 const a = AccelerationOnBody<kMaxThreads>(e, frames, positions, b,
                                                     lane, n_systems);
      if (active) {
        history.acceleration[HistoryIndex(slot, 0, b, n, n_systems, s)] = a.x;
        history.acceleration[HistoryIndex(slot, 1, b, n, n_systems, s)] = a.y;
        history.acceleration[HistoryIndex(slot, 2, b, n, n_systems, s)] = a.z;
      }
    }
  }
}

// One step of SymmetricLinearMultistepIntegrator::Instance::Solve,
// symmetric_linear_multistep_integrator_body.hpp:69-149, with the velocity of
// ComputeVelocityUsingCohenHubbardOesterwinter (:281-321).  `oldest` is the
// ring slot of previous_steps.front(); the new step overwrites it.  kOrder is
// method.order (loops over the history unrolled, its loads issued together).
template<int kOwn, int kMaxThreads, int kOrder>
__global__ void __launch_bounds__(kMaxThreads)
EnsembleMultistepStepKernel(
    EphemerisView const e, MultistepMethod const method, double const h,
    double const* __restrict__ frames, long long const n_systems,
    double* __restrict__ state, MultistepHistory const history,
    int const oldest) {
  extern __shared__ double positions[];
  int const lane = threadIdx.x % kEnsembleLanes;
  int const warp = threadIdx.x / kEnsembleLanes;
  int const warps = blockDim.x / kEnsembleLanes;
  int const n = e.n_bodies;
  long long const system =
      static_cast<long long>(blockIdx.x) * kEnsembleLanes + lane;
  bool const active = system < n_systems;
  long long const s = active ? system : n_systems - 1;
  constexpr int k = kOrder;
  auto slot_of = [&](int const j) {
    int const slot = oldest + j;
    return slot >= k ? slot - k : slot;
  };
  // Nothing but indices is live across the right-hand side: the new
  // displacement goes straight into the ring slot of the oldest step (whose
  // data this thread has just consumed), and the velocity formula reads the
  // history again (L2) instead of holding 36 doubles in spilled registers.
  int const newest = oldest;
#pragma unroll
  for (int o = 0; o < kOwn; ++o) {
    int const b = warp + o * warps;
    if (b < n) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        DP displacement[k];
        double acceleration[k];
#pragma unroll
        for (int j = 0; j < k; ++j) {
          long long const hi = HistoryIndex(slot_of(j), c, b, n, n_systems, s);
          displacement[j] = {history.displacement_value[hi],
                             history.displacement_error[hi]};
          acceleration[j] = history.acceleration[hi];
        }
        // j = 0.
        DP sum = Scale(-method.alpha[0], displacement[0]);
        double weighted = method.beta_numerator[0] * acceleration[0];
        // Generic j paired with k - j.
#pragma unroll
        for (int j = 1; j < k / 2; ++j) {
          sum = Subtract(sum, Scale(method.alpha[j], displacement[j]));
          sum = Subtract(sum, Scale(method.alpha[j], displacement[k - j]));
          weighted += method.beta_numerator[j] *
                      (acceleration[j] + acceleration[k - j]);
        }
        // j = k / 2.
        sum = Subtract(sum, Scale(method.alpha[k / 2], displacement[k / 2]));
        weighted += method.beta_numerator[k / 2] * acceleration[k / 2];
        Increment(sum, h * h * weighted / method.beta_denominator);
        DP const current_position = Add(DP{0.0, 0.0}, sum);
        positions[(3 * b + c) * kEnsembleLanes + lane] =
            current_position.value;
        if (active) {
          long long const hi = HistoryIndex(newest, c, b, n, n_systems, s);
          history.displacement_value[hi] = sum.value;
          history.displacement_error[hi] = sum.error;
          state[EnsembleIndex(c, b, n, n_systems, s)] = current_position.value;
          state[EnsembleIndex(3 + c, b, n, n_systems, s)] =
              current_position.error;
        }
      }
    }
  }
  __syncthreads();
  // Push: the new step replaces the oldest one and becomes previous_steps.back.
#pragma unroll
  for (int o = 0; o < kOwn; ++o) {
    int const b = warp + o * warps;
    if (b < n) {
      Vec3 const acceleration = AccelerationOnBody<kMaxThreads>(
          e, frames, positions, b, lane, n_systems);
      double const a[3] = {acceleration.x, acceleration.y, acceleration.z};
      if (active) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          long long const hi = HistoryIndex(newest, c, b, n, n_systems, s);
          DP const displacement{history.displacement_value[hi],
                                history.displacement_error[hi]};
          // Velocity; previous_steps.back() before the push is j = k - 1.
          long long const pi =
              HistoryIndex(slot_of(k - 1), c, b, n, n_systems, s);
          DP const displacement_change = Subtract(
              displacement, DP{history.displacement_value[pi],
                               history.displacement_error[pi]});
          double velocity =
              (displacement_change.value + displacement_change.error) / h;
          double weighted_accelerations = 0.0;
#pragma unroll
          for (int i = 0; i < k; ++i) {
            // it = previous_steps.rbegin() + i: the newest step first, then
            // the old steps j = k - 1, k - 2, ... 1 (j = 0 is overwritten).
            double const ai =
                i == 0 ? a[c]
                       : history.acceleration[HistoryIndex(
                             slot_of(k - i), c, b, n, n_systems, s)];
            weighted_accelerations += method.cho_numerators[i] * ai;
          }
          velocity += weighted_accelerations * h / method.cho_denominator;
          history.acceleration[hi] = a[c];
          state[EnsembleIndex(6 + c, b, n, n_systems, s)] = velocity;
          state[EnsembleIndex(9 + c, b, n, n_systems, s)] = 0.0;
        }
      }
    }
  }
}

// Surface-frame axes of the bodies with a frame slot at `n_times` instants:
// out[time][6 n_frames].
static __global__ void FrameTableKernel(EphemerisView const e,
                                 double const* __restrict__ times,
                                 int const n_times, double* __restrict__ out) {
  int const index = blockIdx.x * blockDim.x + threadIdx.x;
  if (index >= n_times * e.n_bodies) {
    return;
  }
  int const time_index = index / e.n_bodies;
  int const b = index % e.n_bodies;
  int const slot = e.bodies[b].frame_slot;
  if (slot >= 0) {
    Vec3 x_hat, y_hat;
    SurfaceFrameAxes(e.bodies[b], times[time_index], x_hat, y_hat);
    double* const frame = out + (static_cast<long long>(time_index) *
                                 e.n_frames + slot) * 6;
    frame[0] = x_hat.x;
    frame[1] = x_hat.y;
    frame[2] = x_hat.z;
    frame[3] = y_hat.x;
    frame[4] = y_hat.y;
    frame[5] = y_hat.z;
  }
}

}  // namespace pb
