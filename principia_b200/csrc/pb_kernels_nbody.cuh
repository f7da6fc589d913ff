// Massive-body flows: ensembles of small independent systems (one thread per
// system, the reference's pair order and third-law updates preserved) and the
// large-N all-pairs system (tiled, one thread per target).
#pragma once

#include <cuda_runtime.h>

#include "pb_accel.cuh"
#include "pb_model.h"

namespace pb {

constexpr int kEnsembleThreads = 64;
constexpr int kMaxEnsembleBodies = 64;
constexpr int kMaxMultistepOrder = 12;
// symmetric_linear_multistep_integrator_body.hpp:20.
constexpr int kStartupStepDivisor = 16;

// Ephemeris::ComputeGravitationalAccelerationBetweenAllMassiveBodies,
// physics/ephemeris_body.hpp:1503-1552, for one system held in arrays.
template<int kBodies>
__device__ __forceinline__ void AccelerationsBetweenAllMassiveBodies(
    EphemerisView const& e, double const* const frames, Vec3 const* const q,
    Vec3* const a) {
  int const n = kBodies > 0 ? kBodies : e.n_bodies;
  for (int b = 0; b < n; ++b) {
    a[b] = {0.0, 0.0, 0.0};
  }
  // Not unrolled: one inlined copy of the pair interaction (with its
  // geopotential code) per kernel.
#pragma unroll 1
  for (int b1 = 0; b1 < n; ++b1) {
#pragma unroll 1
    for (int b2 = b1 + 1; b2 < n; ++b2) {
      MassivePairInteraction<true>(e, b1, b2, frames, q[b1], q[b2], a[b1],
                                   a[b2]);
    }
  }
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
// per evaluation).
template<int kBodies>
__global__ void __launch_bounds__(kEnsembleThreads)
EnsembleSrknStepKernel(EphemerisView const e, SrknTableau const tableau,
                       double const h, double const* __restrict__ frames,
                       long long const n_systems, double* __restrict__ state) {
  long long const s =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n_systems) {
    return;
  }
  constexpr int kCapacity = kBodies > 0 ? kBodies : kMaxEnsembleBodies;
  int const n = kBodies > 0 ? kBodies : e.n_bodies;
  Vec3 dq[kCapacity], dv[kCapacity], q_stage[kCapacity], g[kCapacity];
  auto value = [&](int const plane, int const b) {
    return state[EnsembleIndex(plane, b, n, n_systems, s)];
  };
  for (int b = 0; b < n; ++b) {
    dq[b] = {0.0, 0.0, 0.0};
    dv[b] = {0.0, 0.0, 0.0};
  }
  if (tableau.first_stage == 1) {
    double const ha = h * tableau.a[0];
    for (int b = 0; b < n; ++b) {
      dq[b].x += ha * (value(6, b) + dv[b].x);
      dq[b].y += ha * (value(7, b) + dv[b].y);
      dq[b].z += ha * (value(8, b) + dv[b].z);
    }
  }
  int const frame_stride = 6 * e.n_frames;
  for (int st = tableau.first_stage; st < tableau.stages; ++st) {
    for (int b = 0; b < n; ++b) {
      q_stage[b] = {value(0, b) + dq[b].x, value(1, b) + dq[b].y,
                    value(2, b) + dq[b].z};
    }
    AccelerationsBetweenAllMassiveBodies<kBodies>(
        e, frames + (st - tableau.first_stage) * frame_stride, q_stage, g);
    double const hb = h * tableau.b[st];
    double const ha = h * tableau.a[st];
    for (int b = 0; b < n; ++b) {
      dv[b].x += hb * g[b].x;
      dv[b].y += hb * g[b].y;
      dv[b].z += hb * g[b].z;
      dq[b].x += ha * (value(6, b) + dv[b].x);
      dq[b].y += ha * (value(7, b) + dv[b].y);
      dq[b].z += ha * (value(8, b) + dv[b].z);
    }
  }
  for (int b = 0; b < n; ++b) {
    double const increments[6] = {dq[b].x, dq[b].y, dq[b].z,
                                  dv[b].x, dv[b].y, dv[b].z};
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
template<int kBodies>
__global__ void __launch_bounds__(kEnsembleThreads)
EnsembleFillStepKernel(EphemerisView const e,
                       double const* __restrict__ frames,
                       long long const n_systems,
                       double const* __restrict__ state,
                       MultistepHistory const history, int const slot) {
  long long const s =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n_systems) {
    return;
  }
  constexpr int kCapacity = kBodies > 0 ? kBodies : kMaxEnsembleBodies;
  int const n = kBodies > 0 ? kBodies : e.n_bodies;
  Vec3 q[kCapacity], a[kCapacity];
  for (int b = 0; b < n; ++b) {
    double position[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      DP const p{state[EnsembleIndex(c, b, n, n_systems, s)],
                 state[EnsembleIndex(3 + c, b, n, n_systems, s)]};
      DP const displacement = Subtract(p, DP{0.0, 0.0});
      long long const hi = HistoryIndex(slot, c, b, n, n_systems, s);
      history.displacement_value[hi] = displacement.value;
      history.displacement_error[hi] = displacement.error;
      position[c] = p.value;
    }
    q[b] = {position[0], position[1], position[2]};
  }
  AccelerationsBetweenAllMassiveBodies<kBodies>(e, frames, q, a);
  for (int b = 0; b < n; ++b) {
    history.acceleration[HistoryIndex(slot, 0, b, n, n_systems, s)] = a[b].x;
    history.acceleration[HistoryIndex(slot, 1, b, n, n_systems, s)] = a[b].y;
    history.acceleration[HistoryIndex(slot, 2, b, n, n_systems, s)] = a[b].z;
  }
}

// integrators/methods.hpp:998-1007,1040-1055 and
// cohen_hubbard_oesterwinter_body.hpp.
struct MultistepMethod {
  int order;
  double alpha[kMaxMultistepOrder / 2 + 1];
  double beta_numerator[kMaxMultistepOrder / 2 + 1];
  double beta_denominator;
  double cho_numerators[kMaxMultistepOrder];
  double cho_denominator;
};

// One step of SymmetricLinearMultistepIntegrator::Instance::Solve,
// symmetric_linear_multistep_integrator_body.hpp:69-149, with the velocity of
// ComputeVelocityUsingCohenHubbardOesterwinter (:281-321).  `oldest` is the
// ring slot of previous_steps.front(); the new step overwrites it.
template<int kBodies>
__global__ void __launch_bounds__(kEnsembleThreads)
EnsembleMultistepStepKernel(EphemerisView const e, MultistepMethod const method,
                            double const h, double const* __restrict__ frames,
                            long long const n_systems,
                            double* __restrict__ state,
                            MultistepHistory const history, int const oldest) {
  long long const s =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n_systems) {
    return;
  }
  constexpr int kCapacity = kBodies > 0 ? kBodies : kMaxEnsembleBodies;
  int const n = kBodies > 0 ? kBodies : e.n_bodies;
  int const k = method.order;
  auto slot_of = [&](int const j) { return (oldest + j) % k; };
  Vec3 positions[kCapacity], accelerations[kCapacity];
  // The new displacements are kept until the old slot may be overwritten.
  DP new_displacement[3 * kCapacity];
  for (int b = 0; b < n; ++b) {
    double position[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      auto displacement = [&](int const j) {
        long long const hi = HistoryIndex(slot_of(j), c, b, n, n_systems, s);
        return DP{history.displacement_value[hi], history.displacement_error[hi]};
      };
      auto acceleration = [&](int const j) {
        return history.acceleration[HistoryIndex(slot_of(j), c, b, n,
                                                 n_systems, s)];
      };
      // j = 0.
      DP sum = Scale(-method.alpha[0], displacement(0));
      double weighted = method.beta_numerator[0] * acceleration(0);
      // Generic j paired with k - j.
      for (int j = 1; j < k / 2; ++j) {
        sum = Subtract(sum, Scale(method.alpha[j], displacement(j)));
        sum = Subtract(sum, Scale(method.alpha[j], displacement(k - j)));
        weighted += method.beta_numerator[j] *
                    (acceleration(j) + acceleration(k - j));
      }
      // j = k / 2.
      sum = Subtract(sum, Scale(method.alpha[k / 2], displacement(k / 2)));
      weighted += method.beta_numerator[k / 2] * acceleration(k / 2);
      Increment(sum, h * h * weighted / method.beta_denominator);
      new_displacement[3 * b + c] = sum;
      DP const current_position = Add(DP{0.0, 0.0}, sum);
      position[c] = current_position.value;
      state[EnsembleIndex(c, b, n, n_systems, s)] = current_position.value;
      state[EnsembleIndex(3 + c, b, n, n_systems, s)] = current_position.error;
    }
    positions[b] = {position[0], position[1], position[2]};
  }
  AccelerationsBetweenAllMassiveBodies<kBodies>(e, frames, positions,
                                                accelerations);
  // Push: the new step replaces the oldest one and becomes previous_steps.back.
  int const newest = oldest;
  int const previous = (oldest + k - 1) % k;
  for (int b = 0; b < n; ++b) {
    double const a[3] = {accelerations[b].x, accelerations[b].y,
                         accelerations[b].z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      long long const hi = HistoryIndex(newest, c, b, n, n_systems, s);
      DP const displacement = new_displacement[3 * b + c];
      history.displacement_value[hi] = displacement.value;
      history.displacement_error[hi] = displacement.error;
      history.acceleration[hi] = a[c];
      // Velocity.
      long long const pi = HistoryIndex(previous, c, b, n, n_systems, s);
      DP const displacement_change = Subtract(
          displacement,
          DP{history.displacement_value[pi], history.displacement_error[pi]});
      double velocity =
          (displacement_change.value + displacement_change.error) / h;
      double weighted_accelerations = 0.0;
      for (int i = 0; i < k; ++i) {
        // it = previous_steps.rbegin() + i: the newest step first.
        int const slot = (newest + k - i) % k;
        double const ai = i == 0 ? a[c]
                                 : history.acceleration[HistoryIndex(
                                       slot, c, b, n, n_systems, s)];
        weighted_accelerations += method.cho_numerators[i] * ai;
      }
      velocity += weighted_accelerations * h / method.cho_denominator;
      state[EnsembleIndex(6 + c, b, n, n_systems, s)] = velocity;
      state[EnsembleIndex(9 + c, b, n, n_systems, s)] = 0.0;
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

// ===========================================================================
// Large-N all-pairs point masses.  One thread per target i, sources staged in
// shared memory tile by tile; every ordered interaction is evaluated (no
// third-law sharing across threads): 18 algorithmic flops each
// (3 sub, 5 norm², sqrt, mul, div, mul, 3 mul, 3 add).  The arithmetic of one
// interaction is the reference's (ephemeris_body.hpp:1366-1374); the
// summation runs over sources j = 0..N-1 in increasing order per target, which
// differs from the reference's pair order (b1 < b2 with scattered updates), so
// parity with the oracle is to a tolerance, stated in the tests.
constexpr int kAllPairsThreads = 128;

// positions / mu: full arrays (3 planes of n / n).  Targets [i_begin, i_end).
static __global__ void __launch_bounds__(kAllPairsThreads, 4)
AllPairsAccelerationKernel(long long const n, long long const i_begin,
                           long long const i_end,
                           double const* __restrict__ qx,
                           double const* __restrict__ qy,
                           double const* __restrict__ qz,
                           double const* __restrict__ mu,
                           double* __restrict__ ax, double* __restrict__ ay,
                           double* __restrict__ az) {
  __shared__ double sx[kAllPairsThreads];
  __shared__ double sy[kAllPairsThreads];
  __shared__ double sz[kAllPairsThreads];
  __shared__ double sm[kAllPairsThreads];
  long long const i =
      i_begin + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  bool const active = i < i_end;
  double const xi = active ? qx[i] : 0.0;
  double const yi = active ? qy[i] : 0.0;
  double const zi = active ? qz[i] : 0.0;
  double accx = 0.0, accy = 0.0, accz = 0.0;
  bool in_range = true;
  for (long long tile = 0; tile < n; tile += kAllPairsThreads) {
    long long const j = tile + threadIdx.x;
    if (j < n) {
      sx[threadIdx.x] = qx[j];
      sy[threadIdx.x] = qy[j];
      sz[threadIdx.x] = qz[j];
      sm[threadIdx.x] = mu[j];
    }
    __syncthreads();
    int const count =
        static_cast<int>(n - tile < kAllPairsThreads ? n - tile
                                                     : kAllPairsThreads);
    if (active) {
      int const self = static_cast<int>(i - tile);  // Index of i in the tile.
#pragma unroll 4
      for (int jj = 0; jj < count; ++jj) {
        // Δq = position of the source - position of the target: the
        // acceleration on b2 = i due to b1 = j.  Branch-free: the square root
        // and the division are the correctly rounded FastSqrt / FastDivide
        // (pb_math.cuh), so that the scheduler interleaves four interactions;
        // the self-interaction is computed and discarded by a select.
        bool const is_self = jj == self;
        double const dx = sx[jj] - xi;
        double const dy = sy[jj] - yi;
        double const dz = sz[jj] - zi;
        double const d2 = dx * dx + dy * dy + dz * dz;
        in_range = in_range && (InFastRange(d2) || is_self);
        double const d = FastSqrt(d2);
        double const one_over_d3 = FastDivide(d, d2 * d2);
        double const mu_over_d3 = is_self ? 0.0 : sm[jj] * one_over_d3;
        accx += is_self ? 0.0 : dx * mu_over_d3;
        accy += is_self ? 0.0 : dy * mu_over_d3;
        accz += is_self ? 0.0 : dz * mu_over_d3;
      }
    }
    __syncthreads();
  }
  if (active && !in_range) {
    // Some distance is outside the range of the branch-free operators (two
    // bodies at the same place, non-finite state): redo this target with the
    // IEEE operators.
    accx = accy = accz = 0.0;
    for (long long j = 0; j < n; ++j) {
      if (j == i) {
        continue;
      }
      double const dx = qx[j] - xi;
      double const dy = qy[j] - yi;
      double const dz = qz[j] - zi;
      double const d2 = dx * dx + dy * dy + dz * dz;
      double const d = sqrt(d2);
      double const one_over_d3 = d / (d2 * d2);
      double const mu_over_d3 = mu[j] * one_over_d3;
      accx += dx * mu_over_d3;
      accy += dy * mu_over_d3;
      accz += dz * mu_over_d3;
    }
  }
  if (active) {
    long long const o = i - i_begin;
    ax[o] = accx;
    ay[o] = accy;
    az[o] = accz;
  }
}

}  // namespace pb
