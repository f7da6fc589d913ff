// Adaptive massless flow: EmbeddedExplicitRungeKuttaNyströmIntegrator::
// Instance::Solve (integrators/embedded_explicit_runge_kutta_nyström_
// integrator_body.hpp:44-258) with DormandالمكاوىPrince1986RKN434FM
// (integrators/methods.hpp:574-597), driven like
// Ephemeris::FlowODEWithAdaptiveStep (physics/ephemeris_body.hpp:1624-1717):
// one independent trajectory per thread; step size, error control, rejection
// loop and FSAL state all in registers.
#pragma once

#include <cuda_runtime.h>

#include "pb_accel.cuh"
#include "pb_model.h"

namespace pb {

constexpr int kAdaptiveThreads = 64;

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
    int const end = e.segment_offset[b1 + 1];
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

// state: 12 planes of n; time: 2 planes of n (value, error).
template<bool kUnrolled>
__global__ void __launch_bounds__(kAdaptiveThreads)
ErknFlowKernel(EphemerisView const e, AdaptiveParameters const parameters,
               long long const n, double* __restrict__ state,
               double* __restrict__ time, int32_t* __restrict__ status_out,
               long long* __restrict__ accepted_out,
               long long* __restrict__ rejected_out,
               long long* __restrict__ evaluations_out,
               double* __restrict__ step_log,
               double* __restrict__ state_log) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  // DormandالمكاوىPrince1986RKN434FM, methods.hpp:574-597.
  constexpr int stages = 4;
  constexpr double c[4] = {0.0, 1.0 / 4.0, 7.0 / 10.0, 1.0};
  constexpr double a10 = 1.0 / 32.0;
  constexpr double a20 = 7.0 / 1000.0, a21 = 119.0 / 500.0;
  constexpr double a30 = 1.0 / 14.0, a31 = 8.0 / 27.0, a32 = 25.0 / 189.0;
  constexpr double b_hat[4] = {1.0 / 14.0, 8.0 / 27.0, 25.0 / 189.0, 0.0};
  constexpr double b_hat_prime[4] = {1.0 / 14.0, 32.0 / 81.0, 250.0 / 567.0,
                                     5.0 / 54.0};
  constexpr double b[4] = {-7.0 / 150.0, 67.0 / 150.0, 3.0 / 20.0, -1.0 / 20.0};
  constexpr double b_prime[4] = {13.0 / 21.0, -20.0 / 27.0, 275.0 / 189.0,
                                 -1.0 / 3.0};
  constexpr double safety_factor = 0.9;

  DP t{time[i], time[n + i]};
  DP q[3], v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    q[k].value = state[(0 + k) * n + i];
    q[k].error = state[(3 + k) * n + i];
    v[k].value = state[(6 + k) * n + i];
    v[k].error = state[(9 + k) * n + i];
  }
  long long accepted = 0, rejected = 0, evaluations = 0;
  int32_t final_status = PB_OK;
  double const t_final = parameters.t_final;

  // FlowODEWithAdaptiveStep, ephemeris_body.hpp:1634-1637.
  if (t.value != parameters.t_requested) {
    double h = t_final - t.value;  // first_step.
    if (!(h > 0.0)) {
      // CHECK_GT(first_step, 0): "Flow back to the future".
      final_status = PB_INVALID_ARGUMENT;
    } else {
      Vec3 g[4];
      bool at_end = false;
      double tolerance_to_error_ratio = 0.0;
      int first_stage = 0;
      long long step_count = 0;
      int32_t status = PB_OK;
      int32_t step_status = PB_OK;
      bool first_iteration = true;
      int hint = 0;
      bool done = false;
      Vec3 dq_hat, dv_hat;
      while (!at_end && !done) {
        do {
          if (!first_iteration) {
            step_status = PB_OK;
            // Root<lower_order + 1> = x^(1/4).
            h *= safety_factor * CrRoot4(tolerance_to_error_ratio);
            if (t.value + (t.error + h) == t.value) {
              final_status = PB_FAILED_PRECONDITION;  // VanishingStepSize.
              done = true;
              break;
            }
          }
          first_iteration = false;
          // last_step_is_exact.
          double const time_to_end = (t_final - t.value) - t.error;
          at_end = h >= time_to_end;
          if (at_end) {
            h = time_to_end;
          }
          double const h2 = h * h;
          for (int s = first_stage; s < stages; ++s) {
            double const t_stage = (at_end && c[s] == 1.0)
                                       ? t_final
                                       : t.value + (t.error + c[s] * h);
            double qs[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              double const g0 = k == 0 ? g[0].x : (k == 1 ? g[0].y : g[0].z);
              double const g1 = k == 0 ? g[1].x : (k == 1 ? g[1].y : g[1].z);
              double const g2 = k == 0 ? g[2].x : (k == 1 ? g[2].y : g[2].z);
              double sum = 0.0;
              if (s == 1) {
                sum += a10 * g0;
              } else if (s == 2) {
                sum += a20 * g0;
                sum += a21 * g1;
              } else if (s == 3) {
                sum += a30 * g0;
                sum += a31 * g1;
                sum += a32 * g2;
              }
              qs[k] = q[k].value + h * c[s] * v[k].value + h2 * sum;
            }
            Vec3 const q_stage{qs[0], qs[1], qs[2]};
            Vec3 acceleration;
            int32_t const rhs_error = AccelerationOnMasslessBodyAtTime<kUnrolled>(
                e, t_stage, q_stage, acceleration, hint);
            g[s] = acceleration;
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
            for (int s = 0; s < stages; ++s) {
              double const gs = k == 0 ? g[s].x : (k == 1 ? g[s].y : g[s].z);
              sum_b_hat += b_hat[s] * gs;
              sum_b += b[s] * gs;
              sum_b_hat_prime += b_hat_prime[s] * gs;
              sum_b_prime += b_prime[s] * gs;
            }
            dqh[k] = h * v[k].value + h2 * sum_b_hat;
            double const dq_low = h * v[k].value + h2 * sum_b;
            dvh[k] = h * sum_b_hat_prime;
            double const dv_low = h * sum_b_prime;
            position_error[k] = dq_low - dqh[k];
            velocity_error[k] = dv_low - dvh[k];
          }
          dq_hat = {dqh[0], dqh[1], dqh[2]};
          dv_hat = {dvh[0], dvh[1], dvh[2]};
          // Ephemeris::ToleranceToErrorRatio, ephemeris_body.hpp:1698-1717.
          Vec3 const pe{position_error[0], position_error[1],
                        position_error[2]};
          Vec3 const ve{velocity_error[0], velocity_error[1],
                        velocity_error[2]};
          double const max_length_error = StdMax(0.0, sqrt(Norm2(pe)));
          double const max_speed_error = StdMax(0.0, sqrt(Norm2(ve)));
          tolerance_to_error_ratio = StdMin(
              parameters.length_integration_tolerance / max_length_error,
              parameters.speed_integration_tolerance / max_speed_error);
          if (tolerance_to_error_ratio < 1.0) {
            ++rejected;
          }
        } while (tolerance_to_error_ratio < 1.0);
        if (done) {
          break;
        }
        if (status == PB_OK) {
          status = step_status;
        }
        // first_same_as_last: swap(g.front(), g.back()).
        {
          Vec3 const front = g[0];
          g[0] = g[3];
          g[3] = front;
          first_stage = 1;
        }
        Increment(t, h);
        Increment(q[0], dq_hat.x);
        Increment(q[1], dq_hat.y);
        Increment(q[2], dq_hat.z);
        Increment(v[0], dv_hat.x);
        Increment(v[1], dv_hat.y);
        Increment(v[2], dv_hat.z);
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
        ++accepted;
        ++step_count;
        if (step_count == parameters.max_steps && !at_end) {
          final_status = PB_RESOURCE_EXHAUSTED;  // ReachedMaximalStepCount.
          done = true;
        }
      }
      if (final_status == PB_OK) {
        final_status = status;
      }
      // ephemeris_body.hpp:1676-1695: collisions are swallowed; not reaching t
      // is DeadlineExceeded.
      if (final_status == PB_OUT_OF_RANGE) {
        final_status = PB_OK;
      }
      if (final_status == PB_OK && t_final != parameters.t_requested) {
        final_status = PB_DEADLINE_EXCEEDED;
      }
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
}

}  // namespace pb
