// Fixed-step flow of massless bodies with a symmetric linear multistep
// integrator (the plugin's history integrator is Quinlan1999Order8A at 10 s,
// ksp_plugin/integrators.cpp:66-72):
// SymmetricLinearMultistepIntegrator::Instance::Solve,
// integrators/symmetric_linear_multistep_integrator_body.hpp:22-152, one
// right-hand-side evaluation per step.  One thread per probe; the history of
// the last `order` steps is a ring in global memory, [slot][xyz][probe], that
// stays in L2 (576 B per probe at order 8).  Included by principia_b200.cu
// after pb_kernels_massless.cuh (it uses its constant-bank right-hand side).
#pragma once

#include <cuda_runtime.h>

#include "pb_kernels_massless.cuh"
#include "pb_model.h"

namespace pb {

struct MasslessHistory {
  double* displacement_value;
  double* displacement_error;
  double* acceleration;
};

__device__ __forceinline__ long long MasslessHistoryIndex(
    int const slot, int const component, long long const n,
    long long const probe) {
  return (static_cast<long long>(slot) * 3 + component) * n + probe;
}

// Starter::FillStepFromState, symmetric_linear_multistep_integrator_body.hpp:
// 246-263, for every probe: displacement = position - DoublePrecision<
// Position>(), acceleration = RHS at the state.  `stage` is the stage-table
// entry of the state's time.
static __global__ void __launch_bounds__(kFlowThreads)
MasslessFillStepKernel(EphemerisView const e, double const* __restrict__ stage,
                       double const time, double const* __restrict__ burns,
                       long long const n, double const* __restrict__ state,
                       MasslessHistory const history, int const slot,
                       int32_t* status) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  double position[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    DP const p{state[(0 + c) * n + i], state[(3 + c) * n + i]};
    DP const displacement = Subtract(p, DP{0.0, 0.0});
    long long const hi = MasslessHistoryIndex(slot, c, n, i);
    history.displacement_value[hi] = displacement.value;
    history.displacement_error[hi] = displacement.error;
    position[c] = p.value;
  }
  Vec3 a;
  int32_t const error = AccelerationOnMasslessBodyHot(
      e, stage, Vec3{position[0], position[1], position[2]}, a);
  if (burns != nullptr) {
    a += BurnIntrinsicAcceleration(burns, n, i, time);
  }
  history.acceleration[MasslessHistoryIndex(slot, 0, n, i)] = a.x;
  history.acceleration[MasslessHistoryIndex(slot, 1, n, i)] = a.y;
  history.acceleration[MasslessHistoryIndex(slot, 2, n, i)] = a.z;
  if (error != 0) {
    atomicOr(status, error);
  }
}

// `n_steps` steps of Instance::Solve (:69-149) with the velocity of
// ComputeVelocityUsingCohenHubbardOesterwinter (:281-321).  `oldest` is the
// ring slot of previous_steps.front() at the first step; every step overwrites
// the oldest slot.  stage_table: one entry per step (the time after the
// increment).  The ODE::State (12 planes) is left at the last step; sampling as
// in SrknFlowKernel.
// Probes [i_begin, i_end) of n (the planes have n entries): the host cuts a
// launch into probe slices and step sub-chunks on helper streams, like the SRKN
// flow (principia_b200.cu).
template<int kOrder, int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
MasslessMultistepFlowKernel(EphemerisView const e, MultistepMethod const method,
                            double const h,
                            double const* __restrict__ stage_table,
                            double const* __restrict__ stage_times,
                            double const* __restrict__ burns,
                            int const n_steps, long long const n,
                            long long const i_begin, long long const i_end,
                            double* __restrict__ state,
                            MasslessHistory const history, int const oldest,
                            long long const steps_before,
                            long long const sample_every,
                            long long const max_samples,
                            double* __restrict__ samples, int32_t* status) {
  long long const i = i_begin +
                      static_cast<long long>(blockIdx.x) * blockDim.x +
                      threadIdx.x;
  if (i >= i_end) {
    return;
  }
  constexpr int k = kOrder;
  int32_t error = 0;
  int front = oldest;
  DP last_position[3];
  double last_velocity[3] = {0.0, 0.0, 0.0};
  for (int s = 0; s < n_steps; ++s) {
    auto slot_of = [&](int const j) {
      int const slot = front + j;
      return slot >= k ? slot - k : slot;
    };
    DP new_displacement[3];
    DP previous_displacement[3];
    double position[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      DP displacement[k];
      double acceleration[k];
#pragma unroll
      for (int j = 0; j < k; ++j) {
        long long const hi = MasslessHistoryIndex(slot_of(j), c, n, i);
        displacement[j] = {history.displacement_value[hi],
                           history.displacement_error[hi]};
        acceleration[j] = history.acceleration[hi];
      }
      previous_displacement[c] = displacement[k - 1];
      // j = 0.
      DP sum = Scale(-method.alpha[0], displacement[0]);
      double weighted = method.beta_numerator[0] * acceleration[0];
      // Generic j paired with k - j.
#pragma unroll
      for (int j = 1; j < k / 2; ++j) {
        sum = Subtract(sum, Scale(method.alpha[j], displacement[j]));
        sum = Subtract(sum, Scale(method.alpha[j], displacement[k - j]));
        weighted +=
            method.beta_numerator[j] * (acceleration[j] + acceleration[k - j]);
      }
      // j = k / 2.
      sum = Subtract(sum, Scale(method.alpha[k / 2], displacement[k / 2]));
      weighted += method.beta_numerator[k / 2] * acceleration[k / 2];
      Increment(sum, h * h * weighted / method.beta_denominator);
      new_displacement[c] = sum;
      DP const current_position = Add(DP{0.0, 0.0}, sum);
      last_position[c] = current_position;
      position[c] = current_position.value;
    }
    Vec3 a;
    error |= AccelerationOnMasslessBodyHot(
        e, stage_table + static_cast<long long>(s) * e.stage_stride,
        Vec3{position[0], position[1], position[2]}, a);
    if (burns != nullptr) {
      a += BurnIntrinsicAcceleration(burns, n, i, stage_times[s]);
    }
    double const a_new[3] = {a.x, a.y, a.z};
    // Push: the new step replaces the oldest one and becomes
    // previous_steps.back(); the velocity uses the history newest first.
    int const newest = front;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      DP const displacement_change =
          Subtract(new_displacement[c], previous_displacement[c]);
      double velocity =
          (displacement_change.value + displacement_change.error) / h;
      double weighted_accelerations = 0.0;
#pragma unroll
      for (int j = 0; j < k; ++j) {
        // it = previous_steps.rbegin() + j: the new step, then the old steps
        // k - 1, k - 2, ... 1 (0 is being overwritten).
        double const aj =
            j == 0 ? a_new[c]
                   : history.acceleration[MasslessHistoryIndex(
                         slot_of(k - j), c, n, i)];
        weighted_accelerations += method.cho_numerators[j] * aj;
      }
      velocity += weighted_accelerations * h / method.cho_denominator;
      last_velocity[c] = velocity;
      long long const hi = MasslessHistoryIndex(newest, c, n, i);
      history.displacement_value[hi] = new_displacement[c].value;
      history.displacement_error[hi] = new_displacement[c].error;
      history.acceleration[hi] = a_new[c];
    }
    front = front + 1 == k ? 0 : front + 1;
    if (sample_every > 0) {
      long long const step_number = steps_before + s + 1;
      if (step_number % sample_every == 0) {
        long long const slot = step_number / sample_every - 1;
        if (slot < max_samples) {
          double* const out = samples + slot * 6 * n;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            out[c * n + i] = last_position[c].value;
            out[(3 + c) * n + i] = last_velocity[c];
          }
        }
      }
    }
  }
  if (n_steps > 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      state[(0 + c) * n + i] = last_position[c].value;
      state[(3 + c) * n + i] = last_position[c].error;
      state[(6 + c) * n + i] = last_velocity[c];
      state[(9 + c) * n + i] = 0.0;
    }
  }
  if (error != 0) {
    atomicOr(status, error);
  }
}

}  // namespace pb
