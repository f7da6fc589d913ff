// Jerk and Jacobian queries of physics::Ephemeris (point masses only, like the
// reference: "This doesn't take high-order geopotential into account",
// ephemeris_body.hpp:499,568): ComputeGravitationalJerkOnMasslessBody
// (:524-561, batched over probes), ComputeGravitationalJerkOnMassiveBody
// (:563-605,1317-1340,1211-1247) and ComputeJacobianOnMassiveBody
// (:495-522,1174-1209) for every body of the system.
#pragma once

#include <cuda_runtime.h>

#include "pb_accel.cuh"
#include "pb_model.h"

namespace pb {

// Velocity of body b at time t: the derivative of the segment polynomial,
// ContinuousTrajectory::EvaluateVelocityLocked,
// continuous_trajectory_body.hpp:697-706.
PB_HD Vec3 EvaluateSegmentDerivative(EphemerisView const& e, int const segment,
                                     double const t) {
  double const x = t - e.segment_origin[segment];
  double const* const c =
      e.segment_coefficients + static_cast<long long>(segment) * 3 *
                                   kSegmentCoefficients;
  int const degree = e.segment_degree[segment];
  return {EstrinEvaluateDerivative<1>(c, degree, x),
          EstrinEvaluateDerivative<1>(c + kSegmentCoefficients, degree, x),
          EstrinEvaluateDerivative<1>(c + 2 * kSegmentCoefficients, degree, x)};
}

static __global__ void VelocityTableKernel(EphemerisView const e,
                                           double const t,
                                           double* __restrict__ velocities) {
  int const b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= e.n_bodies) {
    return;
  }
  int const segment = FindSegment(e.segment_t_max, e.segment_offset[b],
                                  e.segment_end[b], t);
  Vec3 const v = EvaluateSegmentDerivative(e, segment, t);
  velocities[3 * b] = v.x;
  velocities[3 * b + 1] = v.y;
  velocities[3 * b + 2] = v.z;
}

// −𝟙/‖Δq‖³ + 3 Δq⊗Δq/‖Δq‖⁵ as the R3x3Matrix arithmetic the reference's
// expression expands to (geometry/symmetric_bilinear_form_body.hpp:279-357,
// 406-412; geometry/r3x3_matrix_body.hpp:277-306,354-398): (−𝟙)/n³ entry by
// entry (off the diagonal that is −0/n³), (3·(Δq_i·Δq_j))/n⁵, then the sum.
struct TidalForm {
  double m[3][3];
};

PB_HD TidalForm MakeTidalForm(Vec3 const& dq) {
  double const dq2 = dq.x * dq.x + dq.y * dq.y + dq.z * dq.z;
  double const dq_norm = sqrt(dq2);
  double const dq_norm3 = dq2 * dq_norm;
  double const dq_norm5 = dq_norm3 * dq2;
  double const c[3] = {dq.x, dq.y, dq.z};
  TidalForm form;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double const identity = (i == j ? -1.0 : -0.0) / dq_norm3;
      double const square = (3 * (c[i] * c[j])) / dq_norm5;
      form.m[i][j] = identity + square;
    }
  }
  return form;
}

// form * Δv: one Dot per row, r3x3_matrix_body.hpp:327-337.
PB_HD Vec3 ApplyTidalForm(TidalForm const& f, Vec3 const& v) {
  return {f.m[0][0] * v.x + f.m[0][1] * v.y + f.m[0][2] * v.z,
          f.m[1][0] * v.x + f.m[1][1] * v.y + f.m[1][2] * v.z,
          f.m[2][0] * v.x + f.m[2][1] * v.y + f.m[2][2] * v.z};
}

// One thread per probe; the massive bodies' positions and velocities at t are
// broadcast from two small tables.
static __global__ void __launch_bounds__(128)
MasslessJerkKernel(EphemerisView const e,
                   double const* __restrict__ body_positions,
                   double const* __restrict__ body_velocities,
                   long long const n, double const* __restrict__ q,
                   double const* __restrict__ v, double* __restrict__ jerk) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  Vec3 const q1{q[i], q[n + i], q[2 * n + i]};
  Vec3 const v1{v[i], v[n + i], v[2 * n + i]};
  Vec3 sum{0.0, 0.0, 0.0};
  for (int b2 = 0; b2 < e.n_bodies; ++b2) {
    Vec3 const dq{q1.x - __ldg(body_positions + 3 * b2),
                  q1.y - __ldg(body_positions + 3 * b2 + 1),
                  q1.z - __ldg(body_positions + 3 * b2 + 2)};
    Vec3 const dv{v1.x - __ldg(body_velocities + 3 * b2),
                  v1.y - __ldg(body_velocities + 3 * b2 + 1),
                  v1.z - __ldg(body_velocities + 3 * b2 + 2)};
    Vec3 const vector = ApplyTidalForm(MakeTidalForm(dq), dv);
    double const mu2 = e.bodies[b2].mu;
    sum.x = sum.x + mu2 * vector.x;
    sum.y = sum.y + mu2 * vector.y;
    sum.z = sum.z + mu2 * vector.z;
  }
  jerk[i] = sum.x;
  jerk[n + i] = sum.y;
  jerk[2 * n + i] = sum.z;
}

// Thread b1 walks the other bodies in increasing internal order.  `jerks`
// (needs `velocities`) and `jacobians` (9 per body, row-major) may each be
// null.
static __global__ void MassiveJerkJacobianKernel(
    EphemerisView const e, double const* __restrict__ positions,
    double const* __restrict__ velocities, double* __restrict__ jerks,
    double* __restrict__ jacobians) {
  int const b1 = blockIdx.x * blockDim.x + threadIdx.x;
  if (b1 >= e.n_bodies) {
    return;
  }
  Vec3 const q1{positions[3 * b1], positions[3 * b1 + 1], positions[3 * b1 + 2]};
  Vec3 v1{0.0, 0.0, 0.0};
  if (jerks != nullptr) {
    v1 = {velocities[3 * b1], velocities[3 * b1 + 1], velocities[3 * b1 + 2]};
  }
  Vec3 jerk{0.0, 0.0, 0.0};
  TidalForm jacobian;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      jacobian.m[i][j] = 0.0;
    }
  }
  for (int b2 = 0; b2 < e.n_bodies; ++b2) {
    if (b2 == b1) {
      continue;
    }
    Vec3 const dq{q1.x - positions[3 * b2], q1.y - positions[3 * b2 + 1],
                  q1.z - positions[3 * b2 + 2]};
    TidalForm const form = MakeTidalForm(dq);
    double const mu2 = e.bodies[b2].mu;
    if (jerks != nullptr) {
      Vec3 const dv{v1.x - velocities[3 * b2], v1.y - velocities[3 * b2 + 1],
                    v1.z - velocities[3 * b2 + 2]};
      Vec3 const vector = ApplyTidalForm(form, dv);
      jerk.x = jerk.x + mu2 * vector.x;
      jerk.y = jerk.y + mu2 * vector.y;
      jerk.z = jerk.z + mu2 * vector.z;
    }
    if (jacobians != nullptr) {
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
          jacobian.m[i][j] = jacobian.m[i][j] + mu2 * form.m[i][j];
        }
      }
    }
  }
  if (jerks != nullptr) {
    jerks[3 * b1] = jerk.x;
    jerks[3 * b1 + 1] = jerk.y;
    jerks[3 * b1 + 2] = jerk.z;
  }
  if (jacobians != nullptr) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        jacobians[9 * b1 + 3 * i + j] = jacobian.m[i][j];
      }
    }
  }
}

}  // namespace pb
