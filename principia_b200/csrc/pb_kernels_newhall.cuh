// ContinuousTrajectory production on the device: for a
// batch of trajectories appended in lockstep (the bodies of an ephemeris, or of
// every member of an ensemble), ContinuousTrajectory::Append +
// ComputeBestNewhallApproximation (physics/continuous_trajectory_body.hpp
// :82-127,758-858) with NewhallApproximationInMonomialBasis
// (numerics/newhall_body.hpp:52-117,259-287), and the evaluation of the result
// (:686-722).  One thread per trajectory; every array has the trajectory as its
// fastest index, so all loads and stores are coalesced.
#pragma once

#include <cuda_runtime.h>

#include "pb_math.cuh"
#include "pb_model.h"

namespace pb {

constexpr int kNewhallDivisions = 8;      // continuous_trajectory_body.hpp:39-40.
constexpr int kNewhallMinDegree = 3;      // :35-37.
constexpr int kNewhallMaxDegree = 17;
constexpr int kNewhallMaxDegreeAge = 100;  // :38.
constexpr int kNewhallRing = kNewhallDivisions + 1;

struct NewhallTables {
  double const* monomial;          // pb_newhall_monomial.
  double const* chebyshev_last_row;  // pb_newhall_chebyshev_last_row.
  int monomial_offset[15];
};

// The adaptive state of one ContinuousTrajectory (:229-247 of the header),
// SoA over the trajectories.
struct NewhallState {
  double* adjusted_tolerance;
  int32_t* degree;
  int32_t* degree_age;
  int32_t* is_unstable;
};

struct NewhallPoints {
  double const* ring;  // [slot][q xyz, v xyz][n].
  int first_slot;      // Slot of the oldest of the 9 points.
};

// MultiplyMatrixRowByColumnVector, newhall_body.hpp:52-68.
PB_HD void NewhallRowTimesColumn(double const* __restrict__ row,
                                 double const (&rq)[kNewhallRing][3],
                                 double const (&rv)[kNewhallRing][3],
                                 double (&result)[3]) {
  double sx = 0, sy = 0, sz = 0;
#pragma unroll
  for (int i = 0; i < kNewhallRing; ++i) {
#if defined(__CUDA_ARCH__)
    double2 const l = __ldg(reinterpret_cast<double2 const*>(row) + i);
    double const l0 = l.x;
    double const l1 = l.y;
#else
    double const l0 = row[2 * i];
    double const l1 = row[2 * i + 1];
#endif
    sx = sx + l0 * rq[i][0];
    sy = sy + l0 * rq[i][1];
    sz = sz + l0 * rq[i][2];
    sx = sx + l1 * rv[i][0];
    sy = sy + l1 * rv[i][1];
    sz = sz + l1 * rv[i][2];
  }
  result[0] = sx;
  result[1] = sy;
  result[2] = sz;
}

PB_HD double NewhallErrorEstimate(NewhallTables const& tables, int const degree,
                                  double const (&rq)[kNewhallRing][3],
                                  double const (&rv)[kNewhallRing][3]) {
  double e[3];
  NewhallRowTimesColumn(
      tables.chebyshev_last_row + (degree - kNewhallMinDegree) * 18, rq, rv, e);
  return sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
}

// ComputeBestNewhallApproximation for trajectory i on its last 9 points;
// writes segment `segment` ([segment][xyz][18][n] coefficients,
// [segment][n] degrees) and ORs PB_INVALID_ARGUMENT into *status on the
// "apocalypse" (:851-857).
static __global__ void __launch_bounds__(128)
NewhallFitKernel(NewhallTables const tables, NewhallState const state,
                 NewhallPoints const points, long long const n,
                 double const tolerance, double const t_min, double const t_max,
                 long long const segment, int32_t* __restrict__ degrees,
                 double* __restrict__ coefficients, int32_t* status) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  // newhall_body.hpp:272-281 with :70-91: the points with the largest time
  // first, velocities scaled by half the duration.
  double const duration_over_two = 0.5 * (t_max - t_min);
  double rq[kNewhallRing][3], rv[kNewhallRing][3];
#pragma unroll
  for (int p = 0; p < kNewhallRing; ++p) {
    int const slot = (points.first_slot + p) % kNewhallRing;
    double const* const point = points.ring + static_cast<long long>(slot) * 6 * n;
    int const j = kNewhallDivisions - p;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      rq[j][c] = point[c * n + i] - 0.0;
      rv[j][c] = point[(3 + c) * n + i] * duration_over_two;
    }
  }

  double adjusted_tolerance = state.adjusted_tolerance[i];
  int degree = state.degree[i];
  int degree_age = state.degree_age[i];
  bool is_unstable = state.is_unstable[i] != 0;

  double const previous_adjusted_tolerance = adjusted_tolerance;
  if (degree_age >= kNewhallMaxDegreeAge) {
    is_unstable = false;
    adjusted_tolerance = tolerance;
    degree = kNewhallMinDegree;
    degree_age = 0;
  }
  // The polynomial kept is the one of the last degree tried; only the error
  // estimates decide, so the coefficients are computed once, at the end.
  int fitted_degree = degree;
  double error_estimate = NewhallErrorEstimate(tables, degree, rq, rv);
  double previous_error_estimate = error_estimate + error_estimate;
  if (is_unstable && error_estimate > adjusted_tolerance) {
    is_unstable = false;
    adjusted_tolerance = tolerance;
    degree = kNewhallMinDegree - 1;
    degree_age = 0;
    previous_error_estimate = 1.7976931348623157e308;
    error_estimate = 0.5 * previous_error_estimate;
  }
  while (error_estimate > adjusted_tolerance &&
         error_estimate < previous_error_estimate &&
         degree < kNewhallMaxDegree) {
    ++degree;
    fitted_degree = degree;
    previous_error_estimate = error_estimate;
    error_estimate = NewhallErrorEstimate(tables, degree, rq, rv);
  }
  if (error_estimate >= previous_error_estimate) {
    if (degree > kNewhallMinDegree) {
      --degree;
    }
    is_unstable = true;
    error_estimate = previous_error_estimate;
    adjusted_tolerance = adjusted_tolerance > error_estimate ? adjusted_tolerance
                                                             : error_estimate;
  }
  ++degree_age;
  state.adjusted_tolerance[i] = adjusted_tolerance;
  state.degree[i] = degree;
  state.degree_age[i] = degree_age;
  state.is_unstable[i] = is_unstable ? 1 : 0;
  if (!(adjusted_tolerance < 1e6 * previous_adjusted_tolerance)) {
    atomicOr(status, 3);  // PB_INVALID_ARGUMENT.
  }

  // newhall_body.hpp:283-286 with Dehomogeneize :93-117.
  degrees[segment * n + i] = fitted_degree;
  double* const out = coefficients + segment * 3 * kSegmentCoefficients * n + i;
  double const* const monomial =
      tables.monomial + tables.monomial_offset[fitted_degree - kNewhallMinDegree];
  double const scale = 1.0 / duration_over_two;
  for (int k = 0; k < kSegmentCoefficients; ++k) {
    double c[3] = {0.0, 0.0, 0.0};
    if (k <= fitted_degree) {
      double homogeneous[3];
      NewhallRowTimesColumn(monomial + 18 * k, rq, rv, homogeneous);
      if (k == 0) {
        c[0] = homogeneous[0] + 0.0;
        c[1] = homogeneous[1] + 0.0;
        c[2] = homogeneous[2] + 0.0;
      } else {
        double const factor = PowInt(scale, k);
        c[0] = homogeneous[0] * factor;
        c[1] = homogeneous[1] * factor;
        c[2] = homogeneous[2] * factor;
      }
    }
    out[(0 * kSegmentCoefficients + k) * n] = c[0];
    out[(1 * kSegmentCoefficients + k) * n] = c[1];
    out[(2 * kSegmentCoefficients + k) * n] = c[2];
  }
}

// EvaluatePosition / EvaluateVelocity of every trajectory in segment
// `segment` (found by the host: the trajectories share their times).
static __global__ void __launch_bounds__(128)
ContinuousTrajectoriesEvaluateKernel(long long const n, long long const segment,
                                     double const x,
                                     int32_t const* __restrict__ degrees,
                                     double const* __restrict__ coefficients,
                                     double* __restrict__ positions,
                                     double* __restrict__ velocities) {
  long long const i =
      static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) {
    return;
  }
  int const degree = degrees[segment * n + i];
  double const* const in =
      coefficients + segment * 3 * kSegmentCoefficients * n + i;
  for (int c = 0; c < 3; ++c) {
    double local[kSegmentCoefficients];
    for (int k = 0; k < kSegmentCoefficients; ++k) {
      local[k] = k <= degree ? in[(c * kSegmentCoefficients + k) * n] : 0.0;
    }
    if (positions != nullptr) {
      positions[c * n + i] = EstrinEvaluate<1>(local, degree, x);
    }
    if (velocities != nullptr) {
      velocities[c * n + i] = EstrinEvaluateDerivative<1>(local, degree, x);
    }
  }
}

}  // namespace pb
