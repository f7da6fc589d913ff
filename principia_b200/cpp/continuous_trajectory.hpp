// Host-side ContinuousTrajectory of the facade: equally spaced degrees of
// freedom in, piecewise polynomials (monomial basis, Newhall fit) out --
// physics/continuous_trajectory_body.hpp:82-127,742-858 and
// numerics/newhall_body.hpp:52-117,259-287 of the reference.  The polynomials
// are evaluated on the GPU (pb_ephemeris_append_segments); this class only
// produces them.  Header-only C++17, no CUDA.
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

#include "../../include/principia_b200.h"
#include "../csrc/generated_tables.h"

namespace principia_b200 {

struct DegreesOfFreedom {
  double q[3];
  double v[3];
};

struct Segment {
  double t_max;
  double origin;
  int32_t degree;
  double coefficients[18][3];
};

class ContinuousTrajectory {
 public:
  static constexpr int max_degree = 17;
  static constexpr int min_degree = 3;
  static constexpr int max_degree_age = 100;
  static constexpr int divisions = 8;

  ContinuousTrajectory(double const step, double const tolerance)
      : step_(step), tolerance_(tolerance), adjusted_tolerance_(tolerance) {}

  bool empty() const { return segments_.empty(); }
  double t_min() const { return first_time_; }
  double t_max() const {
    return segments_.empty() ? -std::numeric_limits<double>::infinity()
                             : segments_.back().t_max;
  }
  std::vector<Segment> const& segments() const { return segments_; }

  // continuous_trajectory_body.hpp:82-127.  Returns PB_OK or
  // PB_INVALID_ARGUMENT (the "apocalypse").
  pb_status_t Append(double const time, DegreesOfFreedom const& dof) {
    if (!has_first_time_) {
      first_time_ = time;
      has_first_time_ = true;
    }
    pb_status_t status = PB_OK;
    if (static_cast<int>(last_points_.size()) == divisions) {
      std::vector<DegreesOfFreedom> all;
      for (auto const& point : last_points_) {
        all.push_back(point.second);
      }
      all.push_back(dof);
      status = ComputeBestNewhallApproximation(time, all);
      last_points_.clear();
    }
    last_points_.emplace_back(time, dof);
    return status;
  }

 private:
  // Pow<k>, numerics/elementary_functions_body.hpp:171-192.
  static double PowInt(double const x, int const exponent) {
    if (exponent == 0) {
      return 1;
    } else if (exponent == 1) {
      return x;
    }
    double const y = PowInt(x, exponent / 2);
    double const y2 = y * y;
    return exponent % 2 == 1 ? y2 * x : y2;
  }

  // NewhallApproximationInMonomialBasis, newhall_body.hpp:259-287 with
  // MultiplyMatrixRowByColumnVector (:52-68) and Dehomogeneize (:93-117).
  Segment Newhall(int const degree, std::vector<DegreesOfFreedom> const& qvs,
                  double const t_min, double const t_max,
                  double (&error_estimate)[3]) const {
    double const duration_over_two = 0.5 * (t_max - t_min);
    double rq[divisions + 1][3], rv[divisions + 1][3];
    for (int i = 0, j = divisions; i < divisions + 1 && j >= 0; ++i, --j) {
      for (int c = 0; c < 3; ++c) {
        rq[j][c] = qvs[i].q[c] - 0.0;
        rv[j][c] = qvs[i].v[c] * duration_over_two;
      }
    }
    auto const row_times_column = [&](double const* const row,
                                      double (&result)[3]) {
      for (int c = 0; c < 3; ++c) {
        double sum = 0;
        for (int i = 0; i < divisions + 1; ++i) {
          sum += row[2 * i] * rq[i][c];
          sum += row[2 * i + 1] * rv[i][c];
        }
        result[c] = sum;
      }
    };
    row_times_column(&pb_newhall_chebyshev_last_row[(degree - min_degree) * 18],
                     error_estimate);
    double const* const monomial =
        &pb_newhall_monomial[pb_newhall_monomial_offset[degree - min_degree]];
    Segment s;
    s.t_max = t_max;
    s.degree = degree;
    // Barycentre({t_min, t_max}), geometry/barycentre_calculator_body.hpp:77-96.
    double total = 0.0;
    total += t_min - 0.0;
    total += t_max - 0.0;
    s.origin = 0.0 + total / 2;
    double const scale = 1.0 / duration_over_two;
    for (int k = 0; k <= degree; ++k) {
      double homogeneous[3];
      row_times_column(monomial + 18 * k, homogeneous);
      for (int c = 0; c < 3; ++c) {
        s.coefficients[k][c] = k == 0 ? homogeneous[c] + 0.0
                                      : homogeneous[c] * PowInt(scale, k);
      }
    }
    for (int k = degree + 1; k < 18; ++k) {
      s.coefficients[k][0] = s.coefficients[k][1] = s.coefficients[k][2] = 0;
    }
    return s;
  }

  static double Norm(double const (&v)[3]) {
    return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  }

  // continuous_trajectory_body.hpp:758-858.
  pb_status_t ComputeBestNewhallApproximation(
      double const time, std::vector<DegreesOfFreedom> const& all) {
    double const previous_adjusted_tolerance = adjusted_tolerance_;
    if (degree_age_ >= max_degree_age) {
      is_unstable_ = false;
      adjusted_tolerance_ = tolerance_;
      degree_ = min_degree;
      degree_age_ = 0;
    }
    double const t_min = last_points_.front().first;
    double displacement_error_estimate[3];
    segments_.push_back(
        Newhall(degree_, all, t_min, time, displacement_error_estimate));
    double error_estimate = Norm(displacement_error_estimate);
    double previous_error_estimate = error_estimate + error_estimate;
    if (is_unstable_ && error_estimate > adjusted_tolerance_) {
      is_unstable_ = false;
      adjusted_tolerance_ = tolerance_;
      degree_ = min_degree - 1;
      degree_age_ = 0;
      previous_error_estimate = std::numeric_limits<double>::max();
      error_estimate = 0.5 * previous_error_estimate;
    }
    while (error_estimate > adjusted_tolerance_ &&
           error_estimate < previous_error_estimate && degree_ < max_degree) {
      ++degree_;
      segments_.back() =
          Newhall(degree_, all, t_min, time, displacement_error_estimate);
      previous_error_estimate = error_estimate;
      error_estimate = Norm(displacement_error_estimate);
    }
    if (error_estimate >= previous_error_estimate) {
      if (degree_ > min_degree) {
        --degree_;
      }
      is_unstable_ = true;
      error_estimate = previous_error_estimate;
      adjusted_tolerance_ = std::max(adjusted_tolerance_, error_estimate);
    }
    ++degree_age_;
    return adjusted_tolerance_ < 1e6 * previous_adjusted_tolerance
               ? PB_OK
               : PB_INVALID_ARGUMENT;
  }

  double step_;
  double tolerance_;
  double adjusted_tolerance_;
  bool is_unstable_ = false;
  int degree_ = min_degree;
  int degree_age_ = 0;
  bool has_first_time_ = false;
  double first_time_ = 0;
  std::vector<std::pair<double, DegreesOfFreedom>> last_points_;
  std::vector<Segment> segments_;
};

}  // namespace principia_b200
