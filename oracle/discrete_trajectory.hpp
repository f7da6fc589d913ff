// TEST INFRASTRUCTURE (CPU oracle): DiscreteTrajectorySegment::Append with
// downsampling, restated from the reference:
//   physics/discrete_trajectory_segment_body.hpp:414-479 (Append), :611-625
//   (MakeHermite3), :132-155 (EvaluatePosition / EvaluateVelocity);
//   numerics/hermite3_body.hpp:66-73,154-194,237-271,289-297 (Hermite3,
//   LInfinityL₂Norm, MakePolynomial, operator-);
//   numerics/polynomial_in_monomial_basis_body.hpp:687-705,843-870 (products);
//   numerics/polynomial_evaluators_body.hpp:220-262 (Horner with FMA, the
//   default evaluator up to degree 3, polynomial_in_monomial_basis.hpp:112-115);
//   numerics/root_finders_body.hpp:459-502 (SolveQuadraticEquation).
#pragma once

#include <cmath>
#include <vector>

#include "numerics.hpp"

namespace orc {

inline Vec3 Fma(Vec3 const& a, double const x, Vec3 const& b) {
  return {std::fma(a.x, x, b.x), std::fma(a.y, x, b.y), std::fma(a.z, x, b.z)};
}

// Hermite3<Position, Instant>: the polynomial has its origin at `lower`.
struct Hermite3 {
  double lower;
  double upper;
  Vec3 a0, a1, a2, a3;
};

// Hermite3::MakePolynomial, hermite3_body.hpp:237-271.
inline Hermite3 MakeHermite3(double const t1, double const t2, Vec3 const& q1,
                             Vec3 const& q2, Vec3 const& v1, Vec3 const& v2) {
  Hermite3 h;
  h.lower = t1;
  h.upper = t2;
  h.a0 = q1;
  h.a1 = v1;
  h.a2 = {0.0, 0.0, 0.0};
  h.a3 = {0.0, 0.0, 0.0};
  double const delta_argument = t2 - t1;
  bool const values_differ = q1.x != q2.x || q1.y != q2.y || q1.z != q2.z;
  bool const derivatives_differ =
      v1.x != v2.x || v1.y != v2.y || v1.z != v2.z;
  if (delta_argument != 0.0 || values_differ || derivatives_differ) {
    double const one_over = 1.0 / delta_argument;
    double const one_over2 = one_over * one_over;
    double const one_over3 = one_over * one_over2;
    Vec3 const delta_value = q2 - q1;
    h.a2 = 3.0 * delta_value * one_over2 - (2.0 * v1 + v2) * one_over;
    h.a3 = -2.0 * delta_value * one_over3 + (v1 + v2) * one_over2;
  }
  return h;
}

// Horner with FMA, polynomial_evaluators_body.hpp:220-241.
inline Vec3 Evaluate(Hermite3 const& h, double const t) {
  double const x = t - h.lower;
  return Fma(Fma(Fma(h.a3, x, h.a2), x, h.a1), x, h.a0);
}

// :243-262: b = get<low>(coefficients) * low.
inline Vec3 EvaluateDerivative(Hermite3 const& h, double const t) {
  double const x = t - h.lower;
  return Fma(Fma(h.a3 * 3, x, h.a2 * 2), x, h.a1 * 1);
}

// operator-(Hermite3, Hermite3), hermite3_body.hpp:289-297.
inline Hermite3 Difference(Hermite3 const& left, Hermite3 const& right) {
  return {right.lower, std::min(left.upper, right.upper), left.a0 - right.a0,
          left.a1 - right.a1, left.a2 - right.a2, left.a3 - right.a3};
}

// TwoProduct with FMA, numerics/double_precision_body.hpp:349-362.
inline DP TwoProduct(double const a, double const b) {
  DP r;
  r.value = a * b;
  r.error = std::fma(a, b, -r.value);
  return r;
}

// SolveQuadraticEquation, root_finders_body.hpp:459-502; returns the number of
// roots (0, 1 or 2), sorted.
inline int SolveQuadraticEquation(double const origin, double const a0,
                                  double const a1, double const a2,
                                  double roots[2]) {
  DP const discriminant = TwoProduct(a1, a1) - TwoProduct(4.0 * a0, a2);
  if (discriminant.value == 0.0) {
    roots[0] = origin - 0.5 * a1 / a2;
    return 1;
  } else if (discriminant.value < 0.0) {
    return 0;
  } else {
    double numerator;
    if (a1 > 0.0) {
      numerator = -a1 - std::sqrt(discriminant.value);
    } else {
      numerator = -a1 + std::sqrt(discriminant.value);
    }
    double const x1 = origin + numerator / (2.0 * a2);
    double const x2 = origin + (2.0 * a0) / numerator;
    if (x1 < x2) {
      roots[0] = x1;
      roots[1] = x2;
    } else {
      roots[0] = x2;
      roots[1] = x1;
    }
    return 2;
  }
}

// Hermite3::LInfinityL₂Norm, hermite3_body.hpp:154-194, of a polynomial that
// vanishes with its derivative at `lower`.
inline double LInfinityL2Norm(Hermite3 const& h) {
  // q(t) = p(t) / (t - lower)² = {a2, a3}; monomial = {0, 0.5}.
  Vec3 const q0 = h.a2;
  Vec3 const q1 = h.a3;
  // q.Derivative(): FallingFactorial(1, 1) * a3.
  Vec3 const q_derivative0 = 1 * q1;
  // monomial * q.Derivative(): result[i + j] += l_i * r_j from zero.
  Vec3 const zero{0.0, 0.0, 0.0};
  Vec3 product0 = zero;
  Vec3 product1 = zero;
  product0 += 0.0 * q_derivative0;
  product1 += 0.5 * q_derivative0;
  // q + that.
  Vec3 const r0 = q0 + product0;
  Vec3 const r1 = q1 + product1;
  // PointwiseInnerProduct(q, r): result[i + j] += <l_i, r_j> from zero.
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
  c0 += Dot(q0, r0);
  c1 += Dot(q0, r1);
  c1 += Dot(q1, r0);
  c2 += Dot(q1, r1);
  double roots[2];
  int const n_roots = SolveQuadraticEquation(h.lower, c0, c1, c2, roots);
  double max = Norm2(Evaluate(h, h.upper));
  for (int i = 0; i < n_roots; ++i) {
    if (h.lower < roots[i] && roots[i] < h.upper) {
      double const candidate = Norm2(Evaluate(h, roots[i]));
      max = (max < candidate) ? candidate : max;  // std::max.
    }
  }
  return std::sqrt(max);
}

// One point of the timeline with the interpolation that ends at it.
struct TimelinePoint {
  double time;
  Vec3 q, v;
  bool has_interpolation = false;
  Hermite3 hermite3{};
  double error = 0.0;
};

// DiscreteTrajectorySegment with SetDownsampling({tolerance}) (a negative
// tolerance: no downsampling).
class DownsamplingSegment {
 public:
  explicit DownsamplingSegment(double const tolerance)
      : tolerance_(tolerance) {}

  // discrete_trajectory_segment_body.hpp:414-479.
  void Append(double const t, Vec3 const& q, Vec3 const& v) {
    if (timeline_.empty()) {
      timeline_.push_back({t, q, v});
      return;
    }
    if (timeline_.front().time == t) {
      return;  // "Append at existing time".
    }
    if (timeline_.size() > 1 && tolerance_ >= 0.0) {
      TimelinePoint const& lower = timeline_[timeline_.size() - 2];
      TimelinePoint& last = timeline_.back();
      Hermite3 const hermite3 =
          MakeHermite3(lower.time, t, lower.q, q, lower.v, v);
      Hermite3 const difference = Difference(hermite3, last.hermite3);
      downsampling_error_ += LInfinityL2Norm(difference);
      if (downsampling_error_ < tolerance_) {
        last.time = t;
        last.q = q;
        last.v = v;
        last.hermite3 = hermite3;
        last.error = downsampling_error_;
        return;
      }
    }
    TimelinePoint const& lower = timeline_.back();
    TimelinePoint point{t, q, v};
    point.has_interpolation = true;
    point.hermite3 = MakeHermite3(lower.time, t, lower.q, q, lower.v, v);
    downsampling_error_ = 0.0;
    point.error = downsampling_error_;
    timeline_.push_back(point);
  }

  std::vector<TimelinePoint> const& timeline() const { return timeline_; }

  // EvaluatePosition / EvaluateVelocity, :132-155: the interpolation of the
  // first point at or after t.
  bool Evaluate(double const t, Vec3& q, Vec3& v) const {
    for (TimelinePoint const& point : timeline_) {
      if (t <= point.time) {
        if (point.time == t) {
          q = point.q;
          v = point.v;
          return true;
        }
        if (!point.has_interpolation) {
          return false;
        }
        q = orc::Evaluate(point.hermite3, t);
        v = orc::EvaluateDerivative(point.hermite3, t);
        return true;
      }
    }
    return false;
  }

 private:
  double tolerance_;
  double downsampling_error_ = 0.0;
  std::vector<TimelinePoint> timeline_;
};

}  // namespace orc
