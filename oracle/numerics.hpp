// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link or run anything under oracle/.
//
// CPU restatement, in plain binary64, of the arithmetic building blocks of the
// reference's gravitational-flow path.  Compile with
//   g++ -O2 -ffp-contract=off -mfma
// (reference flags: Makefile:123-169): no contraction anywhere; FMA only where
// written explicitly, which mirrors the reference's FMA-present build.
#pragma once

#include <quadmath.h>

#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <vector>

namespace orc {

// absl::StatusCode numeric values (the reference returns absl::Status).
enum StatusCode : int32_t {
  kOk = 0,
  kCancelled = 1,
  kInvalidArgument = 3,
  kDeadlineExceeded = 4,
  kResourceExhausted = 8,   // termination_condition::ReachedMaximalStepCount
  kFailedPrecondition = 9,  // termination_condition::VanishingStepSize
  kAborted = 10,
  kOutOfRange = 11,         // "Collision detected"
};

// absl::Status::Update: keeps the first error.
inline void Update(int32_t& status, int32_t const new_status) {
  if (status == kOk) {
    status = new_status;
  }
}

// ---------------------------------------------------------------------------
// numerics/double_precision_body.hpp

struct DP {
  double value = 0;
  double error = 0;
};

// :400-410, [HLB07] algorithm 4.
inline DP TwoSum(double const a, double const b) {
  DP r;
  r.value = a + b;
  double const v = r.value - a;
  r.error = (a - (r.value - v)) + (b - v);
  return r;
}

// :364-380, [HLB07] algorithm 3.
inline DP QuickTwoSum(double const a, double const b) {
  DP r;
  r.value = a + b;
  r.error = b - (r.value - a);
  return r;
}

// :412-445.  The Point x Point and Vector x Vector overloads are arithmetically
// identical (w = a - s = -(s - a) exactly), so one restatement serves both.
inline DP TwoDifference(double const a, double const b) {
  DP r;
  r.value = a - b;
  double const v = r.value - a;
  r.error = (a - (r.value - v)) - (b + v);
  return r;
}

// :486-497, [MR22] algorithm 4.
inline DP operator+(DP const& left, DP const& right) {
  DP const s = TwoSum(left.value, right.value);
  DP const t = TwoSum(left.error, right.error);
  double const c = s.error + t.value;
  DP const v = QuickTwoSum(s.value, c);
  double const w = t.error + v.error;
  return QuickTwoSum(v.value, w);
}

// :519-530, [MR22] algorithm 4.
inline DP operator-(DP const& left, DP const& right) {
  DP const s = TwoDifference(left.value, right.value);
  DP const t = TwoDifference(left.error, right.error);
  double const c = s.error + t.value;
  DP const v = QuickTwoSum(s.value, c);
  double const w = t.error + v.error;
  return QuickTwoSum(v.value, w);
}

// :264-281.
inline DP Scale(double const scale, DP const& right) {
  DP r;
  r.value = right.value * scale;
  r.error = right.error * scale;
  return r;
}

// :227-236, Higham algorithm 4.2.
inline void Increment(DP& x, double const right) {
  double const temp = x.value;
  double const y = x.error + right;
  x.value = temp + y;
  x.error = (temp - x.value) + y;
}

// ---------------------------------------------------------------------------
// Correctly-rounded elementary functions.  The reference uses its own
// correctly-rounded Sin/Cos (numerics/sin_cos.cpp, default configuration
// numerics/elementary_functions_body.hpp:36) -- a correctly-rounded result is
// mathematically unique, so any correctly-rounded implementation is
// bit-identical.  Here: binary128 libquadmath rounded once to binary64 (wrong
// only if the exact value lies within 2^-60 ulp of a rounding boundary).

inline double CrSin(double const x) { return static_cast<double>(sinq(x)); }
inline double CrCos(double const x) { return static_cast<double>(cosq(x)); }

// Root(n, x) = std::pow(x, 1.0 / n), numerics/elementary_functions_body.hpp:
// 157-159.  The reference's step sequence depends on the platform libm here
// (physics/ephemeris_test.cpp:462-474 tolerates 8 outcomes); the oracle pins
// one definition: the correctly-rounded value of x^(1/n) for the *rounded*
// exponent 1.0/n, which is what an exact pow would return.
inline double CrPow(double const x, double const y) {
  if (x == 0.0 || std::isinf(x) || std::isnan(x)) {
    return std::pow(x, y);
  }
  return static_cast<double>(powq(static_cast<__float128>(x),
                                  static_cast<__float128>(y)));
}
inline double Root(int const n, double const x) { return CrPow(x, 1.0 / n); }

// Pow<k>, Russian peasant, numerics/elementary_functions_body.hpp:171-192.
inline double PowInt(double const x, int const exponent) {
  if (exponent == 0) {
    return 1;
  } else if (exponent == 1) {
    return x;
  }
  double const y = PowInt(x, exponent / 2);
  double const y2 = y * y;
  return exponent % 2 == 1 ? y2 * x : y2;
}

// ---------------------------------------------------------------------------
// Vectors: geometry/r3_element_body.hpp.

struct Vec3 {
  double x = 0, y = 0, z = 0;
};
inline Vec3 operator+(Vec3 const& a, Vec3 const& b) {
  return {a.x + b.x, a.y + b.y, a.z + b.z};
}
inline Vec3 operator-(Vec3 const& a, Vec3 const& b) {
  return {a.x - b.x, a.y - b.y, a.z - b.z};
}
inline Vec3 operator-(Vec3 const& a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(double const s, Vec3 const& a) {
  return {s * a.x, s * a.y, s * a.z};
}
inline Vec3 operator*(Vec3 const& a, double const s) {
  return {a.x * s, a.y * s, a.z * s};
}
inline Vec3 operator/(Vec3 const& a, double const s) {
  return {a.x / s, a.y / s, a.z / s};
}
inline Vec3& operator+=(Vec3& a, Vec3 const& b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  return a;
}
inline Vec3& operator-=(Vec3& a, Vec3 const& b) {
  a.x -= b.x;
  a.y -= b.y;
  a.z -= b.z;
  return a;
}
// :171-173.
inline double Norm2(Vec3 const& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double Norm(Vec3 const& a) { return std::sqrt(Norm2(a)); }
// :596-600.
inline double Dot(Vec3 const& l, Vec3 const& r) {
  return l.x * r.x + l.y * r.y + l.z * r.z;
}
// :586-594.
inline Vec3 Cross(Vec3 const& l, Vec3 const& r) {
  return {l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x};
}

// Quaternions and rotations: geometry/quaternion_body.hpp:119-125,
// geometry/rotation_body.hpp:67-71,155-164,355-363.
struct Quaternion {
  double real = 1;
  Vec3 imaginary;
};
inline Quaternion operator*(Quaternion const& l, Quaternion const& r) {
  Quaternion q;
  q.real = l.real * r.real - Dot(l.imaginary, r.imaginary);
  q.imaginary = l.real * r.imaginary + r.real * l.imaginary +
                Cross(l.imaginary, r.imaginary);
  return q;
}
inline Quaternion AngleAxis(double const angle, Vec3 const& axis) {
  double const half_angle = 0.5 * angle;
  double const s = CrSin(half_angle);
  double const c = CrCos(half_angle);
  return {c, s * axis};
}
inline Quaternion EulerZXZ(double const alpha, double const beta,
                           double const gamma) {
  return AngleAxis(alpha, {0, 0, 1}) * AngleAxis(beta, {1, 0, 0}) *
         AngleAxis(gamma, {0, 0, 1});
}
inline Vec3 Rotate(Quaternion const& q, Vec3 const& v) {
  return v + 2 * Cross(q.imaginary, Cross(q.imaginary, v) + q.real * v);
}

// PolynomialInЧебышёвBasis::operator() and EvaluateDerivative,
// numerics/polynomial_in_чебышёв_basis_body.hpp:151-202 (Clenshaw).  The
// reference stores one_over_width_ = 1 / (upper - lower) at construction
// (:139-147).  `coefficients` has stride `stride`.
inline double ЧебышёвEvaluate(double const* coefficients, int stride,
                               int degree, double lower_bound,
                               double upper_bound, double argument) {
  double const width = upper_bound - lower_bound;
  double const one_over_width = 1 / width;
  double const scaled_argument =
      ((argument - upper_bound) + (argument - lower_bound)) * one_over_width;
  double const two_scaled_argument = scaled_argument + scaled_argument;
  double const c0 = coefficients[0];
  switch (degree) {
    case 0:
      return c0;
    case 1:
      return c0 + scaled_argument * coefficients[stride];
    default: {
      double b_k2 = coefficients[degree * stride];
      double b_k1 =
          coefficients[(degree - 1) * stride] + two_scaled_argument * b_k2;
      for (int k = degree - 2; k >= 1; --k) {
        double const b_k =
            coefficients[k * stride] + two_scaled_argument * b_k1 - b_k2;
        b_k2 = b_k1;
        b_k1 = b_k;
      }
      return c0 + scaled_argument * b_k1 - b_k2;
    }
  }
}

inline double ЧебышёвEvaluateDerivative(double const* coefficients, int stride,
                                         int degree, double lower_bound,
                                         double upper_bound, double argument) {
  double const width = upper_bound - lower_bound;
  double const one_over_width = 1 / width;
  double const scaled_argument =
      ((argument - upper_bound) + (argument - lower_bound)) * one_over_width;
  double const two_scaled_argument = scaled_argument + scaled_argument;
  double b_k2 = 0;
  double b_k1 = 0;
  for (int k = degree - 1; k >= 1; --k) {
    double const b_k = coefficients[(k + 1) * stride] * (k + 1) +
                       two_scaled_argument * b_k1 - b_k2;
    b_k2 = b_k1;
    b_k1 = b_k;
  }
  // Degree 0 has no coefficients_[1]; the reference's template is only
  // instantiated so for degree >= 1 series in practice, a constant has
  // derivative 0.
  double const c1 = degree >= 1 ? coefficients[stride] : 0.0;
  return (c1 + two_scaled_argument * b_k1 - b_k2) *
         (one_over_width + one_over_width);
}

}  // namespace orc
