// ORACLE -- TEST INFRASTRUCTURE ONLY (see numerics.hpp).
//
// CPU restatement of the reference's physics on the hot path: bodies,
// harmonic damping, geopotential, ContinuousTrajectory (Newhall fit + Estrin
// evaluation) and the Ephemeris acceleration / flow routines.
#pragma once

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <queue>
#include <string>
#include <functional>
#include <vector>

#include "generated_tables.h"
#include "integrators.hpp"
#include "numerics.hpp"

namespace orc {

constexpr double kPi = 3.14159265358979323846264338327950288419716939937511;

inline double LegendreNormalizationFactor(int n, int m) {
  return pb_legendre_normalization_factor[n * (n + 1) / 2 + m];
}
inline double MaxAbsNormalizedAssociatedLegendreFunction(int n, int m) {
  return pb_max_abs_normalized_legendre[n * (n + 1) / 2 + m];
}

// ---------------------------------------------------------------------------
// physics/massive_body_body.hpp, rotating_body_body.hpp:65-81,156-160,
// rotating_body.hpp:181-191, oblate_body_body.hpp.

struct Body {
  std::string name;
  double mu = 0;
  bool rotating = false;
  double min_radius = 0;  // MassiveBody::min_radius() is 0.
  double mean_radius = 0;
  double reference_angle = 0, reference_instant = 0, angular_frequency = 0;
  double right_ascension = 0, declination = 0;
  bool oblate = false;
  int degree = 0;
  bool zonal = false;
  double reference_radius = 0;
  std::vector<double> cos, sin;  // (n, m) -> n(n+1)/2 + m, normalised.
  Vec3 polar_axis, biequatorial, equatorial;

  double Cos(int n, int m) const { return cos[n * (n + 1) / 2 + m]; }
  double Sin(int n, int m) const { return sin[n * (n + 1) / 2 + m]; }

  // geometry/r3_element_body.hpp:213-219 (SphericalCoordinates::ToCartesian).
  static Vec3 ToCartesian(double radius, double latitude, double longitude) {
    double const sin_latitude = CrSin(latitude);
    double const cos_latitude = CrCos(latitude);
    double const sin_longitude = CrSin(longitude);
    double const cos_longitude = CrCos(longitude);
    return {radius * cos_longitude * cos_latitude,
            radius * sin_longitude * cos_latitude,
            radius * sin_latitude};
  }

  void ComputeAxes() {
    if (!rotating) {
      return;
    }
    polar_axis = ToCartesian(1.0, declination, right_ascension);
    biequatorial = ToCartesian(1.0, 0.0, kPi / 2 + right_ascension);
    equatorial = Cross(biequatorial, polar_axis);
  }

  double AngleAt(double const t) const {
    return reference_angle + (t - reference_instant) * angular_frequency;
  }

  Quaternion FromSurfaceFrame(double const t) const {
    return EulerZXZ(kPi / 2 + right_ascension, kPi / 2 - declination,
                    AngleAt(t));
  }
};

// ---------------------------------------------------------------------------
// physics/harmonic_damping_body.hpp:16-60.

struct HarmonicDamping {
  double outer_threshold = std::numeric_limits<double>::infinity();
  double inner_threshold = std::numeric_limits<double>::infinity();
  double c1 = 0, c2 = 0, c3 = 0;

  HarmonicDamping() = default;
  explicit HarmonicDamping(double const s0)
      : outer_threshold(s0 * 3),
        inner_threshold(s0),
        c1(9 / (4 * s0)),
        c2(-3 / (2 * PowInt(s0, 2))),
        c3(1 / (4 * PowInt(s0, 3))) {}

  void ComputeDampedRadialQuantities(double r_norm, double r2,
                                     Vec3 const& r_normalized,
                                     double R_over_r, double R_prime,
                                     double& sigma_R_over_r,
                                     Vec3& grad_sigma_R) const {
    if (r_norm <= inner_threshold) {
      sigma_R_over_r = R_over_r;
      grad_sigma_R = R_prime * r_normalized;
    } else {
      double const r3 = r2 * r_norm;
      double const c3r3 = c3 * r3;
      double const c2r2 = c2 * r2;
      double const c1r = c1 * r_norm;
      double const sigma = c3r3 + c2r2 + c1r;
      double const sigma_prime_r = 3 * c3r3 + 2 * c2r2 + c1r;
      sigma_R_over_r = sigma * R_over_r;
      grad_sigma_R = (sigma_prime_r * R_over_r + R_prime * sigma) * r_normalized;
    }
  }

  void ComputeDampedRadialQuantities(double r_norm, double r2, double R_over_r,
                                     double& sigma_R_over_r) const {
    if (r_norm <= inner_threshold) {
      sigma_R_over_r = R_over_r;
    } else {
      double const r3 = r2 * r_norm;
      double const c3r3 = c3 * r3;
      double const c2r2 = c2 * r2;
      double const c1r = c1 * r_norm;
      double const sigma = c3r3 + c2r2 + c1r;
      sigma_R_over_r = sigma * R_over_r;
    }
  }
};

// ---------------------------------------------------------------------------
// physics/geopotential_body.hpp.

class Geopotential {
 public:
  // :669-733.
  Geopotential(Body const* body, double const tolerance) : body_(body) {
    double const epsilon = tolerance;
    struct Threshold {
      double r;
      int n;
      int m;
    };
    auto const after = [](Threshold const& left, Threshold const& right) {
      return left.r < right.r ||
             (left.r == right.r &&
              (left.m < right.m || (left.m == right.m && left.n < right.n)));
    };
    std::priority_queue<Threshold, std::vector<Threshold>, decltype(after)>
        harmonic_thresholds(after);
    for (int n = 2; n <= body_->degree; ++n) {
      for (int m = 0; m <= n; ++m) {
        double const max_abs_Pnm =
            MaxAbsNormalizedAssociatedLegendreFunction(n, m);
        double const Cnm = body_->Cos(n, m);
        double const Snm = body_->Sin(n, m);
        // Root(n, x) = std::pow(x, 1.0 / n) with the platform libm
        // (numerics/elementary_functions_body.hpp:157-159).
        double const r =
            Cnm == 0 && Snm == 0
                ? 0.0
                : body_->reference_radius *
                      std::pow((max_abs_Pnm * (n + 1) *
                                std::sqrt(PowInt(Cnm, 2) + PowInt(Snm, 2))) /
                                   epsilon,
                               1.0 / n);
        harmonic_thresholds.push({r, n, m});
      }
    }
    double const infinity = std::numeric_limits<double>::infinity();
    harmonic_thresholds.push({infinity, 0, 0});
    harmonic_thresholds.push({infinity, 1, 0});

    while (!harmonic_thresholds.empty()) {
      Threshold const threshold = harmonic_thresholds.top();
      if (threshold.n == 2 && threshold.m == 2) {
        if (degree_damping_.size() > 3) {
          sectoral_damping_ = HarmonicDamping(degree_damping_[3].inner_threshold);
        } else {
          sectoral_damping_ = HarmonicDamping(threshold.r);
        }
      }
      while (threshold.n >= static_cast<int>(degree_damping_.size())) {
        degree_damping_.emplace_back(threshold.r);
      }
      harmonic_thresholds.pop();
    }
  }

  std::vector<HarmonicDamping> const& degree_damping() const {
    return degree_damping_;
  }
  HarmonicDamping const& sectoral_damping() const { return sectoral_damping_; }

  // :910-922.
  int LimitingDegree(double const r_norm) const {
    if (!std::isfinite(r_norm)) {
      return 2;
    }
    return static_cast<int>(
        std::partition_point(degree_damping_.begin(), degree_damping_.end(),
                             [r_norm](HarmonicDamping const& d) {
                               return r_norm < d.outer_threshold;
                             }) -
        degree_damping_.begin());
  }

  // :740-813.  Result in m^-2 (multiply by mu for an acceleration).
  Vec3 GeneralSphericalHarmonicsAcceleration(double t, Vec3 const& r,
                                             double r_norm, double r2,
                                             double one_over_r3) const {
    if (r_norm != r_norm) {
      double const nan = std::numeric_limits<double>::quiet_NaN();
      return nan * Vec3{};
    }
    int const max_degree = LimitingDegree(r_norm) - 1;
    if (max_degree == 1) {
      return {};
    }
    return AllDegreesAcceleration(max_degree, t, r, r_norm, r2, one_over_r3);
  }

  // :822-895.  Result in m^-1.
  double GeneralSphericalHarmonicsPotential(double t, Vec3 const& r,
                                            double r_norm, double r2,
                                            double one_over_r3) const {
    if (r_norm != r_norm) {
      return std::numeric_limits<double>::quiet_NaN();
    }
    int const max_degree = LimitingDegree(r_norm) - 1;
    if (max_degree == 1) {
      return 0.0;
    }
    return AllDegreesPotential(max_degree, t, r, r_norm, r2, one_over_r3);
  }

 private:
  static constexpr int kSize = 51;

  struct Precomputations {
    double r_norm;
    double r2;
    Vec3 r_normalized;
    double sin_beta;
    double cos_beta;
    Vec3 grad_B_vector;
    Vec3 grad_L_vector;
    double R_over_r[kSize];
    double cos_m_lambda[kSize];
    double sin_m_lambda[kSize];
    double cos_beta_to_the_m[kSize];
    double DmPn[kSize * (kSize + 1) / 2];
    double& P(int n, int m) { return DmPn[n * (n + 1) / 2 + m]; }
  };

  // :575-667.
  void InitializePrecomputations(double t, Vec3 const& r, double r_norm,
                                 double r2, double one_over_r3,
                                 Precomputations& p) const {
    Body const& body = *body_;
    bool const is_zonal =
        body.zonal || r_norm > sectoral_damping_.outer_threshold;
    p.r_norm = r_norm;
    p.r2 = r2;

    Vec3 x_hat, y_hat;
    Vec3 const z_hat = body.polar_axis;
    if (is_zonal) {
      x_hat = body.equatorial;
      y_hat = body.biequatorial;
    } else {
      Quaternion const from_surface_frame = body.FromSurfaceFrame(t);
      x_hat = Rotate(from_surface_frame, {1, 0, 0});
      y_hat = Rotate(from_surface_frame, {0, 1, 0});
    }

    double const x = Dot(r, x_hat);
    double const y = Dot(r, y_hat);
    double const z = Dot(r, z_hat);

    double const x2_plus_y2 = x * x + y * y;
    double const r_equatorial = std::sqrt(x2_plus_y2);

    double cos_lambda = 1;
    double sin_lambda = 0;
    if (r_equatorial > 0) {
      double const one_over_r_equatorial = 1 / r_equatorial;
      cos_lambda = x * one_over_r_equatorial;
      sin_lambda = y * one_over_r_equatorial;
    }

    double const one_over_r_norm = 1 / r_norm;
    p.r_normalized = r * one_over_r_norm;

    p.cos_beta = r_equatorial * one_over_r_norm;
    p.sin_beta = z * one_over_r_norm;

    p.grad_B_vector = (-p.sin_beta * cos_lambda) * x_hat -
                      (p.sin_beta * sin_lambda) * y_hat + p.cos_beta * z_hat;
    p.grad_L_vector = cos_lambda * y_hat - sin_lambda * x_hat;

    p.R_over_r[1] = body.reference_radius * one_over_r3;

    p.cos_m_lambda[1] = cos_lambda;
    p.sin_m_lambda[1] = sin_lambda;

    p.cos_beta_to_the_m[0] = 1;
    p.cos_beta_to_the_m[1] = p.cos_beta;

    p.P(0, 0) = 1;
    p.P(1, 0) = p.sin_beta;
    p.P(1, 1) = 1;
  }

  // DegreeNOrderM::UpdatePrecomputations, :249-331.
  static void UpdateOrder(int n, int m, Precomputations& p) {
    if (n == 2 && m == 1) {
      p.P(2, 2) = 3;
      return;
    }
    double const sin_beta = p.sin_beta;
    if (m == n) {
      if (m % 2 == 0) {
        int const h = m / 2;
        double const cos_h = p.cos_m_lambda[h];
        double const sin_h = p.sin_m_lambda[h];
        double const cos_beta_to_the_h = p.cos_beta_to_the_m[h];
        p.sin_m_lambda[m] = 2 * sin_h * cos_h;
        p.cos_m_lambda[m] = (cos_h + sin_h) * (cos_h - sin_h);
        p.cos_beta_to_the_m[m] = cos_beta_to_the_h * cos_beta_to_the_h;
      } else {
        int const h1 = m / 2;
        int const h2 = m - h1;
        double const cos_h1 = p.cos_m_lambda[h1];
        double const sin_h1 = p.sin_m_lambda[h1];
        double const cos_beta_to_the_h1 = p.cos_beta_to_the_m[h1];
        double const cos_h2 = p.cos_m_lambda[h2];
        double const sin_h2 = p.sin_m_lambda[h2];
        double const cos_beta_to_the_h2 = p.cos_beta_to_the_m[h2];
        p.sin_m_lambda[m] = sin_h1 * cos_h2 + cos_h1 * sin_h2;
        p.cos_m_lambda[m] = cos_h1 * cos_h2 - sin_h1 * sin_h2;
        p.cos_beta_to_the_m[m] = cos_beta_to_the_h1 * cos_beta_to_the_h2;
      }
    }
    if (m == 0) {
      p.P(n, 0) = ((2 * n - 1) * sin_beta * p.P(n - 1, 0) -
                   (n - 1) * p.P(n - 2, 0)) /
                  n;
    }
    if (m == n) {
    } else if (m == n - 1) {
      p.P(n, m + 1) = ((2 * n - 1) * (m + 1) * p.P(n - 1, m)) / n;
    } else if (m == n - 2) {
      p.P(n, m + 1) = ((2 * n - 1) * (sin_beta * p.P(n - 1, m + 1) +
                                      (m + 1) * p.P(n - 1, m))) /
                      n;
    } else {
      p.P(n, m + 1) = ((2 * n - 1) * (sin_beta * p.P(n - 1, m + 1) +
                                      (m + 1) * p.P(n - 1, m)) -
                       (n - 1) * p.P(n - 2, m + 1)) /
                      n;
    }
  }

  // DegreeNOrderM::Acceleration, :130-204.
  Vec3 OrderAcceleration(int n, int m, double sigma_R_over_r,
                         Vec3 const& grad_sigma_R, Precomputations& p) const {
    UpdateOrder(n, m, p);
    if (n == 2 && m == 1) {
      return {};
    }
    double const cos_beta = p.cos_beta;
    double const sin_beta = p.sin_beta;
    double const cos_m_lambda = p.cos_m_lambda[m];
    double const sin_m_lambda = p.sin_m_lambda[m];
    double const cos_beta_to_the_m = p.cos_beta_to_the_m[m];
    double const normalization_factor = LegendreNormalizationFactor(n, m);

    double cos_beta_to_the_m_minus_1 = 0;
    double const B = cos_beta_to_the_m * p.P(n, m);

    double grad_B_polynomials = 0;
    if (m < n) {
      grad_B_polynomials = cos_beta * cos_beta_to_the_m * p.P(n, m + 1);
    }
    if (m > 0) {
      cos_beta_to_the_m_minus_1 = p.cos_beta_to_the_m[m - 1];
      grad_B_polynomials -= m * sin_beta * cos_beta_to_the_m_minus_1 * p.P(n, m);
    }

    double const Cnm = body_->Cos(n, m);
    double const Snm = body_->Sin(n, m);
    double L;
    if (m == 0) {
      L = Cnm;
    } else {
      L = Cnm * cos_m_lambda + Snm * sin_m_lambda;
    }

    Vec3 const BL_grad_R = (B * L) * grad_sigma_R;
    Vec3 const RL_grad_B =
        (sigma_R_over_r * L * grad_B_polynomials) * p.grad_B_vector;
    Vec3 grad_RBL = BL_grad_R + RL_grad_B;
    if (m > 0) {
      Vec3 const RB_grad_L =
          (sigma_R_over_r * cos_beta_to_the_m_minus_1 * p.P(n, m) * m *
           (Snm * cos_m_lambda - Cnm * sin_m_lambda)) *
          p.grad_L_vector;
      grad_RBL += RB_grad_L;
    }
    return normalization_factor * grad_RBL;
  }

  // DegreeNOrderM::Potential, :206-247.
  double OrderPotential(int n, int m, double sigma_R_over_r,
                        Precomputations& p) const {
    UpdateOrder(n, m, p);
    if (n == 2 && m == 1) {
      return 0.0;
    }
    double const cos_m_lambda = p.cos_m_lambda[m];
    double const sin_m_lambda = p.sin_m_lambda[m];
    double const cos_beta_to_the_m = p.cos_beta_to_the_m[m];
    double const normalization_factor = LegendreNormalizationFactor(n, m);
    double const sigma_R = p.r_norm * sigma_R_over_r;
    double const B = cos_beta_to_the_m * p.P(n, m);
    double const Cnm = body_->Cos(n, m);
    double const Snm = body_->Sin(n, m);
    double L;
    if (m == 0) {
      L = Cnm;
    } else {
      L = Cnm * cos_m_lambda + Snm * sin_m_lambda;
    }
    return -normalization_factor * sigma_R * B * L;
  }

  // DegreeNAllOrders::UpdatePrecomputations, :476-499.
  static void UpdateDegree(int n, Precomputations& p) {
    if (n % 2 == 0) {
      int const h = n / 2;
      p.R_over_r[n] = p.R_over_r[h] * p.R_over_r[h] * p.r2;
    } else {
      int const h1 = n / 2;
      int const h2 = n - h1;
      p.R_over_r[n] = p.R_over_r[h1] * p.R_over_r[h2] * p.r2;
    }
  }

  // DegreeNAllOrders::Acceleration, :333-413.  `size` is the number of orders
  // (1 when zonal, n + 1 otherwise).
  Vec3 DegreeAcceleration(int n, int size, Precomputations& p) const {
    if (n < 2) {
      return {};
    }
    UpdateDegree(n, p);
    double const R_over_r = p.R_over_r[n];
    double const R_prime = -(n + 1) * R_over_r;

    double sigma_R_over_r;
    Vec3 grad_sigma_R;
    if (n == 2 && size > 1) {
      degree_damping_[2].ComputeDampedRadialQuantities(
          p.r_norm, p.r2, p.r_normalized, R_over_r, R_prime, sigma_R_over_r,
          grad_sigma_R);
      Vec3 const j2_acceleration =
          OrderAcceleration(2, 0, sigma_R_over_r, grad_sigma_R, p);
      sectoral_damping_.ComputeDampedRadialQuantities(
          p.r_norm, p.r2, p.r_normalized, R_over_r, R_prime, sigma_R_over_r,
          grad_sigma_R);
      OrderAcceleration(2, 1, sigma_R_over_r, grad_sigma_R, p);
      Vec3 const c22_s22_acceleration =
          OrderAcceleration(2, 2, sigma_R_over_r, grad_sigma_R, p);
      return j2_acceleration + c22_s22_acceleration;
    } else {
      degree_damping_[n].ComputeDampedRadialQuantities(
          p.r_norm, p.r2, p.r_normalized, R_over_r, R_prime, sigma_R_over_r,
          grad_sigma_R);
      // Evaluated by increasing order (initializer list), summed by a unary
      // right fold `(accelerations[orders] + ...)`: a0 + (a1 + (... + a_last)).
      Vec3 accelerations[kSize];
      for (int m = 0; m < size; ++m) {
        accelerations[m] =
            OrderAcceleration(n, m, sigma_R_over_r, grad_sigma_R, p);
      }
      Vec3 sum = accelerations[size - 1];
      for (int m = size - 2; m >= 0; --m) {
        sum = accelerations[m] + sum;
      }
      return sum;
    }
  }

  double DegreePotential(int n, int size, Precomputations& p) const {
    if (n < 2) {
      return 0.0;
    }
    UpdateDegree(n, p);
    double const R_over_r = p.R_over_r[n];
    double sigma_R_over_r;
    if (n == 2 && size > 1) {
      degree_damping_[2].ComputeDampedRadialQuantities(p.r_norm, p.r2, R_over_r,
                                                       sigma_R_over_r);
      double const j2_potential = OrderPotential(2, 0, sigma_R_over_r, p);
      sectoral_damping_.ComputeDampedRadialQuantities(p.r_norm, p.r2, R_over_r,
                                                      sigma_R_over_r);
      OrderPotential(2, 1, sigma_R_over_r, p);
      double const c22_s22_potential = OrderPotential(2, 2, sigma_R_over_r, p);
      return j2_potential + c22_s22_potential;
    } else {
      degree_damping_[n].ComputeDampedRadialQuantities(p.r_norm, p.r2, R_over_r,
                                                       sigma_R_over_r);
      double potentials[kSize] = {};
      for (int m = 0; m < size; ++m) {
        potentials[m] = OrderPotential(n, m, sigma_R_over_r, p);
      }
      double sum = potentials[size - 1];
      for (int m = size - 2; m >= 0; --m) {
        sum = potentials[m] + sum;
      }
      return sum;
    }
  }

  // AllDegrees::Acceleration, :501-536: degrees 0..max_degree evaluated by
  // increasing degree, summed by a unary right fold.
  Vec3 AllDegreesAcceleration(int max_degree, double t, Vec3 const& r,
                              double r_norm, double r2,
                              double one_over_r3) const {
    bool const is_zonal =
        body_->zonal || r_norm > sectoral_damping_.outer_threshold;
    Precomputations p;
    InitializePrecomputations(t, r, r_norm, r2, one_over_r3, p);
    Vec3 accelerations[kSize];
    for (int n = 0; n <= max_degree; ++n) {
      accelerations[n] = DegreeAcceleration(n, is_zonal ? 1 : n + 1, p);
    }
    Vec3 sum = accelerations[max_degree];
    for (int n = max_degree - 1; n >= 0; --n) {
      sum = accelerations[n] + sum;
    }
    return sum;
  }

  double AllDegreesPotential(int max_degree, double t, Vec3 const& r,
                             double r_norm, double r2,
                             double one_over_r3) const {
    bool const is_zonal =
        body_->zonal || r_norm > sectoral_damping_.outer_threshold;
    Precomputations p;
    InitializePrecomputations(t, r, r_norm, r2, one_over_r3, p);
    double potentials[kSize] = {};
    for (int n = 0; n <= max_degree; ++n) {
      potentials[n] = DegreePotential(n, is_zonal ? 1 : n + 1, p);
    }
    double sum = potentials[max_degree];
    for (int n = max_degree - 1; n >= 0; --n) {
      sum = potentials[n] + sum;
    }
    return sum;
  }

  Body const* body_;
  std::vector<HarmonicDamping> degree_damping_;
  HarmonicDamping sectoral_damping_;
};

// ---------------------------------------------------------------------------
// physics/continuous_trajectory_body.hpp + numerics/newhall_body.hpp +
// numerics/polynomial_evaluators_body.hpp (Estrin with FMA).

struct Segment {
  double t_max;
  double origin;  // t_mid.
  int degree;
  double c[18][3];  // Coefficient k of (t - origin)^k, xyz.
};

// InternalEstrinEvaluator<..., fma = true>::EvaluateDerivative,
// numerics/polynomial_evaluators_body.hpp:104-129,162-192, entered with low = 1,
// subdegree = degree - 1 (:304-330).
inline double EstrinEvaluateDerivative(double const* c, int stride, int low,
                                       int subdegree, double x) {
  if (subdegree == -1) {
    return 0.0;
  }
  if (subdegree == 0) {
    return low * c[low * stride];
  }
  if (subdegree == 1) {
    return std::fma((low + 1) * c[(low + 1) * stride], x,
                    low * c[low * stride]);
  }
  int m = 1;
  while (m * 2 <= subdegree) {
    m *= 2;
  }
  double const a =
      EstrinEvaluateDerivative(c, stride, low + m, subdegree - m, x);
  double const b = EstrinEvaluateDerivative(c, stride, low, m - 1, x);
  double xm = x;
  for (int k = 1; k < m; k *= 2) {
    xm = xm * xm;
  }
  return std::fma(a, xm, b);
}

// InternalEstrinEvaluator<..., fma = true>::Evaluate, :72-160.
inline double EstrinEvaluate(double const* c, int stride, int low,
                             int subdegree, double x) {
  if (subdegree == 0) {
    return c[low * stride];
  }
  if (subdegree == 1) {
    return std::fma(c[(low + 1) * stride], x, c[low * stride]);
  }
  int m = 1;
  while (m * 2 <= subdegree) {
    m *= 2;
  }
  double const a = EstrinEvaluate(c, stride, low + m, subdegree - m, x);
  double const b = EstrinEvaluate(c, stride, low, m - 1, x);
  double const xm = PowInt(x, m);
  return std::fma(a, xm, b);
}

class ContinuousTrajectory {
 public:
  static constexpr int max_degree = 17;
  static constexpr int min_degree = 3;
  static constexpr int max_degree_age = 100;
  static constexpr int divisions = 8;

  ContinuousTrajectory(double step, double tolerance)
      : step_(step), tolerance_(tolerance), adjusted_tolerance_(tolerance) {}

  std::vector<Segment> const& segments() const { return segments_; }
  bool empty() const { return segments_.empty(); }
  double t_min() const { return first_time_; }
  double t_max() const {
    return segments_.empty() ? -std::numeric_limits<double>::infinity()
                             : segments_.back().t_max;
  }

  void AppendSegment(Segment const& s, double t_min_if_first) {
    if (segments_.empty()) {
      first_time_ = t_min_if_first;
    }
    segments_.push_back(s);
  }

  // :82-127.
  int32_t Append(double time, Vec3 const& q, Vec3 const& v) {
    if (!has_first_time_) {
      first_time_ = time;
      has_first_time_ = true;
    }
    int32_t status = kOk;
    if (static_cast<int>(last_points_.size()) == divisions) {
      std::vector<std::pair<Vec3, Vec3>> all;
      for (auto const& p : last_points_) {
        all.push_back({p.q, p.v});
      }
      all.push_back({q, v});
      status = ComputeBestNewhallApproximation(time, all);
      last_points_.clear();
    }
    last_points_.push_back({time, q, v});
    return status;
  }

  // EvaluateVelocity / EvaluateDegreesOfFreedom: the derivative of the same
  // polynomial, continuous_trajectory_body.hpp:697-722.
  Vec3 EvaluateVelocity(double const time) const {
    auto const it = std::lower_bound(
        segments_.begin(), segments_.end(), time,
        [](Segment const& left, double right) { return left.t_max < right; });
    Segment const& s = *it;
    double const x = time - s.origin;
    return {EstrinEvaluateDerivative(&s.c[0][0], 3, 1, s.degree - 1, x),
            EstrinEvaluateDerivative(&s.c[0][1], 3, 1, s.degree - 1, x),
            EstrinEvaluateDerivative(&s.c[0][2], 3, 1, s.degree - 1, x)};
  }

  // :686-695, :860-885 (first polynomial with time <= t_max).
  Vec3 EvaluatePosition(double const time) const {
    auto const it = std::lower_bound(
        segments_.begin(), segments_.end(), time,
        [](Segment const& left, double right) { return left.t_max < right; });
    Segment const& s = *it;
    double const x = time - s.origin;
    return {EstrinEvaluate(&s.c[0][0], 3, 0, s.degree, x),
            EstrinEvaluate(&s.c[0][1], 3, 0, s.degree, x),
            EstrinEvaluate(&s.c[0][2], 3, 0, s.degree, x)};
  }

 private:
  struct Point {
    double time;
    Vec3 q, v;
  };

 public:
  // numerics/newhall_body.hpp:52-117,259-287.
  Segment Newhall(int degree, std::vector<std::pair<Vec3, Vec3>> const& qvs,
                  double t_min, double t_max, Vec3& error_estimate) const {
    double const duration_over_two = 0.5 * (t_max - t_min);
    // Largest time first.
    Vec3 rq[divisions + 1], rv[divisions + 1];
    for (int i = 0, j = divisions; i < divisions + 1 && j >= 0; ++i, --j) {
      rq[j] = qvs[i].first - Vec3{};
      rv[j] = qvs[i].second * duration_over_two;
    }
    auto const row_times_column = [&](double const* row) {
      Vec3 result{};
      for (int i = 0; i < divisions + 1; ++i) {
        double const l0 = row[2 * i];
        double const l1 = row[2 * i + 1];
        result += l0 * rq[i];
        result += l1 * rv[i];
      }
      return result;
    };
    error_estimate = row_times_column(
        &pb_newhall_chebyshev_last_row[(degree - min_degree) * 18]);
    double const* monomial =
        &pb_newhall_monomial[pb_newhall_monomial_offset[degree - min_degree]];
    Segment s;
    s.t_max = t_max;
    s.degree = degree;
    s.origin = Barycentre2(t_min, t_max);
    double const scale = 1.0 / duration_over_two;
    for (int k = 0; k <= degree; ++k) {
      Vec3 const homogeneous = row_times_column(monomial + 18 * k);
      Vec3 c;
      if (k == 0) {
        c = homogeneous + Vec3{};
      } else {
        c = homogeneous * PowInt(scale, k);
      }
      s.c[k][0] = c.x;
      s.c[k][1] = c.y;
      s.c[k][2] = c.z;
    }
    for (int k = degree + 1; k < 18; ++k) {
      s.c[k][0] = s.c[k][1] = s.c[k][2] = 0;
    }
    return s;
  }

  // Barycentre({t_min, t_max}), geometry/barycentre_calculator_body.hpp:77-96
  // (unweighted overload; Instant is affine, origin = Instant{}).
  static double Barycentre2(double const t_min, double const t_max) {
    double total = 0.0;
    total += t_min - 0.0;
    total += t_max - 0.0;
    return 0.0 + total / 2;
  }

  // :758-858.
  int32_t ComputeBestNewhallApproximation(
      double time, std::vector<std::pair<Vec3, Vec3>> const& all) {
    double const previous_adjusted_tolerance = adjusted_tolerance_;
    if (degree_age_ >= max_degree_age) {
      is_unstable_ = false;
      adjusted_tolerance_ = tolerance_;
      degree_ = min_degree;
      degree_age_ = 0;
    }
    double const t_min = last_points_.front().time;
    Vec3 displacement_error_estimate;
    // The reference's tests replace the fit by a mock that scripts the error
    // estimates (continuous_trajectory_test.cpp:60-110).
    auto const Newhall = [this](int const degree,
                                std::vector<std::pair<Vec3, Vec3>> const& qvs,
                                double const t_min, double const t_max,
                                Vec3& error_estimate) {
      if (mock_error_estimate) {
        error_estimate = mock_error_estimate(degree);
        Segment s{};
        s.t_max = t_max;
        s.degree = degree;
        return s;
      }
      return this->Newhall(degree, qvs, t_min, t_max, error_estimate);
    };
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

    if (adjusted_tolerance_ < 1e6 * previous_adjusted_tolerance) {
      return kOk;
    } else {
      return kInvalidArgument;
    }
  }

  // Testing hooks (TestableContinuousTrajectory of the reference's tests).
  std::function<Vec3(int degree)> mock_error_estimate;
  int degree() const { return degree_; }
  double adjusted_tolerance() const { return adjusted_tolerance_; }
  bool is_unstable() const { return is_unstable_; }
  void ResetBestNewhallApproximation() {
    degree_age_ = std::numeric_limits<int>::max();
  }

 private:
  double step_;
  double tolerance_;
  double adjusted_tolerance_;
  bool is_unstable_ = false;
  int degree_ = min_degree;
  int degree_age_ = 0;
  bool has_first_time_ = false;
  double first_time_ = 0;
  std::vector<Point> last_points_;
  std::vector<Segment> segments_;
};

}  // namespace orc
