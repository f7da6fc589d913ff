// Spherical-harmonics acceleration and potential of an oblate body:
// Geopotential::GeneralSphericalHarmonicsAcceleration / ...Potential,
// physics/geopotential_body.hpp:740-895, with the damping of
// physics/harmonic_damping_body.hpp:33-82.  Notation follows the reference's
// documentation/Geopotential.pdf.
//
// Two code paths with the same arithmetic, operation for operation:
//  * `GeopotentialUnrolled<max_degree, zonal>`: compile-time (n, m) loops, all
//    recurrence state in registers -- the hot path (Earth at degree/order 10
//    for LEO probes); instantiated for max_degree 2..kMaxUnrolledDegree.
//  * `GeopotentialGeneric`: run-time loops, state in local memory -- any degree
//    up to 50 (the Moon's degree-30 field at close range, Jupiter/Saturn J12).
//
// Summation order (part of the contract): terms are *evaluated*
// by increasing degree then order, and *summed* by unary right folds
// `(accelerations[orders] + ...)`, geopotential_body.hpp:410,535, i.e.
// a0 + (a1 + (... + a_last)).  The per-order recurrences of one degree are
// independent of the order terms, so all updates of a degree run first and
// the terms are then accumulated from the highest order down, which is the
// right fold without storing the terms.
#pragma once

#include <utility>

#include "pb_math.cuh"
#include "pb_model.h"

namespace pb {

constexpr int kMaxUnrolledDegree = 10;

struct GeopotentialBody {
  DeviceBody const* body;
  Damping const* degree_damping;  // [n_damping].
  double const* cos;
  double const* sin;
  double const* normalization;
};

// HarmonicDamping::ComputeDampedRadialQuantities, harmonic_damping_body.hpp:
// 33-61.
PB_HD void ComputeDampedRadialQuantities(Damping const& d, double const r_norm,
                                         double const r2,
                                         Vec3 const& r_normalized,
                                         double const R_over_r,
                                         double const R_prime,
                                         double& sigma_R_over_r,
                                         Vec3& grad_sigma_R) {
  if (r_norm <= d.inner_threshold) {
    sigma_R_over_r = R_over_r;
    grad_sigma_R = R_prime * r_normalized;
  } else {
    double const r3 = r2 * r_norm;
    double const c3r3 = d.c3 * r3;
    double const c2r2 = d.c2 * r2;
    double const c1r = d.c1 * r_norm;
    double const sigma = c3r3 + c2r2 + c1r;
    double const sigma_prime_r = 3 * c3r3 + 2 * c2r2 + c1r;
    sigma_R_over_r = sigma * R_over_r;
    grad_sigma_R = (sigma_prime_r * R_over_r + R_prime * sigma) * r_normalized;
  }
}

// :63-82.
PB_HD void ComputeDampedRadialQuantities(Damping const& d, double const r_norm,
                                         double const r2, double const R_over_r,
                                         double& sigma_R_over_r) {
  if (r_norm <= d.inner_threshold) {
    sigma_R_over_r = R_over_r;
  } else {
    double const r3 = r2 * r_norm;
    double const c3r3 = d.c3 * r3;
    double const c2r2 = d.c2 * r2;
    double const c1r = d.c1 * r_norm;
    double const sigma = c3r3 + c2r2 + c1r;
    sigma_R_over_r = sigma * R_over_r;
  }
}

// Geopotential::LimitingDegree, geopotential_body.hpp:910-922: partition point
// of r_norm < outer_threshold over non-increasing thresholds.
PB_HD int LimitingDegree(Damping const* degree_damping, int const n_damping,
                         double const r_norm) {
  if (!(fabs(r_norm) <= 1.7976931348623157e308)) {
    return 2;
  }
  int lo = 0;
  int hi = n_damping;
  while (lo < hi) {
    int const mid = (lo + hi) / 2;
    if (r_norm < degree_damping[mid].outer_threshold) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo;
}

// Surface-frame axes at time t: FromSurfaceFrame(t) applied to (1,0,0) and
// (0,1,0), geopotential_body.hpp:621-624, geometry/rotation_body.hpp:67-71,
// 155-164,355-363, geometry/quaternion_body.hpp:119-125.
PB_HD void SurfaceFrameAxes(DeviceBody const& body, double const t, Vec3& x_hat,
                            Vec3& y_hat) {
  double const angle = body.reference_angle +
                       (t - body.reference_instant) * body.angular_frequency;
  double const half_angle = 0.5 * angle;
  double s, c;
  CrSinCos(half_angle, s, c);
  // AngleAxis(angle, z) = (c, s * (0, 0, 1)).
  double const r_real = c;
  Vec3 const r_imaginary = s * Vec3{0.0, 0.0, 1.0};
  // q = q_ab * r.
  double const l_real = body.q_ab_real;
  Vec3 const l_imaginary = body.q_ab_imaginary;
  double const q_real = l_real * r_real - Dot(l_imaginary, r_imaginary);
  Vec3 const q_imaginary = l_real * r_imaginary + r_real * l_imaginary +
                           Cross(l_imaginary, r_imaginary);
  Vec3 const ex{1.0, 0.0, 0.0};
  Vec3 const ey{0.0, 1.0, 0.0};
  x_hat = ex + 2 * Cross(q_imaginary, Cross(q_imaginary, ex) + q_real * ex);
  y_hat = ey + 2 * Cross(q_imaginary, Cross(q_imaginary, ey) + q_real * ey);
}

// The quantities of AllDegrees::InitializePrecomputations that do not depend on
// (n, m), geopotential_body.hpp:575-667.
struct GeopotentialSetup {
  double r_norm;
  double r2;
  Vec3 r_normalized;
  double sin_beta;
  double cos_beta;
  double cos_lambda;
  double sin_lambda;
  Vec3 grad_B_vector;
  Vec3 grad_L_vector;
  double R1_over_r;
};

PB_HD void InitializeSetup(DeviceBody const& body, Vec3 const& x_hat,
                           Vec3 const& y_hat, Vec3 const& r,
                           double const r_norm, double const r2,
                           double const one_over_r3, GeopotentialSetup& p) {
  Vec3 const z_hat = body.polar_axis;
  p.r_norm = r_norm;
  p.r2 = r2;
  double const x = Dot(r, x_hat);
  double const y = Dot(r, y_hat);
  double const z = Dot(r, z_hat);
  double const x2_plus_y2 = x * x + y * y;
  double const r_equatorial = sqrt(x2_plus_y2);
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
  p.cos_lambda = cos_lambda;
  p.sin_lambda = sin_lambda;
  p.grad_B_vector = (-p.sin_beta * cos_lambda) * x_hat -
                    (p.sin_beta * sin_lambda) * y_hat + p.cos_beta * z_hat;
  p.grad_L_vector = cos_lambda * y_hat - sin_lambda * x_hat;
  p.R1_over_r = body.reference_radius * one_over_r3;
}

#if defined(__CUDACC__)
// InitializeSetup with the branch-free square root and divisions (the same
// bits for operands in their range, pb_math.cuh): no slow-path branches in the
// prologue of the hot evaluation.  Returns false -- and garbage -- when an
// operand is out of range (a point on the polar axis: r_equatorial == 0); the
// caller then evaluates with InitializeSetup.
__device__ __forceinline__ bool InitializeSetupFast(
    DeviceBody const& body, Vec3 const& x_hat, Vec3 const& y_hat,
    Vec3 const& r, double const r_norm, double const r2,
    double const one_over_r3, GeopotentialSetup& p) {
  Vec3 const z_hat = body.polar_axis;
  p.r_norm = r_norm;
  p.r2 = r2;
  double const x = Dot(r, x_hat);
  double const y = Dot(r, y_hat);
  double const z = Dot(r, z_hat);
  double const x2_plus_y2 = x * x + y * y;
  bool const in_range = InFastRange(x2_plus_y2) && InFastRange(r2);
#if defined(PB_CONTRACTED_BUILD)
  // Contracted arithmetic: one reciprocal square root (hardware seed and a
  // third-order correction, ~1e-16) gives r_equatorial and its inverse;
  // 1 / r = r^2 / r^3.
  double const safe = in_range ? x2_plus_y2 : 1.0;
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(safe));
  double const residual = fma(safe, -(y0 * y0), 1.0);
  double const one_over_r_equatorial =
      fma(fma(residual, 0.375, 0.5), y0 * residual, y0);
  double const r_equatorial = safe * one_over_r_equatorial;
  double const cos_lambda = x * one_over_r_equatorial;
  double const sin_lambda = y * one_over_r_equatorial;
  double const one_over_r_norm = r2 * one_over_r3;
#else
  double const r_equatorial = FastSqrt(in_range ? x2_plus_y2 : 1.0);
  // r_equatorial > 0 here.
  double const one_over_r_equatorial = FastDivide(1.0, r_equatorial);
  double const cos_lambda = x * one_over_r_equatorial;
  double const sin_lambda = y * one_over_r_equatorial;
  double const one_over_r_norm = FastDivide(1.0, in_range ? r_norm : 1.0);
#endif
  p.r_normalized = r * one_over_r_norm;
  p.cos_beta = r_equatorial * one_over_r_norm;
  p.sin_beta = z * one_over_r_norm;
  p.cos_lambda = cos_lambda;
  p.sin_lambda = sin_lambda;
  p.grad_B_vector = (-p.sin_beta * cos_lambda) * x_hat -
                    (p.sin_beta * sin_lambda) * y_hat + p.cos_beta * z_hat;
  p.grad_L_vector = cos_lambda * y_hat - sin_lambda * x_hat;
  p.R1_over_r = body.reference_radius * one_over_r3;
  return in_range;
}
#endif

// ===========================================================================
// Generic path: run-time loops.

// kSize = largest degree + 1: the thread-local arrays are sized for it, so that
// a kernel whose bodies have small models (degree <= 10) keeps a small frame.
template<int kSize>
struct GenericPrecomputations {
  double R_over_r[kSize];
  double cos_m_lambda[kSize];
  double sin_m_lambda[kSize];
  double cos_beta_to_the_m[kSize];
  // Rolling rows n, n-1, n-2 of DmPn_of_sin_beta.
  double P[3][kSize + 1];
};

// DegreeNOrderM::UpdatePrecomputations, geopotential_body.hpp:249-331.
template<int kSize>
PB_HD void GenericUpdateOrder(int const n, int const m, double const sin_beta,
                              double const one_over_n,
                              GenericPrecomputations<kSize>& p) {
  double* const Pn = p.P[n % 3];
  double const* const Pn1 = p.P[(n + 2) % 3];
  double const* const Pn2 = p.P[(n + 1) % 3];
  if (n == 2 && m == 1) {
    Pn[2] = 3;
    return;
  }
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
  // The divisions by n are correctly rounded (DivideBySmallInteger).
  if (m == 0) {
    Pn[0] = DivideBySmallInteger(
        (2 * n - 1) * sin_beta * Pn1[0] - (n - 1) * Pn2[0], n, one_over_n);
  }
  if (m == n) {
  } else if (m == n - 1) {
    Pn[m + 1] =
        DivideBySmallInteger((2 * n - 1) * (m + 1) * Pn1[m], n, one_over_n);
  } else if (m == n - 2) {
    Pn[m + 1] = DivideBySmallInteger(
        (2 * n - 1) * (sin_beta * Pn1[m + 1] + (m + 1) * Pn1[m]), n,
        one_over_n);
  } else {
    Pn[m + 1] = DivideBySmallInteger(
        (2 * n - 1) * (sin_beta * Pn1[m + 1] + (m + 1) * Pn1[m]) -
            (n - 1) * Pn2[m + 1],
        n, one_over_n);
  }
}

// DegreeNOrderM::Acceleration without the update, :130-204.
PB_HD Vec3 OrderTermAcceleration(int const n, int const m, double const Pnm,
                                 double const Pnm1, double const cos_m_lambda,
                                 double const sin_m_lambda,
                                 double const cos_beta_to_the_m,
                                 double const cos_beta_to_the_m_minus_1,
                                 double const Cnm, double const Snm,
                                 double const normalization_factor,
                                 double const sigma_R_over_r,
                                 Vec3 const& grad_sigma_R,
                                 GeopotentialSetup const& s) {
  double const B = cos_beta_to_the_m * Pnm;
  double grad_B_polynomials = 0;
  if (m < n) {
    grad_B_polynomials = s.cos_beta * cos_beta_to_the_m * Pnm1;
  }
  if (m > 0) {
    grad_B_polynomials -= m * s.sin_beta * cos_beta_to_the_m_minus_1 * Pnm;
  }
  double L;
  if (m == 0) {
    L = Cnm;
  } else {
    L = Cnm * cos_m_lambda + Snm * sin_m_lambda;
  }
  Vec3 const BL_grad_R = (B * L) * grad_sigma_R;
  Vec3 const RL_grad_B =
      (sigma_R_over_r * L * grad_B_polynomials) * s.grad_B_vector;
  Vec3 grad_RBL = BL_grad_R + RL_grad_B;
  if (m > 0) {
    Vec3 const RB_grad_L = (sigma_R_over_r * cos_beta_to_the_m_minus_1 * Pnm *
                            m * (Snm * cos_m_lambda - Cnm * sin_m_lambda)) *
                           s.grad_L_vector;
    grad_RBL += RB_grad_L;
  }
  return normalization_factor * grad_RBL;
}

// DegreeNOrderM::Potential without the update, :206-247.
PB_HD double OrderTermPotential(int const m, double const Pnm,
                                double const cos_m_lambda,
                                double const sin_m_lambda,
                                double const cos_beta_to_the_m,
                                double const Cnm, double const Snm,
                                double const normalization_factor,
                                double const sigma_R_over_r,
                                double const r_norm) {
  double const sigma_R = r_norm * sigma_R_over_r;
  double const B = cos_beta_to_the_m * Pnm;
  double L;
  if (m == 0) {
    L = Cnm;
  } else {
    L = Cnm * cos_m_lambda + Snm * sin_m_lambda;
  }
  return -normalization_factor * sigma_R * B * L;
}

template<int kSize>
PB_HD Vec3 GenericTerm(GeopotentialBody const& g, int const n, int const m,
                       double const sigma_R_over_r, Vec3 const& grad_sigma_R,
                       GeopotentialSetup const& s,
                       GenericPrecomputations<kSize> const& p) {
  double const* const Pn = p.P[n % 3];
  int const k = n * (n + 1) / 2 + m;
  return OrderTermAcceleration(
      n, m, Pn[m], m < n ? Pn[m + 1] : 0.0, p.cos_m_lambda[m],
      p.sin_m_lambda[m], p.cos_beta_to_the_m[m],
      m > 0 ? p.cos_beta_to_the_m[m - 1] : 0.0, g.cos[k], g.sin[k],
      g.normalization[k], sigma_R_over_r, grad_sigma_R, s);
}

template<int kSize>
PB_HD double GenericTermPotential(GeopotentialBody const& g, int const n,
                                  int const m, double const sigma_R_over_r,
                                  GeopotentialSetup const& s,
                                  GenericPrecomputations<kSize> const& p) {
  double const* const Pn = p.P[n % 3];
  int const k = n * (n + 1) / 2 + m;
  return OrderTermPotential(m, Pn[m], p.cos_m_lambda[m], p.sin_m_lambda[m],
                            p.cos_beta_to_the_m[m], g.cos[k], g.sin[k],
                            g.normalization[k], sigma_R_over_r, s.r_norm);
}

template<int kSize>
PB_HD void GenericInitialize(GeopotentialSetup const& s,
                             GenericPrecomputations<kSize>& p) {
  p.R_over_r[1] = s.R1_over_r;
  p.cos_m_lambda[1] = s.cos_lambda;
  p.sin_m_lambda[1] = s.sin_lambda;
  p.cos_beta_to_the_m[0] = 1;
  p.cos_beta_to_the_m[1] = s.cos_beta;
  p.P[0][0] = 1;           // DmPn(0, 0).
  p.P[1][0] = s.sin_beta;  // DmPn(1, 0).
  p.P[1][1] = 1;           // DmPn(1, 1).
}

// DegreeNAllOrders::UpdatePrecomputations, :476-499.
PB_HD void GenericUpdateDegree(int const n, double const r2, double* R_over_r) {
  if (n % 2 == 0) {
    int const h = n / 2;
    R_over_r[n] = R_over_r[h] * R_over_r[h] * r2;
  } else {
    int const h1 = n / 2;
    int const h2 = n - h1;
    R_over_r[n] = R_over_r[h1] * R_over_r[h2] * r2;
  }
}

// AllDegrees::Acceleration, :501-536, for max_degree >= 2.
template<int kSize = kGeopotentialSize>
PB_HD_NOINLINE Vec3 GeopotentialGenericAcceleration(
    GeopotentialBody const& g, int const max_degree, bool const is_zonal,
    GeopotentialSetup const& s) {
  GenericPrecomputations<kSize> p;
  GenericInitialize(s, p);
  Vec3 degree_sum[kSize];
  for (int n = 2; n <= max_degree; ++n) {
    int const size = is_zonal ? 1 : n + 1;
    double const one_over_n = 1.0 / n;
    GenericUpdateDegree(n, s.r2, p.R_over_r);
    double const R_over_r = p.R_over_r[n];
    double const R_prime = -(n + 1) * R_over_r;
    double sigma_R_over_r;
    Vec3 grad_sigma_R;
    if (n == 2 && size > 1) {
      ComputeDampedRadialQuantities(g.degree_damping[2], s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      GenericUpdateOrder(2, 0, s.sin_beta, one_over_n, p);
      Vec3 const j2_acceleration =
          GenericTerm(g, 2, 0, sigma_R_over_r, grad_sigma_R, s, p);
      ComputeDampedRadialQuantities(g.body->sectoral, s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      GenericUpdateOrder(2, 1, s.sin_beta, one_over_n, p);
      GenericUpdateOrder(2, 2, s.sin_beta, one_over_n, p);
      Vec3 const c22_s22_acceleration =
          GenericTerm(g, 2, 2, sigma_R_over_r, grad_sigma_R, s, p);
      degree_sum[n] = j2_acceleration + c22_s22_acceleration;
    } else {
      ComputeDampedRadialQuantities(g.degree_damping[n], s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      for (int m = 0; m < size; ++m) {
        GenericUpdateOrder(n, m, s.sin_beta, one_over_n, p);
      }
      Vec3 sum = GenericTerm(g, n, size - 1, sigma_R_over_r, grad_sigma_R, s, p);
      for (int m = size - 2; m >= 0; --m) {
        sum = GenericTerm(g, n, m, sigma_R_over_r, grad_sigma_R, s, p) + sum;
      }
      degree_sum[n] = sum;
    }
  }
  // Degrees 0 and 1 contribute zero vectors to the fold.
  Vec3 sum = degree_sum[max_degree];
  for (int n = max_degree - 1; n >= 2; --n) {
    sum = degree_sum[n] + sum;
  }
  Vec3 const zero{0.0, 0.0, 0.0};
  sum = zero + sum;
  sum = zero + sum;
  return sum;
}

PB_HD_NOINLINE double GeopotentialGenericPotential(
    GeopotentialBody const& g, int const max_degree, bool const is_zonal,
    GeopotentialSetup const& s) {
  GenericPrecomputations<kGeopotentialSize> p;
  GenericInitialize(s, p);
  double degree_sum[kGeopotentialSize];
  for (int n = 2; n <= max_degree; ++n) {
    int const size = is_zonal ? 1 : n + 1;
    double const one_over_n = 1.0 / n;
    GenericUpdateDegree(n, s.r2, p.R_over_r);
    double const R_over_r = p.R_over_r[n];
    double sigma_R_over_r;
    if (n == 2 && size > 1) {
      ComputeDampedRadialQuantities(g.degree_damping[2], s.r_norm, s.r2,
                                    R_over_r, sigma_R_over_r);
      GenericUpdateOrder(2, 0, s.sin_beta, one_over_n, p);
      double const j2_potential =
          GenericTermPotential(g, 2, 0, sigma_R_over_r, s, p);
      ComputeDampedRadialQuantities(g.body->sectoral, s.r_norm, s.r2, R_over_r,
                                    sigma_R_over_r);
      GenericUpdateOrder(2, 1, s.sin_beta, one_over_n, p);
      GenericUpdateOrder(2, 2, s.sin_beta, one_over_n, p);
      double const c22_s22_potential =
          GenericTermPotential(g, 2, 2, sigma_R_over_r, s, p);
      degree_sum[n] = j2_potential + c22_s22_potential;
    } else {
      ComputeDampedRadialQuantities(g.degree_damping[n], s.r_norm, s.r2,
                                    R_over_r, sigma_R_over_r);
      for (int m = 0; m < size; ++m) {
        GenericUpdateOrder(n, m, s.sin_beta, one_over_n, p);
      }
      double sum = GenericTermPotential(g, n, size - 1, sigma_R_over_r, s, p);
      for (int m = size - 2; m >= 0; --m) {
        sum = GenericTermPotential(g, n, m, sigma_R_over_r, s, p) + sum;
      }
      degree_sum[n] = sum;
    }
  }
  double sum = degree_sum[max_degree];
  for (int n = max_degree - 1; n >= 2; --n) {
    sum = degree_sum[n] + sum;
  }
  sum = 0.0 + sum;
  sum = 0.0 + sum;
  return sum;
}

// ===========================================================================
// Unrolled path: compile-time (n, m).

template<int... Is, typename F>
PB_HD void StaticFor(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template<int kBegin, int kEnd, typename F>
PB_HD void StaticForRange(F&& f) {
  if constexpr (kBegin < kEnd) {
    StaticFor(std::make_integer_sequence<int, kEnd - kBegin>{}, [&](auto i) {
      f(std::integral_constant<int, kBegin + decltype(i)::value>{});
    });
  }
}

template<int kMaxDegree>
struct UnrolledPrecomputations {
  double R_over_r[kMaxDegree + 1];
  // Full triangle; every index is a compile-time constant, so after unrolling
  // these are plain values with short live ranges (rows n-1, n-2).
  double P[(kMaxDegree + 1) * (kMaxDegree + 2) / 2 + 1];
};

// The long-lived state of one evaluation -- cos mλ, sin mλ, cos^m β (created at
// degree m, used by every later degree) and the per-degree sums (kept for the
// final right fold) -- lives behind a Scratch policy: thread-local arrays here,
// shared memory in the flow kernel (108 registers' worth of state that would
// otherwise be spilled).
template<int kMaxDegree>
struct LocalScratch {
  // Whether the contracted build may take the factored evaluation.
  static constexpr bool kFactored = false;
  double cos_m_lambda[kMaxDegree + 1];
  double sin_m_lambda[kMaxDegree + 1];
  double cos_beta_to_the_m[kMaxDegree + 1];
  Vec3 degree_sum[kMaxDegree + 1];
  template<int m> PB_HD double CosMLambda() const { return cos_m_lambda[m]; }
  template<int m> PB_HD double SinMLambda() const { return sin_m_lambda[m]; }
  template<int m> PB_HD double CosBetaToTheM() const {
    return cos_beta_to_the_m[m];
  }
  template<int m> PB_HD void SetCosMLambda(double const x) {
    cos_m_lambda[m] = x;
  }
  template<int m> PB_HD void SetSinMLambda(double const x) {
    sin_m_lambda[m] = x;
  }
  template<int m> PB_HD void SetCosBetaToTheM(double const x) {
    cos_beta_to_the_m[m] = x;
  }
  template<int n> PB_HD Vec3 DegreeSum() const { return degree_sum[n]; }
  template<int n> PB_HD void SetDegreeSum(Vec3 const& x) { degree_sum[n] = x; }
  // The division by the degree in the Legendre recurrences, and whether the
  // evaluation has to be abandoned (never, with the checked division).
  PB_HD double Divide(double const x, int const n, double const reciprocal) {
    return DivideBySmallInteger(x, n, reciprocal);
  }
  PB_HD constexpr bool Failed() const { return false; }
};

#if defined(__CUDACC__)
// The same scratch with the branch-free division: DivideBySmallInteger's fast
// path unconditionally, its operand test folded into a running maximum (two
// integer instructions).  Without the 64 branches of an evaluation the
// recurrences of a degree are one basic block that the scheduler can
// interleave.  When Failed() the results are garbage and the caller evaluates
// again with the checked division; the unrolled evaluation tests it once per
// degree, which also keeps the degrees apart in the instruction stream (a
// single block for all of them makes ptxas hoist far too much and spill).
template<int kMaxDegree>
struct FlaggedScratch : LocalScratch<kMaxDegree> {
  static constexpr bool kFactored = true;
  unsigned flag = 0;
  __device__ __forceinline__ double Divide(double const x, int const n,
                                           double const reciprocal) {
    // exponent < 0x036 or exponent == 0x7ff <=> offset >= 0x7c900000.
#if defined(PB_CONTRACTED_BUILD)
    // Contracted arithmetic: the product with RN(1 / n), within one ulp of the
    // quotient for any operand.
    return x * reciprocal;
#else
    flag = __viaddmax_u32(
        static_cast<unsigned>(__double2hiint(x)) & 0x7ff00000u,
        0u - 0x03600000u, flag);
    double const q0 = x * reciprocal;
    double const r = fma(-static_cast<double>(n), q0, x);
    return fma(r, reciprocal, q0);
#endif
  }
  __device__ __forceinline__ bool Failed() const {
    return flag >= 0x7c900000u;
  }
};
#endif

PB_HD constexpr int Tri(int const n, int const m) {
  return n * (n + 1) / 2 + m;
}

template<int n, int m, int kMaxDegree, typename Scratch>
PB_HD void UnrolledUpdateOrder(double const sin_beta,
                               UnrolledPrecomputations<kMaxDegree>& p,
                               Scratch& w) {
  if constexpr (n == 2 && m == 1) {
    p.P[Tri(2, 2)] = 3;
  } else {
    if constexpr (m == n) {
      if constexpr (m % 2 == 0) {
        constexpr int h = m / 2;
        double const cos_h = w.template CosMLambda<h>();
        double const sin_h = w.template SinMLambda<h>();
        double const cos_beta_to_the_h = w.template CosBetaToTheM<h>();
        w.template SetSinMLambda<m>(2 * sin_h * cos_h);
        w.template SetCosMLambda<m>((cos_h + sin_h) * (cos_h - sin_h));
        w.template SetCosBetaToTheM<m>(cos_beta_to_the_h * cos_beta_to_the_h);
      } else {
        constexpr int h1 = m / 2;
        constexpr int h2 = m - h1;
        double const cos_h1 = w.template CosMLambda<h1>();
        double const sin_h1 = w.template SinMLambda<h1>();
        double const cos_beta_to_the_h1 = w.template CosBetaToTheM<h1>();
        double const cos_h2 = w.template CosMLambda<h2>();
        double const sin_h2 = w.template SinMLambda<h2>();
        double const cos_beta_to_the_h2 = w.template CosBetaToTheM<h2>();
        w.template SetSinMLambda<m>(sin_h1 * cos_h2 + cos_h1 * sin_h2);
        w.template SetCosMLambda<m>(cos_h1 * cos_h2 - sin_h1 * sin_h2);
        w.template SetCosBetaToTheM<m>(cos_beta_to_the_h1 * cos_beta_to_the_h2);
      }
    }
    constexpr double one_over_n = 1.0 / n;
    if constexpr (m == 0) {
      p.P[Tri(n, 0)] = w.Divide(
          (2 * n - 1) * sin_beta * p.P[Tri(n - 1, 0)] -
              (n - 1) * p.P[Tri(n - 2, 0)],
          n, one_over_n);
    }
    if constexpr (m == n) {
    } else if constexpr (m == n - 1) {
      p.P[Tri(n, m + 1)] = w.Divide(
          (2 * n - 1) * (m + 1) * p.P[Tri(n - 1, m)], n, one_over_n);
    } else if constexpr (m == n - 2) {
      p.P[Tri(n, m + 1)] = w.Divide(
          (2 * n - 1) * (sin_beta * p.P[Tri(n - 1, m + 1)] +
                         (m + 1) * p.P[Tri(n - 1, m)]),
          n, one_over_n);
    } else {
      p.P[Tri(n, m + 1)] = w.Divide(
          (2 * n - 1) * (sin_beta * p.P[Tri(n - 1, m + 1)] +
                         (m + 1) * p.P[Tri(n - 1, m)]) -
              (n - 1) * p.P[Tri(n - 2, m + 1)],
          n, one_over_n);
    }
  }
}

// Where the unrolled path reads a body's tables from.  PointerTables: global
// memory through pointers (any body).  A kernel may instead provide a type
// with the same interface backed by __constant__ memory, so that with
// compile-time (n, m) every coefficient becomes a constant-bank operand of the
// FP64 instruction that uses it (no load instruction, no scoreboard wait).
struct PointerTables {
  GeopotentialBody const& g;
  PB_HD double Cos(int const k) const { return g.cos[k]; }
  PB_HD double Sin(int const k) const { return g.sin[k]; }
  PB_HD double Normalization(int const k) const { return g.normalization[k]; }
  PB_HD Damping const& DegreeDamping(int const n) const {
    return g.degree_damping[n];
  }
  PB_HD Damping const& SectoralDamping() const { return g.body->sectoral; }
};

template<int n, int m, int kMaxDegree, typename Tables, typename Scratch>
PB_HD Vec3 UnrolledTerm(Tables const& g, double const sigma_R_over_r,
                        Vec3 const& grad_sigma_R, GeopotentialSetup const& s,
                        UnrolledPrecomputations<kMaxDegree> const& p,
                        Scratch const& w) {
  constexpr int k = Tri(n, m);
  double Pnm1 = 0.0;
  if constexpr (m < n) {
    Pnm1 = p.P[Tri(n, m + 1)];
  }
  double cos_beta_to_the_m_minus_1 = 0.0;
  double cos_m_lambda = 0.0;
  double sin_m_lambda = 0.0;
  if constexpr (m > 0) {
    cos_beta_to_the_m_minus_1 = w.template CosBetaToTheM<m - 1>();
    cos_m_lambda = w.template CosMLambda<m>();
    sin_m_lambda = w.template SinMLambda<m>();
  }
  return OrderTermAcceleration(n, m, p.P[k], Pnm1, cos_m_lambda, sin_m_lambda,
                               w.template CosBetaToTheM<m>(),
                               cos_beta_to_the_m_minus_1, g.Cos(k), g.Sin(k),
                               g.Normalization(k), sigma_R_over_r,
                               grad_sigma_R, s);
}

#if defined(PB_CONTRACTED_BUILD)
// The contracted arithmetic is not bound to the reference's expression order:
// the acceleration of degree n is
//   grad(sigma R) SUM_m N B L + sigma R / r (gradB SUM_m N L B' + gradL SUM_m N B'' L')
// with three scalar sums over the orders and the vectors applied once per
// degree -- about 24 flops per term instead of 41 (OrderTermAcceleration forms
// and scales three vectors per term).  Same recurrences, same damping.
template<int kMaxDegree, bool kZonal, typename Tables, typename Scratch>
__device__ __forceinline__ Vec3 GeopotentialUnrolledAccelerationFactored(
    Tables const& g, GeopotentialSetup const& s, Scratch& w) {
  UnrolledPrecomputations<kMaxDegree> p;
  p.R_over_r[1] = s.R1_over_r;
  w.template SetCosMLambda<1>(s.cos_lambda);
  w.template SetSinMLambda<1>(s.sin_lambda);
  w.template SetCosBetaToTheM<0>(1);
  w.template SetCosBetaToTheM<1>(s.cos_beta);
  p.P[Tri(0, 0)] = 1;
  p.P[Tri(1, 0)] = s.sin_beta;
  p.P[Tri(1, 1)] = 1;
  Vec3 total{0.0, 0.0, 0.0};
  StaticForRange<2, kMaxDegree + 1>([&](auto n_constant) {
    constexpr int n = decltype(n_constant)::value;
    constexpr int size = kZonal ? 1 : n + 1;
    if constexpr (n % 2 == 0) {
      constexpr int h = n / 2;
      p.R_over_r[n] = p.R_over_r[h] * p.R_over_r[h] * s.r2;
    } else {
      constexpr int h1 = n / 2;
      constexpr int h2 = n - h1;
      p.R_over_r[n] = p.R_over_r[h1] * p.R_over_r[h2] * s.r2;
    }
    double const R_over_r = p.R_over_r[n];
    double const R_prime = -(n + 1) * R_over_r;
    double sigma_R_over_r;
    Vec3 grad_sigma_R;
    // The three sums of one order.
    auto const order = [&](auto m_constant, double& radial, double& latitude,
                           double& longitude) {
      constexpr int m = decltype(m_constant)::value;
      constexpr int k = Tri(n, m);
      double const Pnm = p.P[k];
      double const cos_beta_m = w.template CosBetaToTheM<m>();
      double const normalization = g.Normalization(k);
      double L;
      if constexpr (m == 0) {
        L = g.Cos(k);
      } else {
        L = g.Cos(k) * w.template CosMLambda<m>() +
            g.Sin(k) * w.template SinMLambda<m>();
      }
      double const normalized_L = normalization * L;
      radial += normalized_L * (cos_beta_m * Pnm);
      double grad_B = 0;
      if constexpr (m < n) {
        grad_B = s.cos_beta * cos_beta_m * p.P[Tri(n, m + 1)];
      }
      if constexpr (m > 0) {
        double const cos_beta_m1_Pnm =
            w.template CosBetaToTheM<m - 1>() * Pnm;
        grad_B -= m * s.sin_beta * cos_beta_m1_Pnm;
        longitude += normalization *
                     (cos_beta_m1_Pnm * m *
                      (g.Sin(k) * w.template CosMLambda<m>() -
                       g.Cos(k) * w.template SinMLambda<m>()));
      }
      latitude += normalized_L * grad_B;
    };
    auto const add_degree = [&](double const radial, double const latitude,
                                double const longitude) {
      total += radial * grad_sigma_R +
               sigma_R_over_r * (latitude * s.grad_B_vector +
                                 longitude * s.grad_L_vector);
    };
    if constexpr (n == 2 && size > 1) {
      ComputeDampedRadialQuantities(g.DegreeDamping(2), s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      UnrolledUpdateOrder<2, 0>(s.sin_beta, p, w);
      double radial = 0, latitude = 0, longitude = 0;
      order(std::integral_constant<int, 0>{}, radial, latitude, longitude);
      add_degree(radial, latitude, longitude);
      ComputeDampedRadialQuantities(g.SectoralDamping(), s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      UnrolledUpdateOrder<2, 1>(s.sin_beta, p, w);
      UnrolledUpdateOrder<2, 2>(s.sin_beta, p, w);
      radial = latitude = longitude = 0;
      order(std::integral_constant<int, 2>{}, radial, latitude, longitude);
      add_degree(radial, latitude, longitude);
    } else {
      ComputeDampedRadialQuantities(g.DegreeDamping(n), s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      StaticForRange<0, size>([&](auto m_constant) {
        UnrolledUpdateOrder<n, decltype(m_constant)::value>(s.sin_beta, p, w);
      });
      double radial = 0, latitude = 0, longitude = 0;
      StaticForRange<0, size>([&](auto m_constant) {
        order(m_constant, radial, latitude, longitude);
      });
      add_degree(radial, latitude, longitude);
    }
  });
  return total;
}
#endif

// kCapped: the degrees above the run-time `max_degree` (2 <= max_degree <=
// kMaxDegree) are skipped and left out of the fold, which gives the operations
// of GeopotentialUnrolledAcceleration<max_degree>: ONE straight-line body for
// the lanes of a warp that evaluate different degrees (the adaptive flow),
// where the dispatch on the degree would run one body per distinct degree.
template<int kMaxDegree, bool kZonal, bool kCapped = false, typename Tables,
         typename Scratch>
PB_HD Vec3 GeopotentialUnrolledAcceleration(Tables const& g,
                                            GeopotentialSetup const& s,
                                            Scratch& w,
                                            int const max_degree = kMaxDegree) {
#if defined(PB_CONTRACTED_BUILD) && defined(__CUDA_ARCH__)
  if constexpr (Scratch::kFactored && !kCapped) {
    return GeopotentialUnrolledAccelerationFactored<kMaxDegree, kZonal>(g, s,
                                                                        w);
  }
#endif
  UnrolledPrecomputations<kMaxDegree> p;
  p.R_over_r[1] = s.R1_over_r;
  w.template SetCosMLambda<1>(s.cos_lambda);
  w.template SetSinMLambda<1>(s.sin_lambda);
  w.template SetCosBetaToTheM<0>(1);
  w.template SetCosBetaToTheM<1>(s.cos_beta);
  p.P[Tri(0, 0)] = 1;
  p.P[Tri(1, 0)] = s.sin_beta;
  p.P[Tri(1, 1)] = 1;
  StaticForRange<2, kMaxDegree + 1>([&](auto n_constant) {
    constexpr int n = decltype(n_constant)::value;
    constexpr int size = kZonal ? 1 : n + 1;
    if (w.Failed()) {
      return;
    }
    if constexpr (kCapped) {
      if (n > max_degree) {
        w.template SetDegreeSum<n>(Vec3{0.0, 0.0, 0.0});
        return;
      }
    }
    if constexpr (n % 2 == 0) {
      constexpr int h = n / 2;
      p.R_over_r[n] = p.R_over_r[h] * p.R_over_r[h] * s.r2;
    } else {
      constexpr int h1 = n / 2;
      constexpr int h2 = n - h1;
      p.R_over_r[n] = p.R_over_r[h1] * p.R_over_r[h2] * s.r2;
    }
    double const R_over_r = p.R_over_r[n];
    double const R_prime = -(n + 1) * R_over_r;
    double sigma_R_over_r;
    Vec3 grad_sigma_R;
    if constexpr (n == 2 && size > 1) {
      ComputeDampedRadialQuantities(g.DegreeDamping(2), s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      UnrolledUpdateOrder<2, 0>(s.sin_beta, p, w);
      Vec3 const j2_acceleration =
          UnrolledTerm<2, 0>(g, sigma_R_over_r, grad_sigma_R, s, p, w);
      ComputeDampedRadialQuantities(g.SectoralDamping(), s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      UnrolledUpdateOrder<2, 1>(s.sin_beta, p, w);
      UnrolledUpdateOrder<2, 2>(s.sin_beta, p, w);
      Vec3 const c22_s22_acceleration =
          UnrolledTerm<2, 2>(g, sigma_R_over_r, grad_sigma_R, s, p, w);
      w.template SetDegreeSum<n>(j2_acceleration + c22_s22_acceleration);
    } else {
      ComputeDampedRadialQuantities(g.DegreeDamping(n), s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      StaticForRange<0, size>([&](auto m_constant) {
        UnrolledUpdateOrder<n, decltype(m_constant)::value>(s.sin_beta, p, w);
      });
#if defined(PB_H_FENCE_TERMS)
      if (w.Failed()) {
        return;
      }
#endif
      Vec3 sum =
          UnrolledTerm<n, size - 1>(g, sigma_R_over_r, grad_sigma_R, s, p, w);
      // (Evaluating two to four orders in lock step, operation by operation,
      // to interleave their dependency chains was measured 30 % slower at 128
      // registers and 5 % slower at 192: the spills outweigh the latency.)
      StaticForRange<0, size - 1>([&](auto i_constant) {
        constexpr int m = size - 2 - decltype(i_constant)::value;
        sum = UnrolledTerm<n, m>(g, sigma_R_over_r, grad_sigma_R, s, p, w) + sum;
      });
      w.template SetDegreeSum<n>(sum);
    }
  });
  Vec3 sum = w.template DegreeSum<kMaxDegree>();
  StaticForRange<0, kMaxDegree - 2>([&](auto i_constant) {
    constexpr int n = kMaxDegree - 1 - decltype(i_constant)::value;
    if constexpr (kCapped) {
      // The fold starts at max_degree.
      Vec3 const own = w.template DegreeSum<n>();
      Vec3 const folded = own + sum;
      bool const start = n == max_degree;
      bool const below = n < max_degree;
      sum = {start ? own.x : (below ? folded.x : sum.x),
             start ? own.y : (below ? folded.y : sum.y),
             start ? own.z : (below ? folded.z : sum.z)};
    } else {
      sum = w.template DegreeSum<n>() + sum;
    }
  });
  Vec3 const zero{0.0, 0.0, 0.0};
  sum = zero + sum;
  sum = zero + sum;
  return sum;
}

// With thread-local scratch.
template<int kMaxDegree, bool kZonal, typename Tables>
PB_HD Vec3 GeopotentialUnrolledAcceleration(Tables const& g,
                                            GeopotentialSetup const& s) {
  LocalScratch<kMaxDegree> scratch;
  return GeopotentialUnrolledAcceleration<kMaxDegree, kZonal>(g, s, scratch);
}

// Dispatch on the run-time degree; only degrees <= kLimit are instantiated
// (callers guarantee max_degree <= kLimit).
template<bool kZonal, typename Tables, int kLimit = kMaxUnrolledDegree>
PB_HD Vec3 GeopotentialUnrolledDispatch(Tables const& g,
                                        int const max_degree,
                                        GeopotentialSetup const& s) {
#define PB_UNROLLED_CASE(d)                                        \
  if constexpr (kLimit > (d)) {                                    \
    if (max_degree == (d)) {                                       \
      return GeopotentialUnrolledAcceleration<(d), kZonal>(g, s);  \
    }                                                              \
  }
  PB_UNROLLED_CASE(2)
  PB_UNROLLED_CASE(3)
  PB_UNROLLED_CASE(4)
  PB_UNROLLED_CASE(5)
  PB_UNROLLED_CASE(6)
  PB_UNROLLED_CASE(7)
  PB_UNROLLED_CASE(8)
  PB_UNROLLED_CASE(9)
#undef PB_UNROLLED_CASE
  return GeopotentialUnrolledAcceleration<kLimit, kZonal>(g, s);
}

// Same, with the caller's scratch storage (sized for kLimit).
template<bool kZonal, typename Tables, typename Scratch,
         int kLimit = kMaxUnrolledDegree>
PB_HD Vec3 GeopotentialUnrolledDispatchWithScratch(Tables const& g,
                                                   int const max_degree,
                                                   GeopotentialSetup const& s,
                                                   Scratch& w) {
  // The full degree first: it is the common case of a hot body, and every
  // test of the chain below is a jump over tens of kB of unrolled code (an
  // instruction-cache miss each).
  if (max_degree >= kLimit) {
    return GeopotentialUnrolledAcceleration<kLimit, kZonal>(g, s, w);
  }
#define PB_UNROLLED_CASE(d)                                           \
  if constexpr (kLimit > (d)) {                                       \
    if (max_degree == (d)) {                                          \
      return GeopotentialUnrolledAcceleration<(d), kZonal>(g, s, w);  \
    }                                                                 \
  }
  PB_UNROLLED_CASE(2)
  PB_UNROLLED_CASE(3)
  PB_UNROLLED_CASE(4)
  PB_UNROLLED_CASE(5)
  PB_UNROLLED_CASE(6)
  PB_UNROLLED_CASE(7)
  PB_UNROLLED_CASE(8)
  PB_UNROLLED_CASE(9)
#undef PB_UNROLLED_CASE
  return GeopotentialUnrolledAcceleration<kLimit, kZonal>(g, s, w);
}

// ===========================================================================
// Entry points.

// Where the surface-frame axes of a non-zonal evaluation come from: a stage
// table entry (6 doubles, computed once per stage time for all probes) or, for
// per-probe times (adaptive flow), the time itself.
struct FrameSource {
  double const* axes;
  double t;
};

// Geopotential::GeneralSphericalHarmonicsAcceleration, :740-813.  `frame` holds
// the surface-frame x and y axes at the evaluation time (6 doubles) when the
// body has a frame slot; zonal evaluation uses the body's fixed axes.
// kUnrolled selects the register-resident path for max_degree <= 10.
template<bool kUnrolled, int kUnrolledLimit = kMaxUnrolledDegree>
PB_HD Vec3 GeneralSphericalHarmonicsAcceleration(GeopotentialBody const& g,
                                                 FrameSource const frame,
                                                 Vec3 const& r,
                                                 double const r_norm,
                                                 double const r2,
                                                 double const one_over_r3) {
  if (r_norm != r_norm) {
    double const nan = r_norm;
    Vec3 const zero{0.0, 0.0, 0.0};
    return nan * zero;
  }
  DeviceBody const& body = *g.body;
  int const max_degree =
      LimitingDegree(g.degree_damping, body.n_damping, r_norm) - 1;
  if (max_degree == 1) {
    return {0.0, 0.0, 0.0};
  }
  bool const is_zonal =
      body.is_zonal || r_norm > body.sectoral.outer_threshold;
  Vec3 x_hat, y_hat;
  if (is_zonal) {
    x_hat = body.equatorial;
    y_hat = body.biequatorial;
  } else if (frame.axes != nullptr) {
    x_hat = {frame.axes[0], frame.axes[1], frame.axes[2]};
    y_hat = {frame.axes[3], frame.axes[4], frame.axes[5]};
  } else {
    SurfaceFrameAxes(body, frame.t, x_hat, y_hat);
  }
  GeopotentialSetup s;
  InitializeSetup(body, x_hat, y_hat, r, r_norm, r2, one_over_r3, s);
  if constexpr (kUnrolled) {
    if (max_degree <= kUnrolledLimit) {
      PointerTables const tables{g};
      if (is_zonal) {
        return GeopotentialUnrolledDispatch<true, PointerTables,
                                            kUnrolledLimit>(tables, max_degree,
                                                            s);
      } else {
        return GeopotentialUnrolledDispatch<false, PointerTables,
                                            kUnrolledLimit>(tables, max_degree,
                                                            s);
      }
    }
  }
  return GeopotentialGenericAcceleration<>(g, max_degree, is_zonal, s);
}

// AllDegrees::Acceleration with run-time loops, tuned for instruction count:
// the same operations as GeopotentialGenericAcceleration (and therefore the
// same bits), with the special orders (m = 0, n - 2, n - 1, n) peeled off the
// loops, rotating row pointers instead of n % 3, and running table indices.
// RN(1 / n) for DivideBySmallInteger.
// `Tables`: PointerTables or a kernel's constant-bank equivalent.
template<int kSize, typename Tables>
PB_HD_NOINLINE Vec3 GeopotentialRolledAcceleration(
    Tables const& g, int const max_degree, bool const is_zonal,
    GeopotentialSetup const& s) {
  double R_over_r_[kSize];
  double cos_m_lambda[kSize];
  double sin_m_lambda[kSize];
  double cos_beta_to_the_m[kSize];
  double rows[3][kSize + 1];
  Vec3 degree_sum[kSize];
  R_over_r_[1] = s.R1_over_r;
  cos_m_lambda[1] = s.cos_lambda;
  sin_m_lambda[1] = s.sin_lambda;
  cos_beta_to_the_m[0] = 1;
  cos_beta_to_the_m[1] = s.cos_beta;
  double* Pn2 = rows[0];
  double* Pn1 = rows[1];
  double* Pn = rows[2];
  Pn2[0] = 1;           // DmPn(0, 0).
  Pn1[0] = s.sin_beta;  // DmPn(1, 0).
  Pn1[1] = 1;           // DmPn(1, 1).
  double const sin_beta = s.sin_beta;
  // Index of (n, 0) in the triangular tables.
  int row = 3;
  for (int n = 2; n <= max_degree; ++n) {
    double const one_over_n = 1.0 / n;
    {
      int const h1 = n / 2;
      int const h2 = n - h1;
      R_over_r_[n] = R_over_r_[h1] * R_over_r_[h2] * s.r2;
    }
    double const R_over_r = R_over_r_[n];
    double const R_prime = -(n + 1) * R_over_r;
    double sigma_R_over_r;
    Vec3 grad_sigma_R;
    // One term; `m`'s relation to 0 and n is known at each call site.
    auto const term = [&](int const m, double const Pnm1,
                          double const cos_beta_to_the_m_minus_1) {
      return OrderTermAcceleration(
          n, m, Pn[m], Pnm1, cos_m_lambda[m], sin_m_lambda[m],
          cos_beta_to_the_m[m], cos_beta_to_the_m_minus_1, g.Cos(row + m),
          g.Sin(row + m), g.Normalization(row + m), sigma_R_over_r,
          grad_sigma_R, s);
    };
    // m = 0 of DegreeNOrderM::UpdatePrecomputations: DmPn(n, 0), DmPn(n, 1).
    Pn[0] = DivideBySmallInteger(
        (2 * n - 1) * sin_beta * Pn1[0] - (n - 1) * Pn2[0], n, one_over_n);
    if (n == 2) {
      Pn[1] = DivideBySmallInteger(
          (2 * n - 1) * (sin_beta * Pn1[1] + 1 * Pn1[0]), n, one_over_n);
    } else {
      Pn[1] = DivideBySmallInteger(
          (2 * n - 1) * (sin_beta * Pn1[1] + 1 * Pn1[0]) - (n - 1) * Pn2[1], n,
          one_over_n);
    }
    if (is_zonal) {
      ComputeDampedRadialQuantities(g.DegreeDamping(n), s.r_norm, s.r2,
                                    s.r_normalized, R_over_r, R_prime,
                                    sigma_R_over_r, grad_sigma_R);
      degree_sum[n] = OrderTermAcceleration(
          n, 0, Pn[0], Pn[1], 0.0, 0.0, 1.0, 0.0, g.Cos(row), g.Sin(row),
          g.Normalization(row), sigma_R_over_r, grad_sigma_R, s);
    } else {
      // Orders 1 .. n - 3 (general recurrence), n - 2, n - 1.
      for (int m = 1; m <= n - 3; ++m) {
        Pn[m + 1] = DivideBySmallInteger(
            (2 * n - 1) * (sin_beta * Pn1[m + 1] + (m + 1) * Pn1[m]) -
                (n - 1) * Pn2[m + 1],
            n, one_over_n);
      }
      if (n >= 3) {
        int const m = n - 2;
        Pn[m + 1] = DivideBySmallInteger(
            (2 * n - 1) * (sin_beta * Pn1[m + 1] + (m + 1) * Pn1[m]), n,
            one_over_n);
      }
      if (n == 2) {
        Pn[2] = 3;
      } else {
        int const m = n - 1;
        Pn[m + 1] = DivideBySmallInteger((2 * n - 1) * (m + 1) * Pn1[m], n,
                                         one_over_n);
      }
      // m = n: cos nλ, sin nλ, cos^n β from the halves.
      {
        int const h1 = n / 2;
        int const h2 = n - h1;
        double const cos_h1 = cos_m_lambda[h1];
        double const sin_h1 = sin_m_lambda[h1];
        double const cos_beta_to_the_h1 = cos_beta_to_the_m[h1];
        if (h1 == h2) {
          sin_m_lambda[n] = 2 * sin_h1 * cos_h1;
          cos_m_lambda[n] = (cos_h1 + sin_h1) * (cos_h1 - sin_h1);
          cos_beta_to_the_m[n] = cos_beta_to_the_h1 * cos_beta_to_the_h1;
        } else {
          double const cos_h2 = cos_m_lambda[h2];
          double const sin_h2 = sin_m_lambda[h2];
          double const cos_beta_to_the_h2 = cos_beta_to_the_m[h2];
          sin_m_lambda[n] = sin_h1 * cos_h2 + cos_h1 * sin_h2;
          cos_m_lambda[n] = cos_h1 * cos_h2 - sin_h1 * sin_h2;
          cos_beta_to_the_m[n] = cos_beta_to_the_h1 * cos_beta_to_the_h2;
        }
      }
      if (n == 2) {
        ComputeDampedRadialQuantities(g.DegreeDamping(2), s.r_norm, s.r2,
                                      s.r_normalized, R_over_r, R_prime,
                                      sigma_R_over_r, grad_sigma_R);
        Vec3 const j2_acceleration = OrderTermAcceleration(
            2, 0, Pn[0], Pn[1], 0.0, 0.0, 1.0, 0.0, g.Cos(row), g.Sin(row),
            g.Normalization(row), sigma_R_over_r, grad_sigma_R, s);
        ComputeDampedRadialQuantities(g.SectoralDamping(), s.r_norm, s.r2,
                                      s.r_normalized, R_over_r, R_prime,
                                      sigma_R_over_r, grad_sigma_R);
        Vec3 const c22_s22_acceleration = term(2, 0.0, cos_beta_to_the_m[1]);
        degree_sum[n] = j2_acceleration + c22_s22_acceleration;
      } else {
        ComputeDampedRadialQuantities(g.DegreeDamping(n), s.r_norm, s.r2,
                                      s.r_normalized, R_over_r, R_prime,
                                      sigma_R_over_r, grad_sigma_R);
        Vec3 sum = term(n, 0.0, cos_beta_to_the_m[n - 1]);
        for (int m = n - 1; m >= 1; --m) {
          sum = term(m, Pn[m + 1], cos_beta_to_the_m[m - 1]) + sum;
        }
        sum = OrderTermAcceleration(n, 0, Pn[0], Pn[1], 0.0, 0.0, 1.0, 0.0,
                                    g.Cos(row), g.Sin(row), g.Normalization(row),
                                    sigma_R_over_r, grad_sigma_R, s) +
              sum;
        degree_sum[n] = sum;
      }
    }
    row += n + 1;
    double* const recycled = Pn2;
    Pn2 = Pn1;
    Pn1 = Pn;
    Pn = recycled;
  }
  // Degrees 0 and 1 contribute zero vectors to the fold.
  Vec3 sum = degree_sum[max_degree];
  for (int n = max_degree - 1; n >= 2; --n) {
    sum = degree_sum[n] + sum;
  }
  Vec3 const zero{0.0, 0.0, 0.0};
  sum = zero + sum;
  sum = zero + sum;
  return sum;
}

// GeneralSphericalHarmonicsAcceleration for kernels whose warps evaluate many
// different (body, degree) combinations at once (the adaptive flow): run-time
// loops only, a few kB of code that stay in the instruction cache, where the
// unrolled path has one large straight-line body per degree.  The lanes of a
// warp walk the (n, m) loops together and drop out at their own max_degree.
// Surface-frame axes from the time.
PB_HD Vec3 GeneralSphericalHarmonicsAccelerationRolled(
    GeopotentialBody const& g, double const t, Vec3 const& r,
    double const r_norm, double const r2, double const one_over_r3) {
  if (r_norm != r_norm) {
    double const nan = r_norm;
    Vec3 const zero{0.0, 0.0, 0.0};
    return nan * zero;
  }
  DeviceBody const& body = *g.body;
  int const max_degree =
      LimitingDegree(g.degree_damping, body.n_damping, r_norm) - 1;
  if (max_degree == 1) {
    return {0.0, 0.0, 0.0};
  }
  bool const is_zonal =
      body.is_zonal || r_norm > body.sectoral.outer_threshold;
  Vec3 x_hat, y_hat;
  if (is_zonal) {
    x_hat = body.equatorial;
    y_hat = body.biequatorial;
  } else {
    SurfaceFrameAxes(body, t, x_hat, y_hat);
  }
  GeopotentialSetup s;
  InitializeSetup(body, x_hat, y_hat, r, r_norm, r2, one_over_r3, s);
  PointerTables const tables{g};
  if (max_degree <= kMaxUnrolledDegree) {
    // (A variant with the degrees rolled and the orders unrolled, per-order
    // state in registers, was measured 5 % slower here and 50 % slower than
    // the fully unrolled body in the fixed-step kernel: its 19 kB loop body
    // misses the instruction cache next to the rest of the kernel.)
    return GeopotentialRolledAcceleration<kMaxUnrolledDegree + 1>(
        tables, max_degree, is_zonal, s);
  }
  return GeopotentialRolledAcceleration<kGeopotentialSize>(tables, max_degree,
                                                           is_zonal, s);
}

// Geopotential::GeneralSphericalHarmonicsPotential, :822-895.
PB_HD double GeneralSphericalHarmonicsPotential(GeopotentialBody const& g,
                                                FrameSource const frame,
                                                Vec3 const& r,
                                                double const r_norm,
                                                double const r2,
                                                double const one_over_r3) {
  if (r_norm != r_norm) {
    return r_norm;
  }
  DeviceBody const& body = *g.body;
  int const max_degree =
      LimitingDegree(g.degree_damping, body.n_damping, r_norm) - 1;
  if (max_degree == 1) {
    return 0.0;
  }
  bool const is_zonal =
      body.is_zonal || r_norm > body.sectoral.outer_threshold;
  Vec3 x_hat, y_hat;
  if (is_zonal) {
    x_hat = body.equatorial;
    y_hat = body.biequatorial;
  } else if (frame.axes != nullptr) {
    x_hat = {frame.axes[0], frame.axes[1], frame.axes[2]};
    y_hat = {frame.axes[3], frame.axes[4], frame.axes[5]};
  } else {
    SurfaceFrameAxes(body, frame.t, x_hat, y_hat);
  }
  GeopotentialSetup s;
  InitializeSetup(body, x_hat, y_hat, r, r_norm, r2, one_over_r3, s);
  return GeopotentialGenericPotential(g, max_degree, is_zonal, s);
}

}  // namespace pb
