// DiscreteTrajectorySegment::Append with downsampling, the sink of the flows
// (physics/discrete_trajectory_segment_body.hpp:414-479): a new point replaces
// the last one of the timeline as long as the accumulated bound on the error of
// the Hermite interpolation stays below the tolerance.  Arithmetic of
// numerics/hermite3_body.hpp:154-194,237-271,289-297 (MakePolynomial,
// operator-, LInfinityL₂Norm), polynomial_in_monomial_basis_body.hpp:687-705,
// 843-870 (products, accumulated from zero, i outer, j inner),
// polynomial_evaluators_body.hpp:220-241 (Horner with FMA: the default
// evaluator up to degree 3) and root_finders_body.hpp:459-502
// (SolveQuadraticEquation with a TwoProduct discriminant).
#pragma once

#include "pb_math.cuh"

namespace pb {

PB_HD Vec3 FmaVec3(Vec3 const& a, double const x, Vec3 const& b) {
  return {fma(a.x, x, b.x), fma(a.y, x, b.y), fma(a.z, x, b.z)};
}

// Hermite3<Position, Instant>; the polynomial has its origin at `lower`.
struct Hermite3 {
  double lower;
  double upper;
  Vec3 a0, a1, a2, a3;
};

// Hermite3::MakePolynomial, hermite3_body.hpp:237-271.
PB_HD Hermite3 MakeHermite3(double const t1, double const t2, Vec3 const& q1,
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

PB_HD Vec3 Hermite3Evaluate(Hermite3 const& h, double const t) {
  double const x = t - h.lower;
  return FmaVec3(FmaVec3(FmaVec3(h.a3, x, h.a2), x, h.a1), x, h.a0);
}

// polynomial_evaluators_body.hpp:243-262: b = get<low>(coefficients) * low.
PB_HD Vec3 Hermite3EvaluateDerivative(Hermite3 const& h, double const t) {
  double const x = t - h.lower;
  return FmaVec3(FmaVec3(h.a3 * 3, x, h.a2 * 2), x, h.a1 * 1);
}

PB_HD DP TwoProduct(double const a, double const b) {
  DP r;
  r.value = a * b;
  r.error = fma(a, b, -r.value);
  return r;
}

PB_HD int SolveQuadraticEquation(double const origin, double const a0,
                                 double const a1, double const a2,
                                 double roots[2]) {
  DP const discriminant =
      Subtract(TwoProduct(a1, a1), TwoProduct(4.0 * a0, a2));
  if (discriminant.value == 0.0) {
    roots[0] = origin - 0.5 * a1 / a2;
    return 1;
  } else if (discriminant.value < 0.0) {
    return 0;
  } else {
    double numerator;
    if (a1 > 0.0) {
      numerator = -a1 - sqrt(discriminant.value);
    } else {
      numerator = -a1 + sqrt(discriminant.value);
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

// Hermite3::LInfinityL₂Norm of a polynomial that vanishes with its derivative
// at `lower` (the difference of two interpolations from the same point).
PB_HD double Hermite3LInfinityL2Norm(Hermite3 const& h) {
  Vec3 const q0 = h.a2;
  Vec3 const q1 = h.a3;
  Vec3 const q_derivative0 = 1 * q1;
  Vec3 const zero{0.0, 0.0, 0.0};
  Vec3 product0 = zero;
  Vec3 product1 = zero;
  product0 += 0.0 * q_derivative0;
  product1 += 0.5 * q_derivative0;
  Vec3 const r0 = q0 + product0;
  Vec3 const r1 = q1 + product1;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0;
  c0 += Dot(q0, r0);
  c1 += Dot(q0, r1);
  c1 += Dot(q1, r0);
  c2 += Dot(q1, r1);
  double roots[2] = {0.0, 0.0};
  int const n_roots = SolveQuadraticEquation(h.lower, c0, c1, c2, roots);
  double max = Norm2(Hermite3Evaluate(h, h.upper));
  for (int i = 0; i < n_roots; ++i) {
    if (h.lower < roots[i] && roots[i] < h.upper) {
      double const candidate = Norm2(Hermite3Evaluate(h, roots[i]));
      max = (max < candidate) ? candidate : max;
    }
  }
  return sqrt(max);
}

// The timelines of n trajectories, [slot][trajectory] (times, errors) and
// [slot][q xyz, v xyz][trajectory] (points), with the state of
// DiscreteTrajectorySegment that Append needs: the size of the timeline, the
// downsampling error and the coefficients a2, a3 of the last interpolation
// (its a0, a1 are the point it starts from).
struct DownsamplerView {
  long long n;
  long long capacity;
  double tolerance;  // Negative: no downsampling.
  double* times;
  double* points;
  double* errors;
  long long* counts;
  double* downsampling_error;
  double* last_a2a3;  // [6][n].
  int32_t* overflow;  // [n]: an append was dropped for lack of room.
};

// Append for trajectory i.
PB_HD void DownsamplerAppend(DownsamplerView const& d, long long const i,
                             double const t, Vec3 const& q, Vec3 const& v) {
  long long const n = d.n;
  long long count = d.counts[i];
  auto store = [&](long long const slot, double const error) {
    d.times[slot * n + i] = t;
    double* const point = d.points + slot * 6 * n + i;
    point[0] = q.x;
    point[n] = q.y;
    point[2 * n] = q.z;
    point[3 * n] = v.x;
    point[4 * n] = v.y;
    point[5 * n] = v.z;
    d.errors[slot * n + i] = error;
  };
  auto load = [&](long long const slot, double& time, Vec3& position,
                  Vec3& velocity) {
    time = d.times[slot * n + i];
    double const* const point = d.points + slot * 6 * n + i;
    position = {point[0], point[n], point[2 * n]};
    velocity = {point[3 * n], point[4 * n], point[5 * n]};
  };
  if (count == 0) {
    if (d.capacity < 1) {
      d.overflow[i] = 1;
      return;
    }
    store(0, 0.0);
    d.counts[i] = 1;
    return;
  }
  if (d.times[i] == t) {
    return;  // "Append at existing time".
  }
  double last_time;
  Vec3 last_q, last_v;
  load(count - 1, last_time, last_q, last_v);
  if (count > 1 && d.tolerance >= 0.0) {
    double lower_time;
    Vec3 lower_q, lower_v;
    load(count - 2, lower_time, lower_q, lower_v);
    Hermite3 const hermite3 =
        MakeHermite3(lower_time, t, lower_q, q, lower_v, v);
    Hermite3 last_hermite3;
    last_hermite3.lower = lower_time;
    last_hermite3.upper = last_time;
    last_hermite3.a0 = lower_q;
    last_hermite3.a1 = lower_v;
    last_hermite3.a2 = {d.last_a2a3[i], d.last_a2a3[n + i],
                        d.last_a2a3[2 * n + i]};
    last_hermite3.a3 = {d.last_a2a3[3 * n + i], d.last_a2a3[4 * n + i],
                        d.last_a2a3[5 * n + i]};
    // operator-(Hermite3, Hermite3).
    Hermite3 difference;
    difference.lower = last_hermite3.lower;
    difference.upper =
        hermite3.upper < last_hermite3.upper ? hermite3.upper
                                             : last_hermite3.upper;
    difference.a0 = hermite3.a0 - last_hermite3.a0;
    difference.a1 = hermite3.a1 - last_hermite3.a1;
    difference.a2 = hermite3.a2 - last_hermite3.a2;
    difference.a3 = hermite3.a3 - last_hermite3.a3;
    double downsampling_error = d.downsampling_error[i];
    downsampling_error += Hermite3LInfinityL2Norm(difference);
    d.downsampling_error[i] = downsampling_error;
    if (downsampling_error < d.tolerance) {
      store(count - 1, downsampling_error);
      d.last_a2a3[i] = hermite3.a2.x;
      d.last_a2a3[n + i] = hermite3.a2.y;
      d.last_a2a3[2 * n + i] = hermite3.a2.z;
      d.last_a2a3[3 * n + i] = hermite3.a3.x;
      d.last_a2a3[4 * n + i] = hermite3.a3.y;
      d.last_a2a3[5 * n + i] = hermite3.a3.z;
      return;
    }
  }
  if (count >= d.capacity) {
    d.overflow[i] = 1;
    return;
  }
  Hermite3 const hermite3 = MakeHermite3(last_time, t, last_q, q, last_v, v);
  d.downsampling_error[i] = 0.0;
  store(count, 0.0);
  d.last_a2a3[i] = hermite3.a2.x;
  d.last_a2a3[n + i] = hermite3.a2.y;
  d.last_a2a3[2 * n + i] = hermite3.a2.z;
  d.last_a2a3[3 * n + i] = hermite3.a3.x;
  d.last_a2a3[4 * n + i] = hermite3.a3.y;
  d.last_a2a3[5 * n + i] = hermite3.a3.z;
  d.counts[i] = count + 1;
}

}  // namespace pb

// Defined in pb_api_trajectory.cu: the view of a pb_downsampler_t for kernels
// of other translation units.
struct pb_downsampler;
pb::DownsamplerView PbDownsamplerView(pb_downsampler* downsampler);
