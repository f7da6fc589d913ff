// Acceleration routines of physics::Ephemeris as host/device functions over an
// EphemerisView and one stage-table entry (all massive bodies evaluated at one
// instant).
#pragma once

#include "pb_geopotential.cuh"
#include "pb_math.cuh"
#include "pb_model.h"

namespace pb {

PB_HD GeopotentialBody MakeGeopotentialBody(EphemerisView const& e,
                                            int const b) {
  DeviceBody const* const body = e.bodies + b;
  GeopotentialBody g;
  g.body = body;
  g.degree_damping = e.dampings + body->damping_offset;
  g.cos = e.cos + body->coefficient_offset;
  g.sin = e.sin + body->coefficient_offset;
  g.normalization = e.normalization;
  return g;
}

// ContinuousTrajectory::FindPolynomialForInstantLocked +
// EvaluatePositionLocked, physics/continuous_trajectory_body.hpp:686-695,
// 860-885: the first segment with t <= t_max, evaluated at t - origin.
PB_HD int FindSegment(double const* t_max, int const begin, int const end,
                      double const t) {
  int lo = begin;
  int hi = end;
  while (lo < hi) {
    int const mid = (lo + hi) / 2;
    if (t_max[mid] < t) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo < end ? lo : end - 1;
}

PB_HD Vec3 EvaluateSegment(EphemerisView const& e, int const segment,
                           double const t) {
  double const x = t - e.segment_origin[segment];
  double const* const c =
      e.segment_coefficients + static_cast<long long>(segment) * 3 *
                                   kSegmentCoefficients;
  int const degree = e.segment_degree[segment];
  return {EstrinEvaluate<1>(c, degree, x),
          EstrinEvaluate<1>(c + kSegmentCoefficients, degree, x),
          EstrinEvaluate<1>(c + 2 * kSegmentCoefficients, degree, x)};
}

// EvaluateSegment as ONE straight-line tree over blocks of 8 coefficient
// slots, whatever the degree: compact code for kernels whose lanes evaluate
// segments of different degrees at once.  It relies on the slots above the
// degree holding -0.0 (pb_ephemeris_append_segments pads them): the reference's
// tree (polynomial_evaluators_body.hpp:131-160) carries a lone trailing block up
// unchanged, and fma(-0.0, x^m, low) == low for every low, signed zeros
// included, as long as x^m is finite.  Only the leaves need care, because x
// itself may be negative: a leaf whose odd slot is padding returns its even
// slot.  Same operations on the real coefficients, same bits.
PB_HD double EstrinLeaf(double const* const c, int const k, int const degree,
                        double const x) {
  // Slots k (even) and k + 1 of one coordinate: one 16-byte load.
#if defined(__CUDA_ARCH__)
  double2 const pair = __ldg(reinterpret_cast<double2 const*>(c + k));
  double const even = pair.x;
  double const odd = pair.y;
#else
  double const even = c[k];
  double const odd = c[k + 1];
#endif
  return k + 1 <= degree ? fma(odd, x, even) : even;
}

PB_HD double EstrinBlock8(double const* const c, int const base,
                          int const degree, double const x, double const x2,
                          double const x4) {
  double const l0 = EstrinLeaf(c, base, degree, x);
  double const l1 = EstrinLeaf(c, base + 2, degree, x);
  double const l2 = EstrinLeaf(c, base + 4, degree, x);
  double const l3 = EstrinLeaf(c, base + 6, degree, x);
  return fma(fma(l3, x2, l2), x4, fma(l1, x2, l0));
}

PB_HD Vec3 EvaluateSegmentPadded(EphemerisView const& e, int const segment,
                                 double const t) {
  static_assert(kSegmentCoefficients == 18, "two blocks of 8 and one leaf");
  double const x = t - e.segment_origin[segment];
  double const* const cx =
      e.segment_coefficients + static_cast<long long>(segment) * 3 *
                                   kSegmentCoefficients;
  double const* const cy = cx + kSegmentCoefficients;
  double const* const cz = cy + kSegmentCoefficients;
  int const degree = e.segment_degree[segment];
  double const x2 = x * x;
  double const x4 = x2 * x2;
  Vec3 p{EstrinBlock8(cx, 0, degree, x, x2, x4),
         EstrinBlock8(cy, 0, degree, x, x2, x4),
         EstrinBlock8(cz, 0, degree, x, x2, x4)};
  if (degree >= 8) {
    double const x8 = x4 * x4;
    p.x = fma(EstrinBlock8(cx, 8, degree, x, x2, x4), x8, p.x);
    p.y = fma(EstrinBlock8(cy, 8, degree, x, x2, x4), x8, p.y);
    p.z = fma(EstrinBlock8(cz, 8, degree, x, x2, x4), x8, p.z);
    if (degree >= 16) {
      double const x16 = x8 * x8;
      p.x = fma(EstrinLeaf(cx, 16, degree, x), x16, p.x);
      p.y = fma(EstrinLeaf(cy, 16, degree, x), x16, p.y);
      p.z = fma(EstrinLeaf(cz, 16, degree, x), x16, p.z);
    }
  }
  return p;
}

// Fills one stage-table entry for body b at time t.
PB_HD void FillStageEntry(EphemerisView const& e, int const b, double const t,
                          double* const entry) {
  int const segment = FindSegment(e.segment_t_max, e.segment_offset[b],
                                  e.segment_end[b], t);
  Vec3 const position = EvaluateSegment(e, segment, t);
  entry[3 * b] = position.x;
  entry[3 * b + 1] = position.y;
  entry[3 * b + 2] = position.z;
  int const slot = e.bodies[b].frame_slot;
  if (slot >= 0) {
    Vec3 x_hat, y_hat;
    SurfaceFrameAxes(e.bodies[b], t, x_hat, y_hat);
    double* const frame = entry + 3 * e.n_bodies + 6 * slot;
    frame[0] = x_hat.x;
    frame[1] = x_hat.y;
    frame[2] = x_hat.z;
    frame[3] = y_hat.x;
    frame[4] = y_hat.y;
    frame[5] = y_hat.z;
  }
}

// Ephemeris::ComputeGravitationalAccelerationByAllMassiveBodiesOnMasslessBodies
// for one probe, physics/ephemeris_body.hpp:1554-1591 with the inner routine
// :1412-1462.  Returns the OR of the status codes (0 or 11).
template<bool kUnrolled>
PB_HD int32_t AccelerationOnMasslessBody(EphemerisView const& e,
                                         double const* const stage,
                                         Vec3 const& q, Vec3& acceleration) {
  acceleration = {0.0, 0.0, 0.0};
  int32_t error = 0;
  for (int b1 = 0; b1 < e.n_bodies; ++b1) {
    DeviceBody const& body1 = e.bodies[b1];
    double const mu1 = body1.mu;
    Vec3 const position1{stage[3 * b1], stage[3 * b1 + 1], stage[3 * b1 + 2]};
    Vec3 const dq = position1 - q;
    double const dq2 = Norm2(dq);
    double const dq_norm = sqrt(dq2);
    error |= dq_norm > body1.collision_radius ? 0 : 11;
    double const one_over_dq3 = dq_norm / (dq2 * dq2);
    double const mu1_over_dq3 = mu1 * one_over_dq3;
    acceleration += dq * mu1_over_dq3;
    if (b1 < e.n_oblate) {
      GeopotentialBody const g = MakeGeopotentialBody(e, b1);
      FrameSource const frame{
          body1.frame_slot >= 0
              ? stage + 3 * e.n_bodies + 6 * body1.frame_slot
              : nullptr,
          0.0};
      Vec3 const effect = GeneralSphericalHarmonicsAcceleration<kUnrolled>(
          g, frame, -dq, dq_norm, dq2, one_over_dq3);
      acceleration += mu1 * effect;
    }
  }
  return error;
}

// ComputeGravitationalPotentialsOfAllMassiveBodies for one point,
// ephemeris_body.hpp:1464-1501,1593-1621.
PB_HD double PotentialOnMasslessBody(EphemerisView const& e,
                                     double const* const stage, Vec3 const& q) {
  double potential = 0.0;
  for (int b1 = 0; b1 < e.n_bodies; ++b1) {
    DeviceBody const& body1 = e.bodies[b1];
    double const mu1 = body1.mu;
    Vec3 const position1{stage[3 * b1], stage[3 * b1 + 1], stage[3 * b1 + 2]};
    Vec3 const dq = position1 - q;
    double const dq2 = Norm2(dq);
    double const dq_norm = sqrt(dq2);
    double const one_over_dq_norm = 1 / dq_norm;
    potential -= mu1 * one_over_dq_norm;
    if (b1 < e.n_oblate) {
      double const one_over_dq3 =
          one_over_dq_norm * one_over_dq_norm * one_over_dq_norm;
      GeopotentialBody const g = MakeGeopotentialBody(e, b1);
      FrameSource const frame{
          body1.frame_slot >= 0
              ? stage + 3 * e.n_bodies + 6 * body1.frame_slot
              : nullptr,
          0.0};
      double const effect = GeneralSphericalHarmonicsPotential(
          g, frame, -dq, dq_norm, dq2, one_over_dq3);
      potential += mu1 * effect;
    }
  }
  return potential;
}

// One (b1, b2) interaction of
// ComputeGravitationalAccelerationByMassiveBodyOnMassiveBodies,
// ephemeris_body.hpp:1342-1410, b1 < b2 in internal order.  `frames` holds the
// surface-frame axes of the bodies with a frame slot (6 doubles per slot).
// Updates both accelerations (Newton's third law) in the reference's order.
template<bool kUnrolled>
PB_HD void MassivePairInteraction(EphemerisView const& e, int const b1,
                                  int const b2, double const* const frames,
                                  Vec3 const& position_of_b1,
                                  Vec3 const& position_of_b2,
                                  Vec3& acceleration_on_b1,
                                  Vec3& acceleration_on_b2) {
  DeviceBody const& body1 = e.bodies[b1];
  DeviceBody const& body2 = e.bodies[b2];
  double const mu1 = body1.mu;
  double const mu2 = body2.mu;
  Vec3 const dq = position_of_b1 - position_of_b2;
  double const dq2 = Norm2(dq);
  double const dq_norm = sqrt(dq2);
  double const one_over_dq3 = dq_norm / (dq2 * dq2);
  double const mu1_over_dq3 = mu1 * one_over_dq3;
  acceleration_on_b2 += dq * mu1_over_dq3;
  double const mu2_over_dq3 = mu2 * one_over_dq3;
  acceleration_on_b1 -= dq * mu2_over_dq3;
  if (body1.is_oblate) {
    GeopotentialBody const g = MakeGeopotentialBody(e, b1);
    FrameSource const frame{
        body1.frame_slot >= 0 ? frames + 6 * body1.frame_slot : nullptr, 0.0};
    Vec3 const effect = GeneralSphericalHarmonicsAcceleration<kUnrolled>(
        g, frame, -dq, dq_norm, dq2, one_over_dq3);
    acceleration_on_b1 -= mu2 * effect;
    acceleration_on_b2 += mu1 * effect;
  }
  if (body2.is_oblate) {
    GeopotentialBody const g = MakeGeopotentialBody(e, b2);
    FrameSource const frame{
        body2.frame_slot >= 0 ? frames + 6 * body2.frame_slot : nullptr, 0.0};
    Vec3 const effect = GeneralSphericalHarmonicsAcceleration<kUnrolled>(
        g, frame, dq, dq_norm, dq2, one_over_dq3);
    acceleration_on_b1 += mu2 * effect;
    acceleration_on_b2 -= mu1 * effect;
  }
}

// The tabulated intrinsic acceleration of trajectory i at time t: an inertially
// fixed burn, Manœuvre::ComputeIntrinsicAcceleration,
// ksp_plugin/manœuvre_body.hpp:395-406.  burns: 8 planes of n.
PB_HD Vec3 BurnIntrinsicAcceleration(
    double const* const burns, long long const n,
    long long const i, double const t) {
  double const initial_time = burns[6 * n + i];
  double const final_time = burns[7 * n + i];
  if (t >= initial_time && t <= final_time) {
    Vec3 const direction{burns[i], burns[n + i], burns[2 * n + i]};
    double const thrust = burns[3 * n + i];
    double const initial_mass = burns[4 * n + i];
    double const mass_flow = burns[5 * n + i];
    return direction * thrust / (initial_mass - (t - initial_time) * mass_flow);
  } else {
    return {0.0, 0.0, 0.0};
  }
}

}  // namespace pb
