// Device data model of an ephemeris: plain structs, SoA arrays in HBM.
// See DESIGN.md "Data layout in HBM".
#pragma once

#include <cstdint>

#include "pb_math.cuh"

namespace pb {

constexpr int kMaxGeopotentialDegree = 50;
constexpr int kGeopotentialSize = kMaxGeopotentialDegree + 1;
constexpr int kSegmentCoefficients = 18;  // Degrees 3..17.

// physics/harmonic_damping.hpp:55-64.
struct Damping {
  double inner_threshold;
  double outer_threshold;
  double c1, c2, c3;
};

// One massive body in INTERNAL order (oblate bodies first, see
// physics/ephemeris_body.hpp:138-158).
struct DeviceBody {
  double mu;
  double collision_radius;  // 0.99 * min_radius, ephemeris_body.hpp:1425-1426.
  double reference_radius;
  Vec3 polar_axis;    // rotating_body_body.hpp:65-81.
  Vec3 equatorial;
  Vec3 biequatorial;
  // FromSurfaceFrame(t) = (AA(pi/2 + alpha, z) AA(pi/2 - delta, x)) AA(W(t), z),
  // rotating_body.hpp:181-191; the parenthesis is time-independent.
  double q_ab_real;
  Vec3 q_ab_imaginary;
  double reference_angle;
  double reference_instant;
  double angular_frequency;
  Damping sectoral;
  int32_t is_oblate;
  int32_t degree;
  int32_t is_zonal;
  int32_t n_damping;           // degree_damping_.size() = degree + 1.
  int32_t damping_offset;      // Into EphemerisView::dampings.
  int32_t coefficient_offset;  // Into EphemerisView::cos / sin.
  int32_t frame_slot;          // Slot in the stage table, -1 if zonal/spherical.
  int32_t pad;
};

// Read-only view of everything a kernel needs; passed by value.
struct EphemerisView {
  int32_t n_bodies;
  int32_t n_oblate;
  int32_t n_frames;  // Oblate, non-zonal bodies.
  int32_t stage_stride;  // Doubles per time in a stage table.
  // 1 if all bodies have the same segment boundaries (t_max arrays).
  int32_t uniform_segments;
  // Body b's segments start at b * segment_capacity (== segment_offset[b]).
  int32_t segment_capacity;
  // 1: the ensemble kernels may read the bodies from their constant bank
  // (c_ensemble, pb_kernels_nbody.cuh): set by a caller that holds it.
  int32_t ensemble_tables;
  DeviceBody const* bodies;
  Damping const* dampings;
  double const* cos;  // Normalised, (n, m) -> n(n+1)/2 + m per body.
  double const* sin;
  // Legendre normalisation factors, (n, m) -> n(n+1)/2 + m.
  double const* normalization;
  // ContinuousTrajectory segments, per body [segment_offset[b],
  // segment_end[b]).
  int32_t const* segment_offset;
  int32_t const* segment_end;
  double const* segment_t_max;
  double const* segment_origin;
  int32_t const* segment_degree;
  // [segment][xyz][18]: the host API's [segment][18][3] transposed, so that the
  // coefficients of one coordinate are contiguous (16-byte loads of pairs).
  double const* segment_coefficients;
  // Caller index -> internal index.
  int32_t const* internal_of_caller;
};

// Set by FlowPathKernel when some probe of a fixed-step flow is within the
// range of the harmonics of degree > 3 of a body other than the hot one (the
// thread-local-memory path: same bits, several times slower).
constexpr int32_t kPathGenericHarmonics = 0x100;

// Layout of one entry of a stage table (all bodies evaluated at one time):
//   [0, 3 n_bodies)            positions, body-major xyz, internal order;
//   [3 n_bodies, + 6 n_frames) surface-frame x and y axes of the bodies that
//                              have a frame_slot.
PB_HD int StageStride(int const n_bodies, int const n_frames) {
  // Even: the entries are copied 16 bytes at a time (SrknFlowKernel).
  return (3 * n_bodies + 6 * n_frames + 1) / 2 * 2;
}

// Tableau of a symplectic Runge-Kutta-Nyström method, integrators/methods.hpp.
struct SrknTableau {
  int stages;
  int first_stage;  // 0 for BA, 1 for ABA.
  double a[16];
  double b[16];
};

// Per-body scalars that every thread of a flow kernel reads at the same
// address; kernels keep a copy in the constant bank.
constexpr int kMaxConstantBodies = 64;

struct BodyScalars {
  double mu[kMaxConstantBodies];
  double collision_radius[kMaxConstantBodies];
  // Geopotential::LimitingDegree returns 2 (no harmonics) iff
  // !(r < degree_damping[2].outer_threshold); -infinity for spherical bodies.
  double outer_threshold_degree_2[kMaxConstantBodies];
};

constexpr int kMaxMultistepOrder = 12;
// symmetric_linear_multistep_integrator_body.hpp:20.
constexpr int kStartupStepDivisor = 16;

// A symmetric linear multistep method with its velocity formula,
// integrators/methods.hpp:998-1007,1040-1055 and
// cohen_hubbard_oesterwinter_body.hpp.
struct MultistepMethod {
  int order;
  double alpha[kMaxMultistepOrder / 2 + 1];
  double beta_numerator[kMaxMultistepOrder / 2 + 1];
  double beta_denominator;
  double cho_numerators[kMaxMultistepOrder];
  double cho_denominator;
};

}  // namespace pb
