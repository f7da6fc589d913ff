// Host-side construction of the device model from the C-ABI body description:
// the parts of Ephemeris::Ephemeris (physics/ephemeris_body.hpp:94-170),
// RotatingBody (physics/rotating_body_body.hpp:65-81) and
// Geopotential::Geopotential (physics/geopotential_body.hpp:669-733) that run
// once per ephemeris.  Plain C++ (no CUDA) so that the host-logic tests can
// build it without a GPU.
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <vector>

#include "../../include/principia_b200.h"
#include "generated_tables.h"
#include "pb_math.cuh"
#include "pb_model.h"

namespace pb {

constexpr double kPi = 3.14159265358979323846264338327950288419716939937511;
constexpr double kMinRadiusTolerance = 0.99;  // ephemeris_body.hpp:64.

struct HostModel {
  std::vector<int32_t> order;               // internal -> caller.
  std::vector<int32_t> internal_of_caller;  // caller -> internal.
  int32_t n_oblate = 0;
  int32_t n_frames = 0;
  std::vector<DeviceBody> bodies;  // Internal order.
  std::vector<Damping> dampings;
  std::vector<double> cos, sin;
};

// HarmonicDamping::HarmonicDamping, harmonic_damping_body.hpp:16-23.
inline Damping MakeDamping(double const inner_threshold) {
  Damping d;
  d.outer_threshold = inner_threshold * 3;
  d.inner_threshold = inner_threshold;
  d.c1 = 9 / (4 * inner_threshold);
  d.c2 = -3 / (2 * PowInt(inner_threshold, 2));
  d.c3 = 1 / (4 * PowInt(inner_threshold, 3));
  return d;
}

inline Damping DefaultDamping() {
  Damping d;
  d.outer_threshold = std::numeric_limits<double>::infinity();
  d.inner_threshold = std::numeric_limits<double>::infinity();
  d.c1 = d.c2 = d.c3 = 0;
  return d;
}

// SphericalCoordinates::ToCartesian, geometry/r3_element_body.hpp:213-219.
inline Vec3 ToCartesian(double const radius, double const latitude,
                        double const longitude) {
  double sin_latitude, cos_latitude, sin_longitude, cos_longitude;
  CrSinCos(latitude, sin_latitude, cos_latitude);
  CrSinCos(longitude, sin_longitude, cos_longitude);
  return {radius * cos_longitude * cos_latitude,
          radius * sin_longitude * cos_latitude, radius * sin_latitude};
}

struct HostQuaternion {
  double real;
  Vec3 imaginary;
};
inline HostQuaternion AngleAxis(double const angle, Vec3 const& axis) {
  double const half_angle = 0.5 * angle;
  double s, c;
  CrSinCos(half_angle, s, c);
  return {c, s * axis};
}
inline HostQuaternion Multiply(HostQuaternion const& l,
                               HostQuaternion const& r) {
  return {l.real * r.real - Dot(l.imaginary, r.imaginary),
          l.real * r.imaginary + r.real * l.imaginary +
              Cross(l.imaginary, r.imaginary)};
}

// Geopotential::Geopotential, geopotential_body.hpp:669-733.
inline void ComputeDampings(pb_body_t const& b, double const tolerance,
                            std::vector<Damping>& degree_damping,
                            Damping& sectoral_damping) {
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
  for (int n = 2; n <= b.geopotential_degree; ++n) {
    for (int m = 0; m <= n; ++m) {
      int const k = n * (n + 1) / 2 + m;
      double const max_abs_Pnm = pb_max_abs_normalized_legendre[k];
      double const Cnm = b.cos[k];
      double const Snm = b.sin[k];
      // Root(n, x) is std::pow(x, 1.0 / n) in the reference
      // (numerics/elementary_functions_body.hpp:157-159); this runs once per
      // body on the host, with the host's libm like the reference.
      double const r =
          Cnm == 0 && Snm == 0
              ? 0.0
              : b.reference_radius *
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
  sectoral_damping = DefaultDamping();
  degree_damping.clear();
  while (!harmonic_thresholds.empty()) {
    Threshold const threshold = harmonic_thresholds.top();
    if (threshold.n == 2 && threshold.m == 2) {
      if (degree_damping.size() > 3) {
        sectoral_damping = MakeDamping(degree_damping[3].inner_threshold);
      } else {
        sectoral_damping = MakeDamping(threshold.r);
      }
    }
    while (threshold.n >= static_cast<int>(degree_damping.size())) {
      degree_damping.push_back(MakeDamping(threshold.r));
    }
    harmonic_thresholds.pop();
  }
}

inline pb_status_t BuildHostModel(int32_t const n_bodies,
                                  pb_body_t const* const input,
                                  double const geopotential_tolerance,
                                  HostModel& model) {
  if (n_bodies <= 0 || input == nullptr || !(geopotential_tolerance >= 0)) {
    return PB_INVALID_ARGUMENT;
  }
  // Ephemeris::Ephemeris: oblate bodies are inserted at the front (so they end
  // up in reverse caller order), the others appended.
  model.order.clear();
  for (int32_t i = 0; i < n_bodies; ++i) {
    if (input[i].is_oblate) {
      if (input[i].geopotential_degree < 2 ||
          input[i].geopotential_degree > kMaxGeopotentialDegree ||
          input[i].cos == nullptr || input[i].sin == nullptr ||
          !input[i].is_rotating) {
        return PB_INVALID_ARGUMENT;
      }
      model.order.insert(model.order.begin(), i);
      ++model.n_oblate;
    } else {
      model.order.push_back(i);
    }
  }
  model.internal_of_caller.assign(n_bodies, 0);
  model.bodies.clear();
  model.dampings.clear();
  model.cos.clear();
  model.sin.clear();
  model.n_frames = 0;
  for (int32_t internal = 0; internal < n_bodies; ++internal) {
    pb_body_t const& b = input[model.order[internal]];
    model.internal_of_caller[model.order[internal]] = internal;
    DeviceBody d{};
    d.mu = b.gravitational_parameter;
    d.collision_radius =
        kMinRadiusTolerance * (b.is_rotating ? b.min_radius : 0.0);
    d.frame_slot = -1;
    d.sectoral = DefaultDamping();
    if (b.is_rotating) {
      d.polar_axis = ToCartesian(1.0, b.declination_of_pole,
                                 b.right_ascension_of_pole);
      d.biequatorial =
          ToCartesian(1.0, 0.0, kPi / 2 + b.right_ascension_of_pole);
      d.equatorial = Cross(d.biequatorial, d.polar_axis);
      HostQuaternion const q_ab = Multiply(
          AngleAxis(kPi / 2 + b.right_ascension_of_pole, {0.0, 0.0, 1.0}),
          AngleAxis(kPi / 2 - b.declination_of_pole, {1.0, 0.0, 0.0}));
      d.q_ab_real = q_ab.real;
      d.q_ab_imaginary = q_ab.imaginary;
      d.reference_angle = b.reference_angle;
      d.reference_instant = b.reference_instant;
      d.angular_frequency = b.angular_frequency;
    }
    if (b.is_oblate) {
      d.is_oblate = 1;
      d.degree = b.geopotential_degree;
      d.is_zonal = b.is_zonal ? 1 : 0;
      d.reference_radius = b.reference_radius;
      std::vector<Damping> degree_damping;
      ComputeDampings(b, geopotential_tolerance, degree_damping, d.sectoral);
      d.n_damping = static_cast<int32_t>(degree_damping.size());
      d.damping_offset = static_cast<int32_t>(model.dampings.size());
      model.dampings.insert(model.dampings.end(), degree_damping.begin(),
                            degree_damping.end());
      d.coefficient_offset = static_cast<int32_t>(model.cos.size());
      int const size = (d.degree + 1) * (d.degree + 2) / 2;
      model.cos.insert(model.cos.end(), b.cos, b.cos + size);
      model.sin.insert(model.sin.end(), b.sin, b.sin + size);
      if (!d.is_zonal) {
        d.frame_slot = model.n_frames++;
      }
    }
    model.bodies.push_back(d);
  }
  return PB_OK;
}

}  // namespace pb
