// ORACLE -- TEST INFRASTRUCTURE ONLY (see numerics.hpp).
//
// CPU restatement of physics::Ephemeris for the hot path
// (physics/ephemeris_body.hpp): body ordering, Prolong, the three acceleration
// routines, the potential, FlowWithFixedStep and FlowWithAdaptiveStep.
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "integrators.hpp"
#include "numerics.hpp"
#include "physics.hpp"

namespace orc {

enum FixedMethod : int32_t {
  kBlanesMoan2002SRKN14A = 0,
  kMcLachlanAtela1992Order5Optimal = 1,
  kQuinlan1999Order8A = 2,
  kQuinlanTremaine1990Order12 = 3,
};

// A fixed-step instance of either family.
class FixedStepInstance {
 public:
  FixedStepInstance(int32_t method, RightHandSide rhs, State state,
                    AppendState append_state, double step) {
    switch (method) {
      case kBlanesMoan2002SRKN14A:
        srkn_ = std::make_unique<SrknInstance>(BlanesMoan2002SRKN14A(), rhs,
                                               state, append_state, step);
        break;
      case kMcLachlanAtela1992Order5Optimal:
        srkn_ = std::make_unique<SrknInstance>(
            McLachlanAtela1992Order5Optimal(), rhs, state, append_state, step);
        break;
      case kQuinlan1999Order8A:
        slms_ = std::make_unique<SlmsInstance>(Quinlan1999Order8A(), rhs, state,
                                               append_state, step);
        break;
      default:
        slms_ = std::make_unique<SlmsInstance>(QuinlanTremaine1990Order12(),
                                               rhs, state, append_state, step);
    }
  }
  int32_t Solve(double t_final) {
    return srkn_ ? srkn_->Solve(t_final) : slms_->Solve(t_final);
  }
  State const& state() const { return srkn_ ? srkn_->state() : slms_->state(); }
  std::int64_t evaluations() const {
    return srkn_ ? srkn_->evaluations : slms_->evaluations;
  }
  SlmsInstance const* slms() const { return slms_.get(); }

 private:
  std::unique_ptr<SrknInstance> srkn_;
  std::unique_ptr<SlmsInstance> slms_;
};

constexpr double min_radius_tolerance = 0.99;  // ephemeris_body.hpp:64.

class Ephemeris {
 public:
  // ephemeris_body.hpp:94-170.  `bodies` in the caller's order (SolarSystem:
  // alphabetical); internally oblate bodies are front-inserted.
  Ephemeris(std::vector<Body> const& bodies, std::vector<Vec3> const& q,
            std::vector<Vec3> const& v, double initial_time,
            double fitting_tolerance, double geopotential_tolerance,
            int32_t method, double step)
      : step_(step) {
    State state;
    state.time.value = initial_time;
    for (std::size_t i = 0; i < bodies.size(); ++i) {
      auto body = std::make_unique<Body>(bodies[i]);
      body->ComputeAxes();
      auto trajectory =
          std::make_unique<ContinuousTrajectory>(step, fitting_tolerance);
      trajectory->Append(initial_time, q[i], v[i]);
      DP qd[3], vd[3];
      qd[0].value = q[i].x;
      qd[1].value = q[i].y;
      qd[2].value = q[i].z;
      vd[0].value = v[i].x;
      vd[1].value = v[i].y;
      vd[2].value = v[i].z;
      if (body->oblate) {
        geopotentials_.emplace(geopotentials_.begin(), body.get(),
                               geopotential_tolerance);
        order_.insert(order_.begin(), static_cast<int>(i));
        bodies_.insert(bodies_.begin(), std::move(body));
        trajectories_.insert(trajectories_.begin(), std::move(trajectory));
        state.positions.insert(state.positions.begin(), qd, qd + 3);
        state.velocities.insert(state.velocities.begin(), vd, vd + 3);
        ++number_of_oblate_bodies_;
      } else {
        order_.push_back(static_cast<int>(i));
        bodies_.push_back(std::move(body));
        trajectories_.push_back(std::move(trajectory));
        state.positions.insert(state.positions.end(), qd, qd + 3);
        state.velocities.insert(state.velocities.end(), vd, vd + 3);
        ++number_of_spherical_bodies_;
      }
    }
    instance_ = std::make_unique<FixedStepInstance>(
        method,
        [this](double t, std::vector<double> const& positions,
               std::vector<double>& accelerations) {
          return ComputeGravitationalAccelerationBetweenAllMassiveBodies(
              t, positions, accelerations);
        },
        state,
        [this](State const& s) { AppendMassiveBodiesState(s); }, step);
  }

  int size() const { return static_cast<int>(bodies_.size()); }
  std::vector<int> const& order() const { return order_; }
  Body const& body(int internal) const { return *bodies_[internal]; }
  Geopotential const& geopotential(int internal) const {
    return geopotentials_[internal];
  }
  int number_of_oblate_bodies() const { return number_of_oblate_bodies_; }
  ContinuousTrajectory& trajectory(int internal) {
    return *trajectories_[internal];
  }
  FixedStepInstance const& instance() const { return *instance_; }
  int32_t last_severe_integration_status() const { return severe_status_; }

  double t_min() const {
    double t = trajectories_[0]->t_min();
    for (auto const& trajectory : trajectories_) {
      t = std::max(t, trajectory->t_min());
    }
    return t;
  }
  double t_max() const {
    double t = trajectories_[0]->t_max();
    for (auto const& trajectory : trajectories_) {
      t = std::min(t, trajectory->t_max());
    }
    return t;
  }

  // :312-344.
  void Prolong(double t, std::int64_t max_ephemeris_steps =
                             std::numeric_limits<std::int64_t>::max()) {
    double const instance_time = instance_->state().time.value;
    double const desired_t_max =
        std::min(t, instance_time + max_ephemeris_steps * step_);
    double t_final;
    if (desired_t_max <= instance_time) {
      t_final = instance_time + step_;
    } else {
      t_final = desired_t_max;
    }
    while (t_max() < desired_t_max) {
      instance_->Solve(t_final);
      t_final += step_;
    }
  }

  // :1503-1552.  Flat arrays in internal order.
  int32_t ComputeGravitationalAccelerationBetweenAllMassiveBodies(
      double t, std::vector<double> const& positions,
      std::vector<double>& accelerations) const {
    int const n = size();
    std::vector<Vec3> q(n), a(n);
    for (int i = 0; i < n; ++i) {
      q[i] = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
    }
    int const n_oblate = number_of_oblate_bodies_;
    for (int b1 = 0; b1 < n_oblate; ++b1) {
      ByMassiveBodyOnMassiveBodies(true, true, t, b1, b1 + 1, n_oblate, q, a);
      ByMassiveBodyOnMassiveBodies(true, false, t, b1, n_oblate, n, q, a);
    }
    for (int b1 = n_oblate; b1 < n; ++b1) {
      ByMassiveBodyOnMassiveBodies(false, false, t, b1, b1 + 1, n, q, a);
    }
    for (int i = 0; i < n; ++i) {
      accelerations[3 * i] = a[i].x;
      accelerations[3 * i + 1] = a[i].y;
      accelerations[3 * i + 2] = a[i].z;
    }
    return kOk;
  }

  // :1249-1315: acceleration on one massive body (internal index b1).
  Vec3 ComputeGravitationalAccelerationOnMassiveBody(
      int b1, std::vector<Vec3> const& q, double t) const {
    int const n = size();
    int const n_oblate = number_of_oblate_bodies_;
    std::vector<Vec3> a(n);
    if (bodies_[b1]->oblate) {
      ByMassiveBodyOnMassiveBodies(true, true, t, b1, 0, b1, q, a);
      ByMassiveBodyOnMassiveBodies(true, true, t, b1, b1 + 1, n_oblate, q, a);
      ByMassiveBodyOnMassiveBodies(true, false, t, b1, n_oblate, n, q, a);
    } else {
      ByMassiveBodyOnMassiveBodies(false, true, t, b1, 0, n_oblate, q, a);
      ByMassiveBodyOnMassiveBodies(false, false, t, b1, n_oblate, b1, q, a);
      ByMassiveBodyOnMassiveBodies(false, false, t, b1, b1 + 1, n, q, a);
    }
    return a[b1];
  }

  // The point-mass tidal form −𝟙/‖Δq‖³ + 3 Δq⊗Δq/‖Δq‖⁵ of :546-553,
  // :1196-1203, :1234-1241, as the R3x3Matrix arithmetic it expands to
  // (geometry/symmetric_bilinear_form_body.hpp:279-357,406-412;
  // geometry/r3x3_matrix_body.hpp:277-306,354-398): (−𝟙)/n³ entry by entry
  // (the off-diagonal ones are −0/n³), (3·(Δq_i·Δq_j))/n⁵, then the sum.
  struct Form {
    double m[3][3];
  };
  static Form TidalForm(Vec3 const& dq) {
    double const dq2 = Norm2(dq);
    double const dq_norm = std::sqrt(dq2);
    double const dq_norm3 = dq2 * dq_norm;
    double const dq_norm5 = dq_norm3 * dq2;
    double const c[3] = {dq.x, dq.y, dq.z};
    Form form;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        double const identity = (i == j ? -1.0 : -0.0) / dq_norm3;
        double const square = (3 * (c[i] * c[j])) / dq_norm5;
        form.m[i][j] = identity + square;
      }
    }
    return form;
  }
  // form * Δv: a Dot per row (r3x3_matrix_body.hpp:327-337).
  static Vec3 Apply(Form const& f, Vec3 const& v) {
    return {f.m[0][0] * v.x + f.m[0][1] * v.y + f.m[0][2] * v.z,
            f.m[1][0] * v.x + f.m[1][1] * v.y + f.m[1][2] * v.z,
            f.m[2][0] * v.x + f.m[2][1] * v.y + f.m[2][2] * v.z};
  }

  // ComputeGravitationalJerkOnMasslessBody, :524-561 (point masses only).
  Vec3 ComputeGravitationalJerkOnMasslessBody(Vec3 const& q, Vec3 const& v,
                                              double t) const {
    Vec3 jerk{};
    for (int b2 = 0; b2 < size(); ++b2) {
      Vec3 const q2 = trajectories_[b2]->EvaluatePosition(t);
      Vec3 const v2 = trajectories_[b2]->EvaluateVelocity(t);
      Vec3 const dq = q - q2;
      Vec3 const dv = v - v2;
      Vec3 const vector = Apply(TidalForm(dq), dv);
      jerk += bodies_[b2]->mu * vector;
    }
    return jerk;
  }

  // ComputeGravitationalJerkOnMassiveBody, :1317-1340 with :1211-1247: the
  // bodies before b1, then those after it (internal order).
  Vec3 ComputeGravitationalJerkOnMassiveBody(int b1, std::vector<Vec3> const& q,
                                             std::vector<Vec3> const& v) const {
    Vec3 jerk{};
    for (int b2 = 0; b2 < size(); ++b2) {
      if (b2 == b1) {
        continue;
      }
      Vec3 const dq = q[b1] - q[b2];
      Vec3 const dv = v[b1] - v[b2];
      Vec3 const vector = Apply(TidalForm(dq), dv);
      jerk += bodies_[b2]->mu * vector;
    }
    return jerk;
  }

  // ComputeJacobianOnMassiveBody, :495-522 with :1174-1209.
  Form ComputeJacobianOnMassiveBody(int b1, std::vector<Vec3> const& q) const {
    Form jacobian{};
    for (int b2 = 0; b2 < size(); ++b2) {
      if (b2 == b1) {
        continue;
      }
      Form const form = TidalForm(q[b1] - q[b2]);
      double const mu2 = bodies_[b2]->mu;
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
          jacobian.m[i][j] = jacobian.m[i][j] + mu2 * form.m[i][j];
        }
      }
    }
    return jacobian;
  }

  std::vector<Vec3> EvaluateAllVelocities(double t) const {
    std::vector<Vec3> v;
    for (auto const& trajectory : trajectories_) {
      v.push_back(trajectory->EvaluateVelocity(t));
    }
    return v;
  }

  std::vector<Vec3> EvaluateAllPositions(double t) const {
    std::vector<Vec3> q;
    for (auto const& trajectory : trajectories_) {
      q.push_back(trajectory->EvaluatePosition(t));
    }
    return q;
  }

  // :1554-1591.
  int32_t ComputeGravitationalAccelerationByAllMassiveBodiesOnMasslessBodies(
      double t, std::vector<Vec3> const& positions,
      std::vector<Vec3>& accelerations) const {
    accelerations.assign(accelerations.size(), Vec3{});
    int32_t error = kOk;
    for (int b1 = 0; b1 < size(); ++b1) {
      error |= ByMassiveBodyOnMasslessBodies(b1 < number_of_oblate_bodies_, t,
                                             b1, positions, accelerations);
    }
    return error;
  }

  // :1593-1621, :1464-1501.
  void ComputeGravitationalPotentialsOfAllMassiveBodies(
      double t, std::vector<Vec3> const& positions,
      std::vector<double>& potentials) const {
    potentials.assign(potentials.size(), 0.0);
    for (int b1 = 0; b1 < size(); ++b1) {
      bool const oblate = b1 < number_of_oblate_bodies_;
      double const mu1 = bodies_[b1]->mu;
      Vec3 const position1 = trajectories_[b1]->EvaluatePosition(t);
      for (std::size_t b2 = 0; b2 < positions.size(); ++b2) {
        Vec3 const dq = position1 - positions[b2];
        double const dq2 = Norm2(dq);
        double const dq_norm = std::sqrt(dq2);
        double const one_over_dq_norm = 1 / dq_norm;
        potentials[b2] -= mu1 * one_over_dq_norm;
        if (oblate) {
          double const one_over_dq3 =
              one_over_dq_norm * one_over_dq_norm * one_over_dq_norm;
          double const effect =
              geopotentials_[b1].GeneralSphericalHarmonicsPotential(
                  t, -dq, dq_norm, dq2, one_over_dq3);
          potentials[b2] += mu1 * effect;
        }
      }
    }
  }

  // The massless right-hand side on flat arrays (:360-385, no intrinsic
  // accelerations).
  RightHandSide MasslessRightHandSide() const {
    return [this](double t, std::vector<double> const& q,
                  std::vector<double>& g) -> int32_t {
      std::size_t const n = q.size() / 3;
      std::vector<Vec3> positions(n), accelerations(n);
      for (std::size_t i = 0; i < n; ++i) {
        positions[i] = {q[3 * i], q[3 * i + 1], q[3 * i + 2]};
      }
      int32_t const error =
          ComputeGravitationalAccelerationByAllMassiveBodiesOnMasslessBodies(
              t, positions, accelerations);
      for (std::size_t i = 0; i < n; ++i) {
        g[3 * i] = accelerations[i].x;
        g[3 * i + 1] = accelerations[i].y;
        g[3 * i + 2] = accelerations[i].z;
      }
      return error == kOk ? kOk : kOutOfRange;
    };
  }

  // :1698-1717.
  static double ToleranceToErrorRatio(double length_tolerance,
                                      double speed_tolerance,
                                      StateError const& error) {
    double max_length_error = 0;
    double max_speed_error = 0;
    for (std::size_t i = 0; i + 2 < error.position_error.size(); i += 3) {
      Vec3 const e{error.position_error[i], error.position_error[i + 1],
                   error.position_error[i + 2]};
      max_length_error = std::max(max_length_error, Norm(e));
    }
    for (std::size_t i = 0; i + 2 < error.velocity_error.size(); i += 3) {
      Vec3 const e{error.velocity_error[i], error.velocity_error[i + 1],
                   error.velocity_error[i + 2]};
      max_speed_error = std::max(max_speed_error, Norm(e));
    }
    return std::min(length_tolerance / max_length_error,
                    speed_tolerance / max_speed_error);
  }

  // :1624-1696 for one massless trajectory; the ephemeris must already cover
  // [state.time, t] (no Prolong here: the caller prolongs).  `append` receives
  // every accepted state.
  // An inertially fixed burn, Manœuvre::ComputeIntrinsicAcceleration,
  // ksp_plugin/manœuvre_body.hpp:395-406.
  struct Burn {
    Vec3 direction;
    double thrust, initial_mass, mass_flow, initial_time, final_time;
    // >= 0: `direction` is expressed in the Frenet frame of the trajectory in
    // the body-centred non-rotating frame of this body (internal index):
    // Manœuvre::FrenetIntrinsicAcceleration, a velocity-dependent intrinsic
    // acceleration.  -1: inertially fixed.
    int frenet_centre = -1;
    Vec3 IntrinsicAcceleration(double const t) const {
      return IntrinsicAcceleration(t, direction);
    }
    // Manœuvre::ComputeIntrinsicAcceleration(t, direction),
    // ksp_plugin/manœuvre_body.hpp:395-406.
    Vec3 IntrinsicAcceleration(double const t,
                               Vec3 const& inertial_direction) const {
      if (t >= initial_time && t <= final_time) {
        return inertial_direction * thrust /
               (initial_mass - (t - initial_time) * mass_flow);
      } else {
        return {0.0, 0.0, 0.0};
      }
    }
  };

  // The inertial direction of a Frenet-frame burn for a trajectory at
  // (t, q, v) whose gravitational acceleration is `gravitational`:
  // Manœuvre::ComputeFrenetFrame (ksp_plugin/manœuvre_body.hpp:383-393) for a
  // BodyCentredNonRotatingReferenceFrame (physics/
  // body_centred_non_rotating_reference_frame_body.hpp:73-88,118-140):
  // ReferenceFrame::FrenetFrame (physics/reference_frame_body.hpp:38-54) has
  //   tangent  = Normalize(velocity),
  //   normal   = Normalize(acceleration.OrthogonalizationAgainst(velocity)),
  //   binormal = Wedge(tangent, normal),
  // velocity and geometric acceleration (RigidReferenceFrame::
  // GeometricAcceleration, rigid_reference_frame_body.hpp:59-78,364-398: the
  // gravitational acceleration at the point minus the acceleration of the
  // centre; the rotational terms vanish with the angular velocity) RELATIVE to
  // the centre.  The reference rotates the degrees of freedom into the axes of
  // the centre's celestial frame, builds the trihedron there and rotates the
  // direction back (two quaternion rotations, and the gravitational
  // acceleration re-evaluated at the position carried there and back); the
  // trihedron is covariant under that rotation, so it is formed here directly
  // in the inertial axes, from the acceleration the right-hand side has just
  // computed.  Same frame, same formulae; the results differ from the
  // reference's by the rounding of the two rotations (no known answer of the
  // reference pins this routine: its tests use mock frames).
  Vec3 FrenetDirection(Burn const& burn, double const t, Vec3 const& v,
                       Vec3 const& gravitational) const {
    int const n = size();
    std::vector<Vec3> positions(n);
    for (int b = 0; b < n; ++b) {
      positions[b] = trajectories_[b]->EvaluatePosition(t);
    }
    int const c = burn.frenet_centre;
    Vec3 const centre_velocity = trajectories_[c]->EvaluateVelocity(t);
    Vec3 const centre_acceleration =
        ComputeGravitationalAccelerationOnMassiveBody(c, positions, t);
    Vec3 const velocity = v - centre_velocity;
    Vec3 const acceleration = gravitational - centre_acceleration;
    Vec3 const tangent = velocity / Norm(velocity);
    // OrthogonalizationAgainst, geometry/r3_element_body.hpp:186-190.
    Vec3 const normal_acceleration =
        acceleration - Dot(acceleration, tangent) * tangent;
    Vec3 const normal = normal_acceleration / Norm(normal_acceleration);
    Vec3 const binormal = Cross(tangent, normal);
    // Rotation(tangent, normal, binormal) applied to the Frenet coordinates.
    return burn.direction.x * tangent + burn.direction.y * normal +
           burn.direction.z * binormal;
  }

  int32_t FlowWithAdaptiveStep(State& state, double t, double length_tolerance,
                               double speed_tolerance, std::int64_t max_steps,
                               AppendState append, std::int64_t* evaluations,
                               std::int64_t* rejections,
                               Burn const* burn = nullptr,
                               bool generalized = false) const {
    if (state.time.value == t) {
      return kOk;
    }
    double const t_final = std::min(t, t_max());
    AdaptiveParameters parameters;
    parameters.first_step = t_final - state.time.value;
    parameters.safety_factor = 0.9;
    parameters.max_steps = max_steps;
    parameters.last_step_is_exact = true;
    State last = state;
    // ephemeris_body.hpp:422-436: accelerations[0] += intrinsic_acceleration(t).
    RightHandSide const gravitation = MasslessRightHandSide();
    RightHandSide const rhs =
        burn == nullptr
            ? gravitation
            : RightHandSide([&gravitation, burn](
                                double const t, std::vector<double> const& q,
                                std::vector<double>& g) -> int32_t {
                int32_t const error = gravitation(t, q, g);
                Vec3 acceleration{g[0], g[1], g[2]};
                acceleration += burn->IntrinsicAcceleration(t);
                g[0] = acceleration.x;
                g[1] = acceleration.y;
                g[2] = acceleration.z;
                return error;
              });
    // The generalized overload (ephemeris_body.hpp:446-478) with Fine1987RKNG34
    // (the plugin's burn integrator, ksp_plugin/integrators.cpp:54-63).
    ErknInstance instance(
        generalized ? Fine1987RKNG34() : DormandElMikkawyPrince1986RKN434FM(),
        rhs, state,
        [&](State const& s) {
          last = s;
          if (append) {
            append(s);
          }
        },
        [=](double, State const&, StateError const& error) {
          return ToleranceToErrorRatio(length_tolerance, speed_tolerance,
                                       error);
        },
        parameters);
    if (generalized && burn != nullptr && burn->frenet_centre >= 0) {
      // ephemeris_body.hpp:446-478: the generalized overload, whose intrinsic
      // acceleration sees the degrees of freedom.
      instance.generalized_rhs =
          [this, &gravitation, burn](double const t,
                                     std::vector<double> const& q,
                                     std::vector<double> const& v,
                                     std::vector<double>& g) -> int32_t {
        int32_t const error = gravitation(t, q, g);
        Vec3 acceleration{g[0], g[1], g[2]};
        Vec3 const direction =
            FrenetDirection(*burn, t, Vec3{v[0], v[1], v[2]}, acceleration);
        acceleration += burn->IntrinsicAcceleration(t, direction);
        g[0] = acceleration.x;
        g[1] = acceleration.y;
        g[2] = acceleration.z;
        return error;
      };
    }
    int32_t status = instance.Solve(t_final);
    state = last;
    if (evaluations != nullptr) {
      *evaluations = instance.evaluations;
    }
    if (rejections != nullptr) {
      *rejections = instance.rejections;
    }
    if (status == kOutOfRange) {
      status = kOk;
    }
    if (status != kOk || t_final == t) {
      return status;
    } else {
      return kDeadlineExceeded;
    }
  }

 private:
  // :1342-1410.
  void ByMassiveBodyOnMassiveBodies(bool body1_is_oblate, bool body2_is_oblate,
                                    double t, int b1, int b2_begin, int b2_end,
                                    std::vector<Vec3> const& positions,
                                    std::vector<Vec3>& accelerations) const {
    Vec3 const& position_of_b1 = positions[b1];
    Vec3& acceleration_on_b1 = accelerations[b1];
    double const mu1 = bodies_[b1]->mu;
    for (int b2 = b2_begin; b2 < b2_end; ++b2) {
      Vec3& acceleration_on_b2 = accelerations[b2];
      double const mu2 = bodies_[b2]->mu;
      Vec3 const dq = position_of_b1 - positions[b2];
      double const dq2 = Norm2(dq);
      double const dq_norm = std::sqrt(dq2);
      double const one_over_dq3 = dq_norm / (dq2 * dq2);
      double const mu1_over_dq3 = mu1 * one_over_dq3;
      acceleration_on_b2 += dq * mu1_over_dq3;
      double const mu2_over_dq3 = mu2 * one_over_dq3;
      acceleration_on_b1 -= dq * mu2_over_dq3;
      if (body1_is_oblate) {
        Vec3 const effect =
            geopotentials_[b1].GeneralSphericalHarmonicsAcceleration(
                t, -dq, dq_norm, dq2, one_over_dq3);
        acceleration_on_b1 -= mu2 * effect;
        acceleration_on_b2 += mu1 * effect;
      }
      if (body2_is_oblate) {
        Vec3 const effect =
            geopotentials_[b2].GeneralSphericalHarmonicsAcceleration(
                t, dq, dq_norm, dq2, one_over_dq3);
        acceleration_on_b1 += mu2 * effect;
        acceleration_on_b2 -= mu1 * effect;
      }
    }
  }

  // :1412-1462.
  int32_t ByMassiveBodyOnMasslessBodies(bool body1_is_oblate, double t, int b1,
                                        std::vector<Vec3> const& positions,
                                        std::vector<Vec3>& accelerations) const {
    double const mu1 = bodies_[b1]->mu;
    Vec3 const position1 = trajectories_[b1]->EvaluatePosition(t);
    double const body1_collision_radius =
        min_radius_tolerance * bodies_[b1]->min_radius;
    int32_t error = kOk;
    for (std::size_t b2 = 0; b2 < positions.size(); ++b2) {
      Vec3 const dq = position1 - positions[b2];
      double const dq2 = Norm2(dq);
      double const dq_norm = std::sqrt(dq2);
      error |= dq_norm > body1_collision_radius ? kOk : kOutOfRange;
      double const one_over_dq3 = dq_norm / (dq2 * dq2);
      double const mu1_over_dq3 = mu1 * one_over_dq3;
      accelerations[b2] += dq * mu1_over_dq3;
      if (body1_is_oblate) {
        Vec3 const effect =
            geopotentials_[b1].GeneralSphericalHarmonicsAcceleration(
                t, -dq, dq_norm, dq2, one_over_dq3);
        accelerations[b2] += mu1 * effect;
      }
    }
    return error;
  }

  // :1071-1096, :1098-1116.
  void AppendMassiveBodiesState(State const& state) {
    double const time = state.time.value;
    for (int i = 0; i < size(); ++i) {
      Vec3 const q{state.positions[3 * i].value, state.positions[3 * i + 1].value,
                   state.positions[3 * i + 2].value};
      Vec3 const v{state.velocities[3 * i].value,
                   state.velocities[3 * i + 1].value,
                   state.velocities[3 * i + 2].value};
      int32_t const status = trajectories_[i]->Append(time, q, v);
      if (status != kOk) {
        severe_status_ = status;
      }
    }
  }

  double step_;
  std::vector<std::unique_ptr<Body>> bodies_;
  std::vector<std::unique_ptr<ContinuousTrajectory>> trajectories_;
  std::vector<Geopotential> geopotentials_;
  std::vector<int> order_;  // internal index -> caller index.
  int number_of_oblate_bodies_ = 0;
  int number_of_spherical_bodies_ = 0;
  std::unique_ptr<FixedStepInstance> instance_;
  int32_t severe_status_ = kOk;
};

}  // namespace orc
