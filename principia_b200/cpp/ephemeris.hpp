// C++ facade over the C ABI with the surface of the reference's
// physics::Ephemeris<Frame> (physics/ephemeris.hpp:67-561) for the hot path:
// same method names, argument meaning and status behaviour.  Header-only C++17;
// link against libprincipia_b200.so.  Everything computes on the GPU.
//
//   reference                                     here
//   Ephemeris(bodies, initial_state, t0,          Ephemeris(bodies, initial_state,
//             AccuracyParameters,                           t0, AccuracyParameters,
//             FixedStepParameters)                          FixedStepParameters)
//   Prolong(t, max_ephemeris_steps)               Prolong(t, max_ephemeris_steps)
//   NewInstance(trajectories, intrinsic, params)  NewInstance(trajectories, params)
//   FlowWithFixedStep(t, instance)                FlowWithFixedStep(t, instance)
//   FlowWithAdaptiveStep(trajectory, intrinsic,   FlowWithAdaptiveStep(trajectory, t,
//       t, AdaptiveStepParameters, max_steps)         AdaptiveStepParameters, max_steps)
//   ComputeGravitationalAccelerationOnMasslessBody, ...OnMassiveBody,
//   ...OnMassiveBodies, ComputeGravitationalPotential, EvaluateAllPositions,
//   ComputeGravitationalJerkOnMasslessBody, ...OnMassiveBody/Bodies,
//   ComputeJacobianOnMassiveBody,
//   t_min, t_max: same names.
// Intrinsic accelerations: none, or a tabulated inertially fixed burn (host
// callbacks are not supported).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/principia_b200.h"
#include "continuous_trajectory.hpp"

namespace principia_b200 {

// absl::Status stand-in: the code is an absl::StatusCode number.
struct Status {
  pb_status_t code = PB_OK;
  std::string message;
  bool ok() const { return code == PB_OK; }
  static Status FromCall(pb_status_t const code) {
    return code == PB_OK ? Status{} : Status{code, pb_last_error()};
  }
};

using Position = double[3];

// physics/integration_parameters.hpp.
struct FixedStepParameters {
  int32_t integrator;  // PB_BLANES_MOAN_2002_SRKN_14A, ...
  double step;
};
struct AdaptiveStepParameters {
  int32_t integrator = PB_DORMAND_ELMIKKAWY_PRINCE_1986_RKN_434FM;
  std::int64_t max_steps = std::numeric_limits<std::int64_t>::max();
  double length_integration_tolerance = 1e-3;
  double speed_integration_tolerance = 1e-3;
};
struct AccuracyParameters {
  double fitting_tolerance;
  double geopotential_tolerance;
};

// The output sink of the massless flows: times and degrees of freedom
// appended in order (the role of DiscreteTrajectory<Frame>::Append).
struct DiscreteTrajectory {
  std::vector<double> times;
  std::vector<DegreesOfFreedom> degrees_of_freedom;
  void Append(double const t, DegreesOfFreedom const& dof) {
    times.push_back(t);
    degrees_of_freedom.push_back(dof);
  }
  double back_time() const { return times.back(); }
  DegreesOfFreedom const& back() const { return degrees_of_freedom.back(); }
};

class Ephemeris {
 public:
  // An Integrator<NewtonianMotionEquation>::Instance for massless bodies:
  // the ODE::State (SoA, with the compensated-summation errors) of the
  // trajectories it advances.
  class Instance {
   public:
    Instance() = default;
    Instance(Instance const&) = delete;
    Instance& operator=(Instance const&) = delete;
    ~Instance() { pb_fixed_step_instance_destroy(handle_); }
    double time() const { return time_[0]; }

   private:
    friend class Ephemeris;
    FixedStepParameters parameters_;
    std::vector<DiscreteTrajectory*> trajectories_;
    // The integrator instance of the C ABI: its ODE::State and, for the
    // multistep integrators, its previous steps live on the device.
    pb_fixed_step_instance_t* handle_ = nullptr;
    double time_[2] = {0.0, 0.0};
  };

  // physics/ephemeris_body.hpp:94-170.
  Ephemeris(std::vector<pb_body_t> const& bodies,
            std::vector<DegreesOfFreedom> const& initial_state,
            double const initial_time,
            AccuracyParameters const& accuracy_parameters,
            FixedStepParameters const& fixed_step_parameters,
            int32_t const cuda_device = 0)
      : accuracy_parameters_(accuracy_parameters),
        fixed_step_parameters_(fixed_step_parameters),
        n_bodies_(static_cast<int32_t>(bodies.size())) {
    if (bodies.empty() || bodies.size() != initial_state.size()) {
      throw std::invalid_argument("CHECK_EQ(bodies.size(), initial_state.size())");
    }
    Check(pb_ephemeris_create(n_bodies_, bodies.data(),
                              accuracy_parameters.geopotential_tolerance,
                              cuda_device, &ephemeris_));
    order_.resize(n_bodies_);
    Check(pb_ephemeris_internal_order(ephemeris_, order_.data(), nullptr));
    // The massive-body instance: one "ensemble" member, internal order.
    std::vector<double> state(12 * static_cast<std::size_t>(n_bodies_), 0.0);
    for (int32_t b = 0; b < n_bodies_; ++b) {
      DegreesOfFreedom const& dof = initial_state[order_[b]];
      for (int c = 0; c < 3; ++c) {
        state[(0 + c) * n_bodies_ + b] = dof.q[c];
        state[(6 + c) * n_bodies_ + b] = dof.v[c];
      }
      trajectories_.emplace_back(fixed_step_parameters.step,
                                 accuracy_parameters.fitting_tolerance);
      trajectories_.back().Append(initial_time, dof);
      uploaded_segments_.push_back(0);
    }
    Check(pb_nbody_ensemble_create(ephemeris_, fixed_step_parameters.integrator,
                                   fixed_step_parameters.step, 1, state.data(),
                                   initial_time, &instance_));
    instance_time_ = initial_time;
  }

  ~Ephemeris() {
    pb_nbody_ensemble_destroy(instance_);
    pb_ephemeris_destroy(ephemeris_);
  }
  Ephemeris(Ephemeris const&) = delete;
  Ephemeris& operator=(Ephemeris const&) = delete;

  pb_ephemeris_t* handle() const { return ephemeris_; }
  bool empty() const { return trajectories_.front().empty(); }
  double t_min() const { return pb_ephemeris_t_min(ephemeris_); }
  double t_max() const {
    double t = trajectories_.front().t_max();
    for (auto const& trajectory : trajectories_) {
      t = std::min(t, trajectory.t_max());
    }
    return t;
  }
  Status const& last_severe_integration_status() const {
    return last_severe_integration_status_;
  }
  // The segments of `body` (caller index), for inspection.
  std::vector<Segment> const& segments(int32_t const body) const {
    return trajectories_[InternalIndex(body)].segments();
  }

  // physics/ephemeris_body.hpp:312-344.  The massive bodies are integrated on
  // the GPU (pb_nbody_ensemble_*), every appended state goes through
  // ContinuousTrajectory::Append on the host (AppendMassiveBodiesState,
  // :1071-1116) and the new polynomials are handed to the GPU.
  Status Prolong(double const t,
                 std::int64_t const max_ephemeris_steps =
                     std::numeric_limits<std::int64_t>::max()) {
    double const step = fixed_step_parameters_.step;
    double const instance_time = instance_time_;
    double const desired_t_max = std::min(
        t, max_ephemeris_steps == std::numeric_limits<std::int64_t>::max()
               ? std::numeric_limits<double>::infinity()
               : instance_time + max_ephemeris_steps * step);
    double t_final;
    if (desired_t_max <= instance_time) {
      t_final = instance_time + step;
    } else {
      t_final = desired_t_max;
    }
    while (t_max() < desired_t_max) {
      Status const status = Solve(t_final);
      if (!status.ok()) {
        return status;
      }
      t_final += step;
    }
    return Status{};
  }

  // The IntrinsicAcceleration of the reference is a host function of t
  // (ephemeris.hpp:72-74); here it is tabulated: an inertially fixed burn,
  // Manœuvre::ComputeIntrinsicAcceleration (ksp_plugin/manœuvre_body.hpp:
  // 395-406), or none (NoIntrinsicAcceleration).
  struct IntrinsicAcceleration {
    double direction[3];
    double thrust;
    double initial_mass;
    double mass_flow;
    double initial_time;
    double final_time;
  };
  static constexpr IntrinsicAcceleration const* NoIntrinsicAcceleration =
      nullptr;

  using IntrinsicAccelerations = std::vector<IntrinsicAcceleration const*>;

  // physics/ephemeris_body.hpp:346-412.  All trajectories must end at the same
  // time (CHECK_EQ there).
  std::unique_ptr<Instance> NewInstance(
      std::vector<DiscreteTrajectory*> const& trajectories,
      FixedStepParameters const& parameters) {
    return NewInstance(trajectories, IntrinsicAccelerations{}, parameters);
  }

  // `intrinsic_accelerations` is empty (NoIntrinsicAccelerations) or has one
  // entry per trajectory, null for none.
  std::unique_ptr<Instance> NewInstance(
      std::vector<DiscreteTrajectory*> const& trajectories,
      IntrinsicAccelerations const& intrinsic_accelerations,
      FixedStepParameters const& parameters) {
    if (!intrinsic_accelerations.empty() &&
        intrinsic_accelerations.size() != trajectories.size()) {
      throw std::invalid_argument(
          "CHECK(intrinsic_accelerations.empty() || "
          "intrinsic_accelerations.size() == trajectories.size())");
    }
    if (trajectories.empty()) {
      throw std::invalid_argument("CHECK(!trajectories.empty())");
    }
    auto instance = std::make_unique<Instance>();
    instance->parameters_ = parameters;
    instance->trajectories_ = trajectories;
    std::size_t const n = trajectories.size();
    std::vector<double> state(12 * n, 0.0);
    double const last_time = trajectories.front()->back_time();
    for (std::size_t i = 0; i < n; ++i) {
      if (trajectories[i]->back_time() != last_time) {
        throw std::invalid_argument("CHECK_EQ(last_time, trajectory_last_time)");
      }
      DegreesOfFreedom const& dof = trajectories[i]->back();
      for (int c = 0; c < 3; ++c) {
        state[(0 + c) * n + i] = dof.q[c];
        state[(6 + c) * n + i] = dof.v[c];
      }
    }
    instance->time_[0] = last_time;
    pb_status_t const call = pb_fixed_step_instance_create(
        ephemeris_, parameters.integrator, parameters.step,
        static_cast<std::int64_t>(n), state.data(), instance->time_,
        &instance->handle_);
    if (call != PB_OK) {
      throw std::invalid_argument(pb_last_error());
    }
    if (!intrinsic_accelerations.empty()) {
      // A trajectory without intrinsic acceleration gets a burn that never
      // fires: the reference adds nothing for a null function, the table adds
      // the zero vector (the same bits unless an acceleration is -0.0).
      std::vector<double> burns(PB_BURN_PLANES * n, 0.0);
      for (std::size_t i = 0; i < n; ++i) {
        IntrinsicAcceleration const* const a = intrinsic_accelerations[i];
        if (a == nullptr) {
          burns[6 * n + i] = std::numeric_limits<double>::infinity();
          burns[7 * n + i] = -std::numeric_limits<double>::infinity();
          continue;
        }
        for (int c = 0; c < 3; ++c) {
          burns[c * n + i] = a->direction[c];
        }
        burns[3 * n + i] = a->thrust;
        burns[4 * n + i] = a->initial_mass;
        burns[5 * n + i] = a->mass_flow;
        burns[6 * n + i] = a->initial_time;
        burns[7 * n + i] = a->final_time;
      }
      if (pb_fixed_step_instance_set_intrinsic_accelerations(
              instance->handle_, burns.data()) != PB_OK) {
        throw std::invalid_argument(pb_last_error());
      }
    }
    Prolong(last_time + parameters.step);
    return instance;
  }

  // physics/ephemeris_body.hpp:480-493.  Every step is appended to the
  // trajectories (AppendMasslessBodiesStateToTrajectories, :1118-1131).
  Status FlowWithFixedStep(double const t, Instance& instance) {
    if (empty() || t > t_max()) {
      Status const status = Prolong(t);
      if (!status.ok()) {
        return status;
      }
    }
    if (instance.time_[0] == t && instance.time_[1] == 0.0) {
      return Status{};
    }
    std::int64_t const n = static_cast<std::int64_t>(instance.trajectories_.size());
    double const step = instance.parameters_.step;
    // Upper bound on the number of steps for the sample buffer.
    std::int64_t const max_samples = static_cast<std::int64_t>(
        std::abs((t - instance.time_[0]) / step)) + 2;
    std::vector<double> samples(static_cast<std::size_t>(max_samples) * 6 * n);
    std::vector<double> sample_times(max_samples);
    std::int64_t n_samples = 0, n_steps = 0;
    pb_status_t flow_status = PB_OK;
    pb_status_t call = pb_fixed_step_instance_flow(
        instance.handle_, t, /*sample_every=*/1, max_samples, samples.data(),
        sample_times.data(), &n_steps, &n_samples, &flow_status);
    if (call == PB_OK) {
      call = pb_fixed_step_instance_get_state(instance.handle_, nullptr,
                                              instance.time_);
    }
    if (call != PB_OK) {
      return Status::FromCall(call);
    }
    for (std::int64_t s = 0; s < n_samples; ++s) {
      for (std::int64_t i = 0; i < n; ++i) {
        DegreesOfFreedom dof;
        for (int c = 0; c < 3; ++c) {
          dof.q[c] = samples[(s * 6 + c) * n + i];
          dof.v[c] = samples[(s * 6 + 3 + c) * n + i];
        }
        instance.trajectories_[i]->Append(sample_times[s], dof);
      }
    }
    return flow_status == PB_OK ? Status{}
                                : Status{flow_status, "Collision detected"};
  }

  // physics/ephemeris_body.hpp:415-444,1624-1696.
  Status FlowWithAdaptiveStep(DiscreteTrajectory* const trajectory,
                              double const t,
                              AdaptiveStepParameters const& parameters,
                              std::int64_t const max_ephemeris_steps) {
    return FlowWithAdaptiveStep(trajectory, NoIntrinsicAcceleration, t,
                                parameters, max_ephemeris_steps);
  }

  Status FlowWithAdaptiveStep(
      DiscreteTrajectory* const trajectory,
      IntrinsicAcceleration const* const intrinsic_acceleration,
      double const t, AdaptiveStepParameters const& parameters,
      std::int64_t const max_ephemeris_steps) {
    double const last_time = trajectory->back_time();
    if (last_time == t) {
      return Status{};
    }
    Status const prolonged = Prolong(t, max_ephemeris_steps);
    if (!prolonged.ok()) {
      return prolonged;
    }
    // The number of accepted steps is not known in advance: the (deterministic)
    // flow is re-run with a larger log if the log overflows.
    int32_t status = PB_OK;
    std::int64_t accepted = 0;
    std::vector<double> times, states;
    for (std::int64_t capacity = 4096;; capacity *= 8) {
      double state[12] = {};
      double time[2] = {last_time, 0.0};
      for (int c = 0; c < 3; ++c) {
        state[c] = trajectory->back().q[c];
        state[6 + c] = trajectory->back().v[c];
      }
      times.assign(capacity, 0.0);
      states.assign(capacity * 6, 0.0);
      pb_adaptive_step_flow_t flow{};
      flow.method = parameters.integrator;
      flow.t_final = t;
      flow.length_integration_tolerance =
          parameters.length_integration_tolerance;
      flow.speed_integration_tolerance = parameters.speed_integration_tolerance;
      flow.max_steps = parameters.max_steps;
      flow.n = 1;
      flow.state = state;
      flow.time = time;
      flow.status = &status;
      flow.accepted_steps = &accepted;
      flow.step_log_capacity = capacity;
      flow.step_log = times.data();
      flow.state_log = states.data();
      double burn[PB_BURN_PLANES];
      if (intrinsic_acceleration != nullptr) {
        for (int c = 0; c < 3; ++c) {
          burn[c] = intrinsic_acceleration->direction[c];
        }
        burn[3] = intrinsic_acceleration->thrust;
        burn[4] = intrinsic_acceleration->initial_mass;
        burn[5] = intrinsic_acceleration->mass_flow;
        burn[6] = intrinsic_acceleration->initial_time;
        burn[7] = intrinsic_acceleration->final_time;
        flow.burns = burn;
      }
      pb_status_t const call = pb_flow_with_adaptive_step(ephemeris_, &flow);
      if (call != PB_OK) {
        return Status::FromCall(call);
      }
      if (accepted <= capacity) {
        break;
      }
    }
    // append_state for every accepted step.
    for (std::int64_t k = 0; k < accepted; ++k) {
      DegreesOfFreedom dof;
      for (int c = 0; c < 3; ++c) {
        dof.q[c] = states[k * 6 + c];
        dof.v[c] = states[k * 6 + 3 + c];
      }
      trajectory->Append(times[k], dof);
    }
    return status == PB_OK ? Status{} : Status{status, "adaptive flow"};
  }

  // physics/ephemeris_body.hpp:621-633.
  Status ComputeGravitationalAccelerationOnMasslessBody(
      Position const& position, double const t, double (&acceleration)[3]) {
    pb_status_t ignored;
    return Status::FromCall(
        pb_compute_gravitational_acceleration_on_massless_bodies(
            ephemeris_, t, 1, position, acceleration, &ignored));
  }

  // Batched form: q and a are 3 planes of n.
  Status ComputeGravitationalAccelerationOnMasslessBodies(
      double const t, std::int64_t const n, double const* const q,
      double* const a, pb_status_t* const collision) {
    return Status::FromCall(
        pb_compute_gravitational_acceleration_on_massless_bodies(
            ephemeris_, t, n, q, a, collision));
  }

  // physics/ephemeris_body.hpp:646-664: accelerations of all massive bodies at
  // t (positions evaluated from the trajectories); [3 * body], caller order.
  Status ComputeGravitationalAccelerationOnMassiveBodies(
      double const t, std::vector<double>& accelerations) {
    accelerations.assign(3 * static_cast<std::size_t>(n_bodies_), 0.0);
    return Status::FromCall(
        pb_compute_gravitational_acceleration_on_massive_bodies(
            ephemeris_, t, nullptr, accelerations.data()));
  }

  // physics/ephemeris_body.hpp:666-686: positions given by the caller.
  Status ComputeGravitationalAccelerationOnMassiveBodies(
      std::vector<double> const& positions, double const t,
      std::vector<double>& accelerations) {
    accelerations.assign(3 * static_cast<std::size_t>(n_bodies_), 0.0);
    return Status::FromCall(
        pb_compute_gravitational_acceleration_on_massive_bodies(
            ephemeris_, t, positions.data(), accelerations.data()));
  }

  // physics/ephemeris_body.hpp:688-696.
  Status ComputeGravitationalPotential(Position const& position, double const t,
                                       double& potential) {
    return Status::FromCall(pb_compute_gravitational_potential(
        ephemeris_, t, 1, position, &potential));
  }

  // physics/ephemeris_body.hpp:524-561 (point masses only, like the
  // reference).
  Status ComputeGravitationalJerkOnMasslessBody(
      DegreesOfFreedom const& degrees_of_freedom, double const t,
      double (&jerk)[3]) {
    return Status::FromCall(pb_compute_gravitational_jerk_on_massless_bodies(
        ephemeris_, t, 1, degrees_of_freedom.q, degrees_of_freedom.v, jerk));
  }

  // Batched form: q, v and jerk are 3 planes of n.
  Status ComputeGravitationalJerkOnMasslessBodies(
      double const t, std::int64_t const n, double const* const q,
      double const* const v, double* const jerk) {
    return Status::FromCall(pb_compute_gravitational_jerk_on_massless_bodies(
        ephemeris_, t, n, q, v, jerk));
  }

  // physics/ephemeris_body.hpp:563-583: the jerk on the massive `body` (caller
  // index) at t.
  Status ComputeGravitationalJerkOnMassiveBody(int32_t const body,
                                               double const t,
                                               double (&jerk)[3]) {
    std::vector<double> jerks(3 * static_cast<std::size_t>(n_bodies_));
    Status const status = Status::FromCall(
        pb_compute_gravitational_jerk_and_jacobian_on_massive_bodies(
            ephemeris_, t, nullptr, nullptr, jerks.data(), nullptr));
    for (int c = 0; c < 3; ++c) {
      jerk[c] = jerks[3 * body + c];
    }
    return status;
  }

  // physics/ephemeris_body.hpp:585-605: every body, on degrees of freedom
  // precomputed by the caller ([3 * body], caller order).
  Status ComputeGravitationalJerkOnMassiveBodies(
      std::vector<double> const& positions,
      std::vector<double> const& velocities, std::vector<double>& jerks) {
    jerks.assign(3 * static_cast<std::size_t>(n_bodies_), 0.0);
    return Status::FromCall(
        pb_compute_gravitational_jerk_and_jacobian_on_massive_bodies(
            ephemeris_, 0.0, positions.data(), velocities.data(), jerks.data(),
            nullptr));
  }

  // physics/ephemeris_body.hpp:495-522: row-major 3x3.
  Status ComputeJacobianOnMassiveBody(int32_t const body, double const t,
                                      double (&jacobian)[9]) {
    std::vector<double> jacobians(9 * static_cast<std::size_t>(n_bodies_));
    Status const status = Status::FromCall(
        pb_compute_gravitational_jerk_and_jacobian_on_massive_bodies(
            ephemeris_, t, nullptr, nullptr, nullptr, jacobians.data()));
    for (int k = 0; k < 9; ++k) {
      jacobian[k] = jacobians[9 * body + k];
    }
    return status;
  }

  // trajectory(body)->EvaluateVelocity(t) for every body; [3 * body].
  Status EvaluateAllVelocities(double const t, std::vector<double>& velocities) {
    velocities.assign(3 * static_cast<std::size_t>(n_bodies_), 0.0);
    return Status::FromCall(
        pb_ephemeris_evaluate_all_velocities(ephemeris_, t, velocities.data()));
  }

  // physics/ephemeris_body.hpp:226-236; [3 * body], caller order.
  Status EvaluateAllPositions(double const t, std::vector<double>& positions) {
    positions.assign(3 * static_cast<std::size_t>(n_bodies_), 0.0);
    return Status::FromCall(
        pb_ephemeris_evaluate_all_positions(ephemeris_, t, positions.data()));
  }

 private:
  static void Check(pb_status_t const status) {
    if (status != PB_OK) {
      throw std::runtime_error(std::string("principia_b200: ") +
                               pb_last_error());
    }
  }

  int32_t InternalIndex(int32_t const caller) const {
    for (int32_t b = 0; b < n_bodies_; ++b) {
      if (order_[b] == caller) {
        return b;
      }
    }
    throw std::out_of_range("body");
  }

  // instance_->Solve(t_final) with AppendMassiveBodiesState as the callback.
  Status Solve(double const t_final) {
    constexpr std::int64_t capacity = 64;
    std::size_t const plane = static_cast<std::size_t>(n_bodies_);
    std::vector<double> times(capacity);
    std::vector<double> states(capacity * 6 * plane);
    for (;;) {
      std::int64_t recorded = 0;
      pb_status_t const call = pb_nbody_ensemble_solve_recording(
          instance_, t_final, capacity, times.data(), states.data(), &recorded);
      if (call != PB_OK) {
        return Status::FromCall(call);
      }
      for (std::int64_t r = 0; r < recorded; ++r) {
        double const* const state = states.data() + r * 6 * plane;
        for (int32_t b = 0; b < n_bodies_; ++b) {
          DegreesOfFreedom dof;
          for (int c = 0; c < 3; ++c) {
            dof.q[c] = state[(0 + c) * plane + b];
            dof.v[c] = state[(3 + c) * plane + b];
          }
          pb_status_t const appended = trajectories_[b].Append(times[r], dof);
          if (appended != PB_OK) {
            last_severe_integration_status_ =
                Status{appended, "Error extending trajectory: apocalypse"};
          }
        }
        instance_time_ = times[r];
      }
      Status const uploaded = UploadNewSegments();
      if (!uploaded.ok()) {
        return uploaded;
      }
      if (recorded < capacity) {
        return Status{};
      }
    }
  }

  Status UploadNewSegments() {
    for (int32_t b = 0; b < n_bodies_; ++b) {
      auto const& segments = trajectories_[b].segments();
      std::size_t const first = uploaded_segments_[b];
      if (segments.size() == first) {
        continue;
      }
      std::size_t const count = segments.size() - first;
      std::vector<double> t_max(count), origin(count),
          coefficients(count * 54);
      std::vector<int32_t> degree(count);
      for (std::size_t s = 0; s < count; ++s) {
        Segment const& segment = segments[first + s];
        t_max[s] = segment.t_max;
        origin[s] = segment.origin;
        degree[s] = segment.degree;
        std::copy(&segment.coefficients[0][0], &segment.coefficients[0][0] + 54,
                  coefficients.begin() + 54 * s);
      }
      pb_status_t const call = pb_ephemeris_append_segments(
          ephemeris_, order_[b], static_cast<int32_t>(count),
          trajectories_[b].t_min(), t_max.data(), origin.data(), degree.data(),
          coefficients.data());
      if (call != PB_OK) {
        return Status::FromCall(call);
      }
      uploaded_segments_[b] = segments.size();
    }
    return Status{};
  }

  AccuracyParameters accuracy_parameters_;
  FixedStepParameters fixed_step_parameters_;
  int32_t n_bodies_;
  pb_ephemeris_t* ephemeris_ = nullptr;
  pb_nbody_ensemble_t* instance_ = nullptr;
  double instance_time_ = 0;
  std::vector<int32_t> order_;  // internal -> caller.
  std::vector<ContinuousTrajectory> trajectories_;  // Internal order.
  std::vector<std::size_t> uploaded_segments_;
  Status last_severe_integration_status_;
};

}  // namespace principia_b200
