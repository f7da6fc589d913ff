// Test shim: exposes the C++ facade (principia_b200/cpp/ephemeris.hpp) to
// pytest through a few extern "C" functions.  Links libprincipia_b200.so.
#include <cstring>
#include <memory>
#include <vector>

#include "../../principia_b200/cpp/ephemeris.hpp"

using namespace principia_b200;

extern "C" {

void* ft_create(int32_t n, pb_body_t const* bodies, double const* q,
                double const* v, double t0, double fitting_tolerance,
                double geopotential_tolerance, int32_t method, double step) {
  std::vector<pb_body_t> b(bodies, bodies + n);
  std::vector<DegreesOfFreedom> dofs(n);
  for (int i = 0; i < n; ++i) {
    for (int c = 0; c < 3; ++c) {
      dofs[i].q[c] = q[3 * i + c];
      dofs[i].v[c] = v[3 * i + c];
    }
  }
  try {
    return new Ephemeris(b, dofs, t0,
                         AccuracyParameters{fitting_tolerance,
                                            geopotential_tolerance},
                         FixedStepParameters{method, step});
  } catch (std::exception const&) {
    return nullptr;
  }
}

void ft_destroy(void* e) { delete static_cast<Ephemeris*>(e); }

int32_t ft_prolong(void* e, double t) {
  return static_cast<Ephemeris*>(e)->Prolong(t).code;
}

double ft_t_max(void* e) { return static_cast<Ephemeris*>(e)->t_max(); }
double ft_t_min(void* e) { return static_cast<Ephemeris*>(e)->t_min(); }
int32_t ft_empty(void* e) { return static_cast<Ephemeris*>(e)->empty() ? 1 : 0; }

int32_t ft_segment_count(void* e, int32_t body) {
  return static_cast<int32_t>(
      static_cast<Ephemeris*>(e)->segments(body).size());
}

void ft_get_segments(void* e, int32_t body, double* t_max, double* origin,
                     int32_t* degree, double* coefficients) {
  auto const& segments = static_cast<Ephemeris*>(e)->segments(body);
  for (std::size_t s = 0; s < segments.size(); ++s) {
    t_max[s] = segments[s].t_max;
    origin[s] = segments[s].origin;
    degree[s] = segments[s].degree;
    std::memcpy(coefficients + 54 * s, segments[s].coefficients,
                54 * sizeof(double));
  }
}

// NewInstance + FlowWithFixedStep on n probes starting at t0; returns the
// status code; points[i] = number of points of trajectory i; last = the last
// degrees of freedom [n][6]; last_time.
int32_t ft_flow_with_fixed_step(void* e, int32_t n, double const* q,
                                double const* v, double t0, int32_t method,
                                double step, double t, int32_t* points,
                                double* last, double* last_time) {
  auto* ephemeris = static_cast<Ephemeris*>(e);
  std::vector<DiscreteTrajectory> trajectories(n);
  std::vector<DiscreteTrajectory*> pointers;
  for (int i = 0; i < n; ++i) {
    DegreesOfFreedom dof;
    for (int c = 0; c < 3; ++c) {
      dof.q[c] = q[3 * i + c];
      dof.v[c] = v[3 * i + c];
    }
    trajectories[i].Append(t0, dof);
    pointers.push_back(&trajectories[i]);
  }
  auto instance =
      ephemeris->NewInstance(pointers, FixedStepParameters{method, step});
  Status const status = ephemeris->FlowWithFixedStep(t, *instance);
  for (int i = 0; i < n; ++i) {
    points[i] = static_cast<int32_t>(trajectories[i].times.size());
    std::memcpy(last + 6 * i, trajectories[i].back().q, 3 * sizeof(double));
    std::memcpy(last + 6 * i + 3, trajectories[i].back().v, 3 * sizeof(double));
  }
  *last_time = trajectories[0].back_time();
  return status.code;
}

// NewInstance, then FlowWithFixedStep(t_mid) and FlowWithFixedStep(t) on the
// same instance (the plugin's pattern for vessel histories).
int32_t ft_flow_with_fixed_step_twice(void* e, int32_t n, double const* q,
                                      double const* v, double t0,
                                      int32_t method, double step, double t_mid,
                                      double t, int32_t* points, double* last,
                                      double* last_time) {
  auto* ephemeris = static_cast<Ephemeris*>(e);
  std::vector<DiscreteTrajectory> trajectories(n);
  std::vector<DiscreteTrajectory*> pointers;
  for (int i = 0; i < n; ++i) {
    DegreesOfFreedom dof;
    for (int c = 0; c < 3; ++c) {
      dof.q[c] = q[3 * i + c];
      dof.v[c] = v[3 * i + c];
    }
    trajectories[i].Append(t0, dof);
    pointers.push_back(&trajectories[i]);
  }
  auto instance =
      ephemeris->NewInstance(pointers, FixedStepParameters{method, step});
  Status status = ephemeris->FlowWithFixedStep(t_mid, *instance);
  if (status.ok()) {
    status = ephemeris->FlowWithFixedStep(t, *instance);
  }
  for (int i = 0; i < n; ++i) {
    points[i] = static_cast<int32_t>(trajectories[i].times.size());
    std::memcpy(last + 6 * i, trajectories[i].back().q, 3 * sizeof(double));
    std::memcpy(last + 6 * i + 3, trajectories[i].back().v, 3 * sizeof(double));
  }
  *last_time = trajectories[0].back_time();
  return status.code;
}

// FlowWithAdaptiveStep on one probe; returns the status; the appended points
// (up to capacity) in times / dofs[k][6].
int32_t ft_flow_with_adaptive_step(void* e, double const* q, double const* v,
                                   double t0, double t, double length_tolerance,
                                   double speed_tolerance, int64_t max_steps,
                                   int64_t capacity, int64_t* n_points,
                                   double* times, double* dofs) {
  auto* ephemeris = static_cast<Ephemeris*>(e);
  DiscreteTrajectory trajectory;
  DegreesOfFreedom dof;
  for (int c = 0; c < 3; ++c) {
    dof.q[c] = q[c];
    dof.v[c] = v[c];
  }
  trajectory.Append(t0, dof);
  AdaptiveStepParameters parameters;
  parameters.max_steps = max_steps;
  parameters.length_integration_tolerance = length_tolerance;
  parameters.speed_integration_tolerance = speed_tolerance;
  Status const status = ephemeris->FlowWithAdaptiveStep(
      &trajectory, t, parameters, std::numeric_limits<std::int64_t>::max());
  *n_points = static_cast<int64_t>(trajectory.times.size()) - 1;
  for (int64_t k = 0; k < *n_points && k < capacity; ++k) {
    times[k] = trajectory.times[k + 1];
    std::memcpy(dofs + 6 * k, trajectory.degrees_of_freedom[k + 1].q,
                3 * sizeof(double));
    std::memcpy(dofs + 6 * k + 3, trajectory.degrees_of_freedom[k + 1].v,
                3 * sizeof(double));
  }
  return status.code;
}

int32_t ft_acceleration_on_massless_body(void* e, double const* position,
                                         double t, double* acceleration) {
  double p[3] = {position[0], position[1], position[2]};
  double a[3];
  Status const status =
      static_cast<Ephemeris*>(e)->ComputeGravitationalAccelerationOnMasslessBody(
          p, t, a);
  std::memcpy(acceleration, a, sizeof(a));
  return status.code;
}

int32_t ft_jerk_and_jacobian(void* e, double const* q, double const* v,
                             int32_t body, double t, double* massless_jerk,
                             double* massive_jerk, double* jacobian) {
  Ephemeris* const ephemeris = static_cast<Ephemeris*>(e);
  DegreesOfFreedom dof;
  for (int c = 0; c < 3; ++c) {
    dof.q[c] = q[c];
    dof.v[c] = v[c];
  }
  double j1[3], j2[3], jac[9];
  Status status = ephemeris->ComputeGravitationalJerkOnMasslessBody(dof, t, j1);
  if (status.code != 0) {
    return status.code;
  }
  status = ephemeris->ComputeGravitationalJerkOnMassiveBody(body, t, j2);
  if (status.code != 0) {
    return status.code;
  }
  status = ephemeris->ComputeJacobianOnMassiveBody(body, t, jac);
  std::memcpy(massless_jerk, j1, sizeof(j1));
  std::memcpy(massive_jerk, j2, sizeof(j2));
  std::memcpy(jacobian, jac, sizeof(jac));
  return status.code;
}

}  // extern "C"
