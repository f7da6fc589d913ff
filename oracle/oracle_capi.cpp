// ORACLE -- TEST INFRASTRUCTURE ONLY (see numerics.hpp).
//
// C entry points over the CPU restatement, shaped like include/principia_b200.h
// (prefix orc_ instead of pb_) so that the parity tests feed the same buffers
// to both.  Also hosts the known-answer tests of the reference's integrators.
#include <omp.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../include/principia_b200.h"
#include "discrete_trajectory.hpp"
#include "ephemeris.hpp"

using namespace orc;

struct orc_ephemeris {
  std::unique_ptr<Ephemeris> e;
  std::vector<int> internal_of_caller;
};

namespace {

Body MakeBody(pb_body_t const& b) {
  Body body;
  body.mu = b.gravitational_parameter;
  body.rotating = b.is_rotating != 0;
  body.min_radius = b.min_radius;
  body.mean_radius = b.mean_radius;
  body.reference_angle = b.reference_angle;
  body.reference_instant = b.reference_instant;
  body.angular_frequency = b.angular_frequency;
  body.right_ascension = b.right_ascension_of_pole;
  body.declination = b.declination_of_pole;
  body.oblate = b.is_oblate != 0;
  if (body.oblate) {
    body.degree = b.geopotential_degree;
    body.zonal = b.is_zonal != 0;
    body.reference_radius = b.reference_radius;
    int const size = (body.degree + 1) * (body.degree + 2) / 2;
    body.cos.assign(b.cos, b.cos + size);
    body.sin.assign(b.sin, b.sin + size);
  }
  return body;
}

State StateFromPlanes(double const* planes, std::int64_t n, std::int64_t first,
                      std::int64_t count, double time_value,
                      double time_error) {
  State s;
  s.time.value = time_value;
  s.time.error = time_error;
  s.positions.resize(3 * count);
  s.velocities.resize(3 * count);
  for (std::int64_t i = 0; i < count; ++i) {
    for (int c = 0; c < 3; ++c) {
      s.positions[3 * i + c].value = planes[(0 + c) * n + first + i];
      s.positions[3 * i + c].error = planes[(3 + c) * n + first + i];
      s.velocities[3 * i + c].value = planes[(6 + c) * n + first + i];
      s.velocities[3 * i + c].error = planes[(9 + c) * n + first + i];
    }
  }
  return s;
}

void PlanesFromState(State const& s, double* planes, std::int64_t n,
                     std::int64_t first, std::int64_t count) {
  for (std::int64_t i = 0; i < count; ++i) {
    for (int c = 0; c < 3; ++c) {
      planes[(0 + c) * n + first + i] = s.positions[3 * i + c].value;
      planes[(3 + c) * n + first + i] = s.positions[3 * i + c].error;
      planes[(6 + c) * n + first + i] = s.velocities[3 * i + c].value;
      planes[(9 + c) * n + first + i] = s.velocities[3 * i + c].error;
    }
  }
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- KATs -----

// integrators/symplectic_runge_kutta_nyström_integrator_test.cpp:589-648 and
// symmetric_linear_multistep_integrator_test.cpp:338-388: 1-D harmonic
// oscillator, q0 = 1 m, v0 = 0, unit stiffness and mass.
int32_t orc_kat_fixed_step_oscillator(int32_t method, double t_final,
                                      double step, double* q_error,
                                      double* v_error, std::int64_t* n_states,
                                      std::int64_t* evaluations) {
  std::int64_t evals = 0;
  RightHandSide rhs = [&evals](double, std::vector<double> const& q,
                               std::vector<double>& g) {
    g[0] = -q[0] * 1.0;
    ++evals;
    return static_cast<int32_t>(kOk);
  };
  State state;
  state.positions.resize(1);
  state.velocities.resize(1);
  state.positions[0].value = 1.0;
  double qe = 0, ve = 0;
  std::int64_t count = 0;
  AppendState append = [&](State const& s) {
    double const t = s.time.value - 0.0;
    double const q = s.positions[0].value;
    double const v = s.velocities[0].value;
    qe = std::max(qe, std::abs(1.0 * CrCos(1.0 * t) - q));
    ve = std::max(ve, std::abs(-1.0 * CrSin(1.0 * t) - v));
    ++count;
  };
  FixedStepInstance instance(method, rhs, state, append, step);
  int32_t const status = instance.Solve(t_final);
  *q_error = qe;
  *v_error = ve;
  *n_states = count;
  *evaluations = evals;
  return status;
}

// The same oscillator from any initial state, with the states appended
// (time, q, v values) returned: the shape of the Symplecticity, Convergence,
// TimeReversibility and Termination tests of
// symplectic_runge_kutta_nyström_integrator_test.cpp:349-587 and
// symmetric_linear_multistep_integrator_test.cpp:161-336.  Returns the status
// of Solve; at most `capacity` states are stored, *n_states counts them all.
int32_t orc_oscillator_solve(int32_t method, double t0, double q0, double q0_error,
                             double v0, double v0_error, double t_final,
                             double step, std::int64_t capacity, double* times,
                             double* q, double* v, double* final_state,
                             std::int64_t* n_states,
                             std::int64_t* evaluations) {
  std::int64_t evals = 0;
  RightHandSide rhs = [&evals](double, std::vector<double> const& q,
                               std::vector<double>& g) {
    g[0] = -q[0] * 1.0;
    ++evals;
    return static_cast<int32_t>(kOk);
  };
  State state;
  state.time = {t0, 0.0};
  state.positions.resize(1);
  state.velocities.resize(1);
  state.positions[0] = {q0, q0_error};
  state.velocities[0] = {v0, v0_error};
  std::int64_t count = 0;
  State last = state;
  AppendState append = [&](State const& s) {
    if (count < capacity) {
      times[count] = s.time.value;
      q[count] = s.positions[0].value;
      v[count] = s.velocities[0].value;
    }
    last = s;
    ++count;
  };
  FixedStepInstance instance(method, rhs, state, append, step);
  int32_t const status = instance.Solve(t_final);
  if (final_state != nullptr) {
    final_state[0] = last.time.value;
    final_state[1] = last.time.error;
    final_state[2] = last.positions[0].value;
    final_state[3] = last.positions[0].error;
    final_state[4] = last.velocities[0].value;
    final_state[5] = last.velocities[0].error;
  }
  *n_states = count;
  *evaluations = evals;
  return status;
}

// embedded_explicit_runge_kutta_nyström_integrator_test.cpp:75-182.  Runs
// forward from (q0, v0, t0) to t_final; outputs the final state so that the
// backward leg can be chained.
int32_t orc_kat_adaptive_oscillator_max_steps(
    double t0, double q0, double q0_error, double v0, double v0_error,
    double t_final, double length_tolerance, double speed_tolerance,
    std::int64_t max_steps, std::int64_t* n_states, std::int64_t* evaluations,
    std::int64_t* initial_rejections, std::int64_t* subsequent_rejections,
    double* final_state /* t, q, qe, v, ve */);

int32_t orc_kat_adaptive_oscillator(double t0, double q0, double q0_error,
                                    double v0, double v0_error, double t_final,
                                    double length_tolerance,
                                    double speed_tolerance,
                                    std::int64_t* n_states,
                                    std::int64_t* evaluations,
                                    std::int64_t* initial_rejections,
                                    std::int64_t* subsequent_rejections,
                                    double* final_state /* t, q, qe, v, ve */) {
  return orc_kat_adaptive_oscillator_max_steps(
      t0, q0, q0_error, v0, v0_error, t_final, length_tolerance,
      speed_tolerance, std::numeric_limits<std::int64_t>::max(), n_states,
      evaluations, initial_rejections, subsequent_rejections, final_state);
}

// The same with the max_steps of the integrator's parameters (the MaxSteps test,
// :184-270).
int32_t orc_kat_adaptive_oscillator_max_steps(
    double t0, double q0, double q0_error, double v0, double v0_error,
    double t_final, double length_tolerance, double speed_tolerance,
    std::int64_t max_steps, std::int64_t* n_states, std::int64_t* evaluations,
    std::int64_t* initial_rejections, std::int64_t* subsequent_rejections,
    double* final_state /* t, q, qe, v, ve */) {
  std::int64_t evals = 0;
  RightHandSide rhs = [&evals](double, std::vector<double> const& q,
                               std::vector<double>& g) {
    g[0] = -q[0] * 1.0;
    ++evals;
    return static_cast<int32_t>(kOk);
  };
  State state;
  state.time.value = t0;
  state.positions.resize(1);
  state.velocities.resize(1);
  state.positions[0] = {q0, q0_error};
  state.velocities[0] = {v0, v0_error};
  std::int64_t count = 0, initial = 0, subsequent = 0;
  bool first_step = true;
  State last = state;
  AppendState append = [&](State const& s) {
    last = s;
    ++count;
  };
  ToleranceToErrorRatio ratio = [&](double, State const&,
                                    StateError const& error) {
    double const r =
        std::min(length_tolerance / std::abs(error.position_error[0]),
                 speed_tolerance / std::abs(error.velocity_error[0]));
    bool const tolerable = r > 1.0;
    if (!tolerable) {
      if (first_step) {
        ++initial;
      } else {
        ++subsequent;
      }
    } else if (first_step) {
      first_step = false;
    }
    return r;
  };
  AdaptiveParameters parameters;
  parameters.first_step = t_final - t0;
  parameters.max_steps = max_steps;
  ErknInstance instance(DormandElMikkawyPrince1986RKN434FM(), rhs, state,
                        append, ratio, parameters);
  int32_t const status = instance.Solve(t_final);
  *n_states = count;
  *evaluations = evals;
  *initial_rejections = initial;
  *subsequent_rejections = subsequent;
  final_state[0] = last.time.value;
  final_state[1] = last.positions[0].value;
  final_state[2] = last.positions[0].error;
  final_state[3] = last.velocities[0].value;
  final_state[4] = last.velocities[0].error;
  return status;
}

// embedded_explicit_runge_kutta_nyström_integrator_test.cpp:338-439 (Restart):
// the oscillator over `duration` with last_step_is_exact = false, then the same
// instance continued to 2 * duration; and again in one call to Solve.
// sizes[0..1]: appended states after the first and the second Solve;
// last_steps[0..1]: the last time differences at those two points; returns 1
// if the one-call solution equals the restarted one state for state.
int32_t orc_kat_adaptive_oscillator_restart(double duration,
                                            double length_tolerance,
                                            double speed_tolerance,
                                            std::int64_t* sizes,
                                            double* last_steps) {
  auto const make_rhs = [] {
    return RightHandSide([](double, std::vector<double> const& q,
                            std::vector<double>& g) {
      g[0] = -q[0] * 1.0;
      return static_cast<int32_t>(kOk);
    });
  };
  ToleranceToErrorRatio const ratio = [&](double, State const&,
                                          StateError const& error) {
    return std::min(length_tolerance / std::abs(error.position_error[0]),
                    speed_tolerance / std::abs(error.velocity_error[0]));
  };
  State initial;
  initial.positions.resize(1);
  initial.velocities.resize(1);
  initial.positions[0].value = 1.0;
  AdaptiveParameters parameters;
  parameters.first_step = duration;
  parameters.last_step_is_exact = false;
  std::vector<State> solution1, solution2;
  {
    ErknInstance instance(
        DormandElMikkawyPrince1986RKN434FM(), make_rhs(), initial,
        [&](State const& s) { solution1.push_back(s); }, ratio, parameters);
    if (instance.Solve(0.0 + duration) != kOk) {
      return -1;
    }
    sizes[0] = static_cast<std::int64_t>(solution1.size());
    last_steps[0] = solution1[solution1.size() - 1].time.value -
                    solution1[solution1.size() - 2].time.value;
    if (instance.Solve(0.0 + 2.0 * duration) != kOk) {
      return -1;
    }
    sizes[1] = static_cast<std::int64_t>(solution1.size());
    last_steps[1] = solution1[solution1.size() - 1].time.value -
                    solution1[solution1.size() - 2].time.value;
  }
  {
    ErknInstance instance(
        DormandElMikkawyPrince1986RKN434FM(), make_rhs(), initial,
        [&](State const& s) { solution2.push_back(s); }, ratio, parameters);
    if (instance.Solve(0.0 + 2.0 * duration) != kOk) {
      return -1;
    }
  }
  if (solution1.size() != solution2.size()) {
    return 0;
  }
  for (std::size_t i = 0; i < solution1.size(); ++i) {
    State const& a = solution1[i];
    State const& b = solution2[i];
    if (a.time.value != b.time.value || a.time.error != b.time.error ||
        a.positions[0].value != b.positions[0].value ||
        a.positions[0].error != b.positions[0].error ||
        a.velocities[0].value != b.velocities[0].value ||
        a.velocities[0].error != b.velocities[0].error) {
      return 0;
    }
  }
  return 1;
}

// embedded_explicit_runge_kutta_nyström_integrator_test.cpp:271-336
// (Singularity): the ideal rocket x"(t) = m' I_sp / (m0 - t m') with unit
// constants, integrated across its singularity at t = 1 s.
int32_t orc_kat_adaptive_rocket(double length_tolerance, double speed_tolerance,
                                std::int64_t* n_states, double* final_time,
                                double* final_position) {
  double const mass_flow = 1, initial_mass = 1, specific_impulse = 1;
  double const t_initial = 0;
  double const t_final = t_initial + 2 * initial_mass / mass_flow;
  RightHandSide rhs = [&](double const t, std::vector<double> const&,
                          std::vector<double>& g) {
    double const mass = initial_mass - (t - t_initial) * mass_flow;
    g.back() = mass_flow * specific_impulse / mass;
    return static_cast<int32_t>(kOk);
  };
  ToleranceToErrorRatio const ratio = [&](double, State const&,
                                          StateError const& error) {
    return std::min(length_tolerance / std::abs(error.position_error[0]),
                    speed_tolerance / std::abs(error.velocity_error[0]));
  };
  State state;
  state.positions.resize(1);
  state.velocities.resize(1);
  std::int64_t count = 0;
  State last = state;
  AdaptiveParameters parameters;
  parameters.first_step = t_final - t_initial;
  ErknInstance instance(
      DormandElMikkawyPrince1986RKN434FM(), rhs, state,
      [&](State const& s) {
        last = s;
        ++count;
      },
      ratio, parameters);
  int32_t const status = instance.Solve(t_final);
  *n_states = count;
  *final_time = last.time.value;
  *final_position = last.positions.back().value;
  return status;
}

// numerics/hermite3_body.hpp as the trajectory sink uses it: the interpolant of
// (q1, v1) at t1 and (q2, v2) at t2 evaluated with its derivative at n times
// (AoS outputs), and its LInfinityL₂Norm (meaningful when q1 and v1 vanish:
// :154-194).
void orc_hermite3(double t1, double t2, double const* q1, double const* q2,
                  double const* v1, double const* v2, std::int64_t n,
                  double const* times, double* values, double* derivatives,
                  double* l_infinity_l2_norm) {
  Hermite3 const h = MakeHermite3(t1, t2, {q1[0], q1[1], q1[2]},
                                  {q2[0], q2[1], q2[2]}, {v1[0], v1[1], v1[2]},
                                  {v2[0], v2[1], v2[2]});
  for (std::int64_t i = 0; i < n; ++i) {
    Vec3 const q = Evaluate(h, times[i]);
    Vec3 const v = EvaluateDerivative(h, times[i]);
    values[3 * i] = q.x;
    values[3 * i + 1] = q.y;
    values[3 * i + 2] = q.z;
    derivatives[3 * i] = v.x;
    derivatives[3 * i + 1] = v.y;
    derivatives[3 * i + 2] = v.z;
  }
  if (l_infinity_l2_norm != nullptr) {
    *l_infinity_l2_norm = LInfinityL2Norm(h);
  }
}

// HarmonicDamping(s0).ComputeDampedRadialQuantities along x,
// physics/harmonic_damping_body.hpp:16-60.
void orc_harmonic_damping(double s0, double r, double R_over_r, double R_prime,
                          double* thresholds, double* sigma_R_over_r,
                          double* grad_sigma_R_x) {
  HarmonicDamping const damping(s0);
  thresholds[0] = damping.inner_threshold;
  thresholds[1] = damping.outer_threshold;
  Vec3 grad;
  damping.ComputeDampedRadialQuantities(r, r * r, {1, 0, 0}, R_over_r, R_prime,
                                        *sigma_R_over_r, grad);
  *grad_sigma_R_x = grad.x;
}

// numerics/double_precision_body.hpp: the primitives the steppers use, on
// {value, error} pairs.  op: 0 Increment(x, a[0]); 1 TwoSum(a[0], b[0]);
// 2 QuickTwoSum; 3 TwoDifference; 4 x + y; 5 x - y; 6 Scale(a[0], y).
void orc_double_precision(int32_t op, double const* a, double const* b,
                          double* out) {
  DP x{a[0], a[1]};
  DP const y{b[0], b[1]};
  DP result{0.0, 0.0};
  switch (op) {
    case 0:
      Increment(x, b[0]);
      result = x;
      break;
    case 1:
      result = TwoSum(a[0], b[0]);
      break;
    case 2:
      result = QuickTwoSum(a[0], b[0]);
      break;
    case 3:
      result = TwoDifference(a[0], b[0]);
      break;
    case 4:
      result = x + y;
      break;
    case 5:
      result = x - y;
      break;
    case 6:
      result = Scale(a[0], y);
      break;
  }
  out[0] = result.value;
  out[1] = result.error;
}

// embedded_explicit_generalized_runge_kutta_nyström_integrator_test.cpp:71-154
// (Legendre): Fine1987RKNG34 on the Legendre differential equation of degree
// 15, p" = (2 x p' - n (n + 1) p) / (1 - x²) (testing_utilities/
// integration_body.hpp:95-110), a right-hand side that depends on the velocity.
int32_t orc_kat_legendre(double tolerance, double derivative_tolerance,
                         std::int64_t capacity, double* times, double* values,
                         double* derivatives, std::int64_t* n_states,
                         std::int64_t* evaluations,
                         std::int64_t* initial_rejections,
                         std::int64_t* subsequent_rejections) {
  constexpr int n = 15;
  std::int64_t evals = 0;
  GeneralizedRightHandSide rhs = [&evals](double const t,
                                          std::vector<double> const& p,
                                          std::vector<double> const& pv,
                                          std::vector<double>& pa) {
    double const x = t - 0.0;
    double const x2 = x * x;
    pa[0] = (2 * x * pv[0] - n * (n + 1) * p[0]) / (1 * 1.0 - x2);
    ++evals;
    return static_cast<int32_t>(kOk);
  };
  std::int64_t initial = 0, subsequent = 0;
  bool first_step = true;
  ToleranceToErrorRatio const ratio = [&](double, State const&,
                                          StateError const& error) {
    double const r =
        std::min(tolerance / std::abs(error.position_error[0]),
                 derivative_tolerance / std::abs(error.velocity_error[0]));
    bool const tolerable = r > 1.0;
    if (!tolerable) {
      if (first_step) {
        ++initial;
      } else {
        ++subsequent;
      }
    } else if (first_step) {
      first_step = false;
    }
    return r;
  };
  State state;
  state.positions.resize(1);
  state.velocities.resize(1);
  state.positions[0].value = 0;
  state.velocities[0].value = -6435 / (2048 * 1.0);
  double const t_final = 0.0 + 0.99 * 1.0;
  std::int64_t count = 0;
  AdaptiveParameters parameters;
  parameters.first_step = t_final - 0.0;
  ErknInstance instance(
      Fine1987RKNG34(), RightHandSide(), state,
      [&](State const& s) {
        if (count < capacity) {
          times[count] = s.time.value;
          values[count] = s.positions[0].value;
          derivatives[count] = s.velocities[0].value;
        }
        ++count;
      },
      ratio, parameters);
  instance.generalized_rhs = rhs;
  int32_t const status = instance.Solve(t_final);
  *n_states = count;
  *evaluations = evals;
  *initial_rejections = initial;
  *subsequent_rejections = subsequent;
  return status;
}

double orc_cr_sin(double x) { return CrSin(x); }
double orc_cr_cos(double x) { return CrCos(x); }
double orc_root(int32_t n, double x) { return Root(n, x); }

// Estrin evaluation of one polynomial (numerics/polynomial_evaluators_body.hpp).
double orc_estrin(double const* coefficients, int32_t degree, double x) {
  return EstrinEvaluate(coefficients, 1, 0, degree, x);
}

double orc_estrin_derivative(double const* coefficients, int32_t degree,
                             double x) {
  return EstrinEvaluateDerivative(coefficients, 1, 1, degree - 1, x);
}

// Чебышёв series (numerics/polynomial_in_чебышёв_basis_body.hpp:151-202).
double orc_chebyshev_evaluate(double const* coefficients, int32_t degree,
                              double lower_bound, double upper_bound,
                              double argument) {
  return ЧебышёвEvaluate(coefficients, 1, degree, lower_bound, upper_bound,
                         argument);
}
double orc_chebyshev_evaluate_derivative(double const* coefficients,
                                         int32_t degree, double lower_bound,
                                         double upper_bound, double argument) {
  return ЧебышёвEvaluateDerivative(coefficients, 1, degree, lower_bound,
                                   upper_bound, argument);
}

// DiscreteTrajectorySegment::Append with SetDownsampling({tolerance})
// (physics/discrete_trajectory_segment_body.hpp:414-479) on one trajectory of
// n_points (q, v: [n_points][3]).  Returns the size of the timeline; the
// retained points go to kept_* (room for n_points), with the downsampling
// error recorded with each interpolation.  If n_evaluations > 0 the segment is
// then evaluated (EvaluatePosition / EvaluateVelocity, :132-155).
int64_t orc_downsample(int64_t n_points, double const* times, double const* q,
                       double const* v, double tolerance, double* kept_times,
                       double* kept_q, double* kept_v, double* kept_errors,
                       int64_t n_evaluations, double const* evaluation_times,
                       double* evaluated_q, double* evaluated_v) {
  DownsamplingSegment segment(tolerance);
  for (int64_t i = 0; i < n_points; ++i) {
    segment.Append(times[i], Vec3{q[3 * i], q[3 * i + 1], q[3 * i + 2]},
                   Vec3{v[3 * i], v[3 * i + 1], v[3 * i + 2]});
  }
  auto const& timeline = segment.timeline();
  for (std::size_t k = 0; k < timeline.size(); ++k) {
    kept_times[k] = timeline[k].time;
    kept_q[3 * k] = timeline[k].q.x;
    kept_q[3 * k + 1] = timeline[k].q.y;
    kept_q[3 * k + 2] = timeline[k].q.z;
    kept_v[3 * k] = timeline[k].v.x;
    kept_v[3 * k + 1] = timeline[k].v.y;
    kept_v[3 * k + 2] = timeline[k].v.z;
    if (kept_errors != nullptr) {
      kept_errors[k] = timeline[k].error;
    }
  }
  for (int64_t j = 0; j < n_evaluations; ++j) {
    Vec3 eq{NAN, NAN, NAN}, ev{NAN, NAN, NAN};
    segment.Evaluate(evaluation_times[j], eq, ev);
    evaluated_q[3 * j] = eq.x;
    evaluated_q[3 * j + 1] = eq.y;
    evaluated_q[3 * j + 2] = eq.z;
    evaluated_v[3 * j] = ev.x;
    evaluated_v[3 * j + 1] = ev.y;
    evaluated_v[3 * j + 2] = ev.z;
  }
  return static_cast<int64_t>(timeline.size());
}

void orc_cr_sin_cos(double x, double* s, double* c) {
  *s = CrSin(x);
  *c = CrCos(x);
}

// Newhall fit of 9 equally spaced (q, v) pairs: returns the monomial
// coefficients (numerics/newhall_body.hpp:259-287) through a one-shot
// ContinuousTrajectory.
int32_t orc_newhall_fit(double const* times, double const* q, double const* v,
                        double step, double tolerance, double* t_max,
                        double* origin, int32_t* degree,
                        double* coefficients) {
  ContinuousTrajectory trajectory(step, tolerance);
  int32_t status = kOk;
  for (int i = 0; i < 9; ++i) {
    status = trajectory.Append(times[i], {q[3 * i], q[3 * i + 1], q[3 * i + 2]},
                               {v[3 * i], v[3 * i + 1], v[3 * i + 2]});
  }
  Segment const& s = trajectory.segments().back();
  *t_max = s.t_max;
  *origin = s.origin;
  *degree = s.degree;
  std::memcpy(coefficients, s.c, sizeof(s.c));
  return status;
}

// Newhall fit of a given degree to 9 equally spaced samples of one scalar
// function and its derivative, evaluated (value and derivative, Estrin) at n
// times: the shape of NewhallTest::CheckNewhallApproximationErrors with the
// MonomialAdapter, numerics/newhall_test.cpp:181-230.
int32_t orc_newhall_fixed_degree(int32_t degree, double const* q,
                                 double const* v, double t_min, double t_max,
                                 double* error_estimate, int64_t n,
                                 double const* times, double* values,
                                 double* derivatives) {
  ContinuousTrajectory trajectory((t_max - t_min) / 8, 0.0);
  std::vector<std::pair<Vec3, Vec3>> qvs;
  for (int i = 0; i < 9; ++i) {
    qvs.push_back({Vec3{q[i], 0, 0}, Vec3{v[i], 0, 0}});
  }
  Vec3 estimate;
  Segment const s = trajectory.Newhall(degree, qvs, t_min, t_max, estimate);
  *error_estimate = estimate.x;
  for (int64_t i = 0; i < n; ++i) {
    double const x = times[i] - s.origin;
    values[i] = EstrinEvaluate(&s.c[0][0], 3, 0, s.degree, x);
    derivatives[i] = EstrinEvaluateDerivative(&s.c[0][0], 3, 1, s.degree - 1, x);
  }
  return kOk;
}

// ContinuousTrajectoryTest.BestNewhallApproximation,
// physics/continuous_trajectory_test.cpp:234-441: ComputeBestNewhallApproximation
// driven by scripted error estimates.  calls[c] fits are consumed by call c,
// in order: the degree each fit must be asked for and the estimate it yields.
// After each call: degree_, adjusted_tolerance_, is_unstable_; `reset_after[c]`
// applies ResetBestNewhallApproximation.  Returns 0, or 1 + the index of the
// first fit requested out of script.
int32_t orc_kat_best_newhall(double step, double tolerance, int32_t n_calls,
                             int32_t const* fits_per_call,
                             int32_t const* scripted_degrees,
                             double const* scripted_estimates,
                             int32_t const* reset_after, int32_t* degrees,
                             double* adjusted_tolerances,
                             int32_t* is_unstable, int32_t* statuses) {
  ContinuousTrajectory trajectory(step, tolerance);
  trajectory.Append(0.0, Vec3{}, Vec3{});
  int32_t next = 0;
  int32_t limit = 0;
  int32_t mismatch = 0;
  trajectory.mock_error_estimate = [&](int const degree) {
    if (mismatch == 0 && (next >= limit || scripted_degrees[next] != degree)) {
      mismatch = 1 + next;
    }
    Vec3 const estimate{scripted_estimates[3 * next],
                        scripted_estimates[3 * next + 1],
                        scripted_estimates[3 * next + 2]};
    ++next;
    return estimate;
  };
  double t = 0.0;
  for (int32_t c = 0; c < n_calls; ++c) {
    limit = next + fits_per_call[c];
    t += step;
    statuses[c] = trajectory.ComputeBestNewhallApproximation(t, {});
    if (mismatch == 0 && next != limit) {
      mismatch = 1 + next;
    }
    degrees[c] = trajectory.degree();
    adjusted_tolerances[c] = trajectory.adjusted_tolerance();
    is_unstable[c] = trajectory.is_unstable() ? 1 : 0;
    if (reset_after[c] != 0) {
      trajectory.ResetBestNewhallApproximation();
    }
  }
  return mismatch;
}

// A ContinuousTrajectory of its own (physics/continuous_trajectory_body.hpp):
// Append, t_min / t_max, EvaluatePosition / EvaluateVelocity.
struct orc_continuous_trajectory {
  ContinuousTrajectory trajectory;
};
orc_continuous_trajectory* orc_continuous_trajectory_create(double step,
                                                            double tolerance) {
  return new orc_continuous_trajectory{ContinuousTrajectory(step, tolerance)};
}
void orc_continuous_trajectory_destroy(orc_continuous_trajectory* t) {
  delete t;
}
int32_t orc_continuous_trajectory_append(orc_continuous_trajectory* t,
                                         double time, double const* q,
                                         double const* v) {
  return t->trajectory.Append(time, {q[0], q[1], q[2]}, {v[0], v[1], v[2]});
}
double orc_continuous_trajectory_t_min(orc_continuous_trajectory const* t) {
  return t->trajectory.t_min();
}
double orc_continuous_trajectory_t_max(orc_continuous_trajectory const* t) {
  return t->trajectory.t_max();
}
int32_t orc_continuous_trajectory_segment_count(
    orc_continuous_trajectory const* t) {
  return static_cast<int32_t>(t->trajectory.segments().size());
}
int32_t orc_continuous_trajectory_degree(orc_continuous_trajectory const* t,
                                         int32_t segment) {
  return t->trajectory.segments()[segment].degree;
}
void orc_continuous_trajectory_get_segment(orc_continuous_trajectory const* t,
                                           int32_t segment, double* t_max,
                                           double* origin, int32_t* degree,
                                           double* coefficients) {
  Segment const& s = t->trajectory.segments()[segment];
  *t_max = s.t_max;
  *origin = s.origin;
  *degree = s.degree;
  std::memcpy(coefficients, s.c, sizeof(s.c));
}
void orc_continuous_trajectory_evaluate(orc_continuous_trajectory const* t,
                                        double time, double* q, double* v) {
  Vec3 const position = t->trajectory.EvaluatePosition(time);
  Vec3 const velocity = t->trajectory.EvaluateVelocity(time);
  q[0] = position.x;
  q[1] = position.y;
  q[2] = position.z;
  v[0] = velocity.x;
  v[1] = velocity.y;
  v[2] = velocity.z;
}

// ------------------------------------------------------------ Ephemeris ----

int32_t orc_ephemeris_create(int32_t n_bodies, pb_body_t const* bodies,
                             double const* q, double const* v,
                             double initial_time, double fitting_tolerance,
                             double geopotential_tolerance, int32_t method,
                             double step, orc_ephemeris** out) {
  std::vector<Body> bs;
  std::vector<Vec3> qs, vs;
  for (int i = 0; i < n_bodies; ++i) {
    bs.push_back(MakeBody(bodies[i]));
    qs.push_back({q[3 * i], q[3 * i + 1], q[3 * i + 2]});
    vs.push_back({v[3 * i], v[3 * i + 1], v[3 * i + 2]});
  }
  auto* result = new orc_ephemeris;
  result->e = std::make_unique<Ephemeris>(bs, qs, vs, initial_time,
                                          fitting_tolerance,
                                          geopotential_tolerance, method, step);
  result->internal_of_caller.resize(n_bodies);
  for (int i = 0; i < n_bodies; ++i) {
    result->internal_of_caller[result->e->order()[i]] = i;
  }
  *out = result;
  return kOk;
}

void orc_ephemeris_destroy(orc_ephemeris* e) { delete e; }

int32_t orc_ephemeris_internal_order(orc_ephemeris const* e, int32_t* order,
                                     int32_t* n_oblate) {
  for (int i = 0; i < e->e->size(); ++i) {
    order[i] = e->e->order()[i];
  }
  if (n_oblate != nullptr) {
    *n_oblate = e->e->number_of_oblate_bodies();
  }
  return kOk;
}

int32_t orc_geopotential_damping(orc_ephemeris const* e, int32_t body,
                                 int32_t* n_degrees, double* inner_thresholds,
                                 double* sectoral_inner_threshold) {
  int const internal = e->internal_of_caller[body];
  if (internal >= e->e->number_of_oblate_bodies()) {
    return kInvalidArgument;
  }
  Geopotential const& g = e->e->geopotential(internal);
  *n_degrees = static_cast<int32_t>(g.degree_damping().size());
  for (std::size_t i = 0; i < g.degree_damping().size(); ++i) {
    inner_thresholds[i] = g.degree_damping()[i].inner_threshold;
  }
  *sectoral_inner_threshold = g.sectoral_damping().inner_threshold;
  return kOk;
}

// Geopotential::GeneralSphericalHarmonicsAcceleration on one displacement.
// Geopotential::GeneralSphericalHarmonicsAcceleration with every argument
// given by the caller (physics/geopotential_test.cpp:651-656 passes r, r * r
// and 1 / Pow<3>(r) for the vector r * direction).
int32_t orc_geopotential_acceleration_explicit(orc_ephemeris const* e,
                                               int32_t body, double t,
                                               double const* r, double r_norm,
                                               double r2, double one_over_r3,
                                               double* out) {
  int const internal = e->internal_of_caller[body];
  Vec3 const a =
      e->e->geopotential(internal).GeneralSphericalHarmonicsAcceleration(
          t, {r[0], r[1], r[2]}, r_norm, r2, one_over_r3);
  out[0] = a.x;
  out[1] = a.y;
  out[2] = a.z;
  return kOk;
}

// Geopotential::GeneralSphericalHarmonicsPotential with the arguments the
// reference's tests form (physics/geopotential_test.cpp:109-119).
double orc_geopotential_potential(orc_ephemeris const* e, int32_t body, double t,
                                  double const* r) {
  int const internal = e->internal_of_caller[body];
  Vec3 const rv{r[0], r[1], r[2]};
  double const r2 = Norm2(rv);
  double const r_norm = std::sqrt(r2);
  double const one_over_r3 = r_norm / (r2 * r2);
  return e->e->geopotential(internal).GeneralSphericalHarmonicsPotential(
      t, rv, r_norm, r2, one_over_r3);
}

// RotatingBody::FromSurfaceFrame(t) (inverse == 0) or ToSurfaceFrame(t)
// (inverse != 0: the conjugate quaternion, geometry/rotation_body.hpp:198-201)
// applied to a vector, physics/rotating_body.hpp:181-198.
int32_t orc_surface_frame_rotate(orc_ephemeris const* e, int32_t body, double t,
                                 int32_t inverse, double const* in,
                                 double* out) {
  int const internal = e->internal_of_caller[body];
  Quaternion q = e->e->body(internal).FromSurfaceFrame(t);
  if (inverse != 0) {
    q.imaginary = -q.imaginary;
  }
  Vec3 const v = Rotate(q, {in[0], in[1], in[2]});
  out[0] = v.x;
  out[1] = v.y;
  out[2] = v.z;
  return kOk;
}

int32_t orc_geopotential_acceleration(orc_ephemeris const* e, int32_t body,
                                      double t, double const* r, double* out) {
  int const internal = e->internal_of_caller[body];
  Vec3 const rv{r[0], r[1], r[2]};
  double const r2 = Norm2(rv);
  double const r_norm = std::sqrt(r2);
  double const one_over_r3 = r_norm / (r2 * r2);
  Vec3 const a = e->e->geopotential(internal).GeneralSphericalHarmonicsAcceleration(
      t, rv, r_norm, r2, one_over_r3);
  out[0] = a.x;
  out[1] = a.y;
  out[2] = a.z;
  return kOk;
}

int32_t orc_ephemeris_prolong(orc_ephemeris* e, double t) {
  e->e->Prolong(t);
  return e->e->last_severe_integration_status();
}

// The state of the massive-body instance (internal order), for KATs.
int32_t orc_ephemeris_instance_state(orc_ephemeris const* e, double* time,
                                     double* q, double* v) {
  State const& s = e->e->instance().state();
  time[0] = s.time.value;
  time[1] = s.time.error;
  for (std::size_t i = 0; i < s.positions.size(); ++i) {
    q[i] = s.positions[i].value;
    v[i] = s.velocities[i].value;
  }
  return kOk;
}

int32_t orc_ephemeris_segment_count(orc_ephemeris* e, int32_t body) {
  return static_cast<int32_t>(
      e->e->trajectory(e->internal_of_caller[body]).segments().size());
}

int32_t orc_ephemeris_get_segments(orc_ephemeris* e, int32_t body,
                                   double* t_min, double* t_max, double* origin,
                                   int32_t* degree, double* coefficients) {
  ContinuousTrajectory& trajectory =
      e->e->trajectory(e->internal_of_caller[body]);
  *t_min = trajectory.t_min();
  auto const& segments = trajectory.segments();
  for (std::size_t s = 0; s < segments.size(); ++s) {
    t_max[s] = segments[s].t_max;
    origin[s] = segments[s].origin;
    degree[s] = segments[s].degree;
    std::memcpy(coefficients + 54 * s, segments[s].c, sizeof(segments[s].c));
  }
  return kOk;
}

int32_t orc_ephemeris_append_segments(orc_ephemeris* e, int32_t body,
                                      int32_t n_segments, double t_min,
                                      double const* t_max, double const* origin,
                                      int32_t const* degree,
                                      double const* coefficients) {
  ContinuousTrajectory& trajectory =
      e->e->trajectory(e->internal_of_caller[body]);
  for (int s = 0; s < n_segments; ++s) {
    Segment segment;
    segment.t_max = t_max[s];
    segment.origin = origin[s];
    segment.degree = degree[s];
    std::memcpy(segment.c, coefficients + 54 * s, sizeof(segment.c));
    trajectory.AppendSegment(segment, t_min);
  }
  return kOk;
}

double orc_ephemeris_t_min(orc_ephemeris const* e) { return e->e->t_min(); }
double orc_ephemeris_t_max(orc_ephemeris const* e) { return e->e->t_max(); }

int32_t orc_ephemeris_evaluate_all_positions(orc_ephemeris* e, double t,
                                             double* positions) {
  auto const q = e->e->EvaluateAllPositions(t);
  for (int i = 0; i < e->e->size(); ++i) {
    int const caller = e->e->order()[i];
    positions[3 * caller] = q[i].x;
    positions[3 * caller + 1] = q[i].y;
    positions[3 * caller + 2] = q[i].z;
  }
  return kOk;
}

// ContinuousTrajectory::EvaluateVelocity of every body, caller order.
int32_t orc_ephemeris_evaluate_all_velocities(orc_ephemeris* e, double t,
                                              double* velocities) {
  auto const v = e->e->EvaluateAllVelocities(t);
  for (int i = 0; i < e->e->size(); ++i) {
    int const caller = e->e->order()[i];
    velocities[3 * caller] = v[i].x;
    velocities[3 * caller + 1] = v[i].y;
    velocities[3 * caller + 2] = v[i].z;
  }
  return kOk;
}

int32_t orc_compute_gravitational_acceleration_on_massless_bodies(
    orc_ephemeris* e, double t, std::int64_t n, double const* q, double* a,
    int32_t* status) {
  int32_t error = kOk;
#pragma omp parallel for reduction(| : error) schedule(static)
  for (std::int64_t chunk = 0; chunk < (n + 255) / 256; ++chunk) {
    std::int64_t const first = chunk * 256;
    std::int64_t const count = std::min<std::int64_t>(256, n - first);
    std::vector<Vec3> positions(count), accelerations(count);
    for (std::int64_t i = 0; i < count; ++i) {
      positions[i] = {q[first + i], q[n + first + i], q[2 * n + first + i]};
    }
    error |=
        e->e->ComputeGravitationalAccelerationByAllMassiveBodiesOnMasslessBodies(
            t, positions, accelerations);
    for (std::int64_t i = 0; i < count; ++i) {
      a[first + i] = accelerations[i].x;
      a[n + first + i] = accelerations[i].y;
      a[2 * n + first + i] = accelerations[i].z;
    }
  }
  if (status != nullptr) {
    *status = error;
  }
  return kOk;
}

int32_t orc_compute_gravitational_potential(orc_ephemeris* e, double t,
                                            std::int64_t n, double const* q,
                                            double* potential) {
  std::vector<Vec3> positions(n);
  std::vector<double> potentials(n);
  for (std::int64_t i = 0; i < n; ++i) {
    positions[i] = {q[i], q[n + i], q[2 * n + i]};
  }
  e->e->ComputeGravitationalPotentialsOfAllMassiveBodies(t, positions,
                                                         potentials);
  std::memcpy(potential, potentials.data(), n * sizeof(double));
  return kOk;
}

int32_t orc_compute_gravitational_acceleration_on_massive_bodies(
    orc_ephemeris* e, double t, double const* positions,
    double* accelerations) {
  int const n = e->e->size();
  std::vector<Vec3> q(n);
  if (positions == nullptr) {
    q = e->e->EvaluateAllPositions(t);
  } else {
    for (int i = 0; i < n; ++i) {
      int const caller = e->e->order()[i];
      q[i] = {positions[3 * caller], positions[3 * caller + 1],
              positions[3 * caller + 2]};
    }
  }
  for (int i = 0; i < n; ++i) {
    Vec3 const a = e->e->ComputeGravitationalAccelerationOnMassiveBody(i, q, t);
    int const caller = e->e->order()[i];
    accelerations[3 * caller] = a.x;
    accelerations[3 * caller + 1] = a.y;
    accelerations[3 * caller + 2] = a.z;
  }
  return kOk;
}

// Ephemeris::ComputeGravitationalJerkOnMasslessBody (batched),
// ephemeris_body.hpp:524-561.  SoA q, v, jerk.
int32_t orc_compute_gravitational_jerk_on_massless_bodies(
    orc_ephemeris* e, double t, std::int64_t n, double const* q,
    double const* v, double* jerk) {
#pragma omp parallel for schedule(static)
  for (std::int64_t i = 0; i < n; ++i) {
    Vec3 const j = e->e->ComputeGravitationalJerkOnMasslessBody(
        {q[i], q[n + i], q[2 * n + i]}, {v[i], v[n + i], v[2 * n + i]}, t);
    jerk[i] = j.x;
    jerk[n + i] = j.y;
    jerk[2 * n + i] = j.z;
  }
  return kOk;
}

namespace {
void DegreesOfFreedomInInternalOrder(orc_ephemeris* e, double t,
                                     double const* positions,
                                     double const* velocities,
                                     std::vector<Vec3>& q,
                                     std::vector<Vec3>& v) {
  int const n = e->e->size();
  q.resize(n);
  v.resize(n);
  if (positions == nullptr) {
    q = e->e->EvaluateAllPositions(t);
    v = e->e->EvaluateAllVelocities(t);
  } else {
    for (int i = 0; i < n; ++i) {
      int const caller = e->e->order()[i];
      q[i] = {positions[3 * caller], positions[3 * caller + 1],
              positions[3 * caller + 2]};
      if (velocities != nullptr) {
        v[i] = {velocities[3 * caller], velocities[3 * caller + 1],
                velocities[3 * caller + 2]};
      }
    }
  }
}
}  // namespace

// Ephemeris::ComputeGravitationalJerkOnMassiveBody(body, t) (positions ==
// nullptr) / ...OnMassiveBodies(bodies, degrees of freedom),
// ephemeris_body.hpp:563-605: jerks[3 * body], caller order.
int32_t orc_compute_gravitational_jerk_on_massive_bodies(
    orc_ephemeris* e, double t, double const* positions,
    double const* velocities, double* jerks) {
  std::vector<Vec3> q, v;
  DegreesOfFreedomInInternalOrder(e, t, positions, velocities, q, v);
  for (int i = 0; i < e->e->size(); ++i) {
    Vec3 const j = e->e->ComputeGravitationalJerkOnMassiveBody(i, q, v);
    int const caller = e->e->order()[i];
    jerks[3 * caller] = j.x;
    jerks[3 * caller + 1] = j.y;
    jerks[3 * caller + 2] = j.z;
  }
  return kOk;
}

// Ephemeris::ComputeJacobianOnMassiveBody, ephemeris_body.hpp:495-522:
// jacobians[9 * body], row-major, caller order.
int32_t orc_compute_jacobian_on_massive_bodies(orc_ephemeris* e, double t,
                                               double const* positions,
                                               double* jacobians) {
  std::vector<Vec3> q, v;
  DegreesOfFreedomInInternalOrder(e, t, positions, nullptr, q, v);
  for (int i = 0; i < e->e->size(); ++i) {
    auto const form = e->e->ComputeJacobianOnMassiveBody(i, q);
    int const caller = e->e->order()[i];
    std::memcpy(jacobians + 9 * caller, form.m, 9 * sizeof(double));
  }
  return kOk;
}

// The RHS of the massive-body equation on a full position set (internal
// order, AoS), ephemeris_body.hpp:1503-1552.
int32_t orc_compute_gravitational_acceleration_between_all_massive_bodies(
    orc_ephemeris* e, double t, double const* positions,
    double* accelerations) {
  int const n = e->e->size();
  std::vector<double> q(positions, positions + 3 * n), a(3 * n);
  e->e->ComputeGravitationalAccelerationBetweenAllMassiveBodies(t, q, a);
  std::memcpy(accelerations, a.data(), 3 * n * sizeof(double));
  return kOk;
}

// FlowWithFixedStep on n probes.  The probes are cut into `n_threads` chunks,
// one integrator instance per chunk, run on an OpenMP team -- the way the
// reference parallelises independent trajectories
// (benchmarks/ephemeris_benchmark.cpp:354-386).  The result does not depend on
// the partition.
int32_t orc_flow_with_fixed_step(orc_ephemeris* e,
                                 pb_fixed_step_flow_t const* flow,
                                 int32_t n_threads) {
  std::int64_t const n = flow->n;
  if (n_threads <= 0) {
    n_threads = omp_get_max_threads();
  }
  std::int64_t const chunks = std::min<std::int64_t>(n_threads, n);
  int32_t status = kOk;
  std::int64_t steps = 0, samples = 0;
  DP final_time{flow->time[0], flow->time[1]};
  std::vector<int32_t> chunk_status(chunks, kOk);
#pragma omp parallel for num_threads(n_threads) schedule(static, 1)
  for (std::int64_t chunk = 0; chunk < chunks; ++chunk) {
    std::int64_t const first = n * chunk / chunks;
    std::int64_t const count = n * (chunk + 1) / chunks - first;
    State state = StateFromPlanes(flow->state, n, first, count, flow->time[0],
                                  flow->time[1]);
    std::int64_t local_steps = 0, local_samples = 0;
    AppendState append = [&](State const& s) {
      ++local_steps;
      if (flow->samples != nullptr && flow->sample_every > 0 &&
          local_steps % flow->sample_every == 0 &&
          local_samples < flow->max_samples) {
        double* out = flow->samples + local_samples * 6 * n;
        for (std::int64_t i = 0; i < count; ++i) {
          for (int c = 0; c < 3; ++c) {
            out[c * n + first + i] = s.positions[3 * i + c].value;
            out[(3 + c) * n + first + i] = s.velocities[3 * i + c].value;
          }
        }
        if (chunk == 0 && flow->sample_times != nullptr) {
          flow->sample_times[local_samples] = s.time.value;
        }
        ++local_samples;
      }
    };
    // ephemeris_body.hpp:366-385: accelerations[i] += intrinsic_acceleration(t).
    RightHandSide const gravitation = e->e->MasslessRightHandSide();
    RightHandSide rhs = gravitation;
    std::vector<Ephemeris::Burn> burns;
    if (flow->burns != nullptr) {
      for (std::int64_t i = 0; i < count; ++i) {
        double const* const b = flow->burns + first + i;
        burns.push_back({{b[0], b[n], b[2 * n]}, b[3 * n], b[4 * n], b[5 * n],
                         b[6 * n], b[7 * n]});
      }
      rhs = [&gravitation, &burns](double const t,
                                   std::vector<double> const& q,
                                   std::vector<double>& g) -> int32_t {
        int32_t const error = gravitation(t, q, g);
        for (std::size_t i = 0; i < burns.size(); ++i) {
          Vec3 acceleration{g[3 * i], g[3 * i + 1], g[3 * i + 2]};
          acceleration += burns[i].IntrinsicAcceleration(t);
          g[3 * i] = acceleration.x;
          g[3 * i + 1] = acceleration.y;
          g[3 * i + 2] = acceleration.z;
        }
        return error;
      };
    }
    FixedStepInstance instance(flow->method, rhs, state, append, flow->step);
    chunk_status[chunk] = instance.Solve(flow->t_final);
    PlanesFromState(instance.state(), flow->state, n, first, count);
    if (chunk == 0) {
      final_time = instance.state().time;
      steps = local_steps;
      samples = local_samples;
    }
  }
  for (int32_t s : chunk_status) {
    Update(status, s);
  }
  flow->time[0] = final_time.value;
  flow->time[1] = final_time.error;
  if (flow->n_steps != nullptr) {
    *flow->n_steps = steps;
  }
  if (flow->n_samples != nullptr) {
    *flow->n_samples = samples;
  }
  if (flow->status != nullptr) {
    *flow->status = status;
  }
  return kOk;
}

// FlowWithAdaptiveStep, one trajectory per probe, OpenMP over probes.
int32_t orc_flow_with_adaptive_step(orc_ephemeris* e,
                                    pb_adaptive_step_flow_t const* flow,
                                    int32_t n_threads) {
  std::int64_t const n = flow->n;
  if (n_threads <= 0) {
    n_threads = omp_get_max_threads();
  }
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 16)
  for (std::int64_t i = 0; i < n; ++i) {
    State state = StateFromPlanes(flow->state, n, i, 1, flow->time[i],
                                  flow->time[n + i]);
    std::int64_t accepted = 0, evaluations = 0, rejections = 0;
    AppendState append = [&](State const& s) {
      if (flow->step_log != nullptr && accepted < flow->step_log_capacity) {
        flow->step_log[accepted * n + i] = s.time.value;
      }
      if (flow->state_log != nullptr && accepted < flow->step_log_capacity) {
        double* const out = flow->state_log + accepted * 6 * n;
        for (int c = 0; c < 3; ++c) {
          out[c * n + i] = s.positions[c].value;
          out[(3 + c) * n + i] = s.velocities[c].value;
        }
      }
      ++accepted;
    };
    Ephemeris::Burn burn;
    if (flow->burns != nullptr) {
      double const* const b = flow->burns + i;
      burn = {{b[0], b[n], b[2 * n]}, b[3 * n], b[4 * n], b[5 * n], b[6 * n],
              b[7 * n]};
      if (flow->frenet_centres != nullptr && flow->frenet_centres[i] >= 0) {
        burn.frenet_centre = e->internal_of_caller[flow->frenet_centres[i]];
      }
    }
    int32_t const status = e->e->FlowWithAdaptiveStep(
        state, flow->t_final, flow->length_integration_tolerance,
        flow->speed_integration_tolerance, flow->max_steps, append,
        &evaluations, &rejections, flow->burns != nullptr ? &burn : nullptr,
        /*generalized=*/flow->method == PB_FINE_1987_RKNG_34);
    PlanesFromState(state, flow->state, n, i, 1);
    flow->time[i] = state.time.value;
    flow->time[n + i] = state.time.error;
    if (flow->status != nullptr) {
      flow->status[i] = status;
    }
    if (flow->accepted_steps != nullptr) {
      flow->accepted_steps[i] = accepted;
    }
    if (flow->rejected_steps != nullptr) {
      flow->rejected_steps[i] = rejections;
    }
    if (flow->evaluations != nullptr) {
      flow->evaluations[i] = evaluations;
    }
  }
  return kOk;
}

// Ensemble of independent massive systems sharing the gravity model of `e`
// (astronomy/trappist_dynamics_test.cpp:275-297).  state: 12 planes,
// [plane][body internal][system]; advanced from initial_time to t_final with
// a fresh instance per system.  Returns the states; OpenMP over systems.
int32_t orc_nbody_ensemble_flow(orc_ephemeris* e, int32_t method, double step,
                                std::int64_t n_systems, double* state,
                                double* time, double t_final,
                                std::int64_t* n_steps, int32_t n_threads) {
  int const nb = e->e->size();
  if (n_threads <= 0) {
    n_threads = omp_get_max_threads();
  }
  DP final_time{time[0], time[1]};
  std::int64_t steps0 = 0;
#pragma omp parallel for num_threads(n_threads) schedule(static)
  for (std::int64_t s = 0; s < n_systems; ++s) {
    State st;
    st.time = {time[0], time[1]};
    st.positions.resize(3 * nb);
    st.velocities.resize(3 * nb);
    auto at = [&](int plane, int b) -> double& {
      return state[(static_cast<std::int64_t>(plane) * nb + b) * n_systems + s];
    };
    for (int b = 0; b < nb; ++b) {
      for (int c = 0; c < 3; ++c) {
        st.positions[3 * b + c] = {at(c, b), at(3 + c, b)};
        st.velocities[3 * b + c] = {at(6 + c, b), at(9 + c, b)};
      }
    }
    std::int64_t steps = 0;
    FixedStepInstance instance(
        method,
        [e](double t, std::vector<double> const& q, std::vector<double>& g) {
          return e->e->ComputeGravitationalAccelerationBetweenAllMassiveBodies(
              t, q, g);
        },
        st, [&](State const&) { ++steps; }, step);
    instance.Solve(t_final);
    State const& out = instance.state();
    for (int b = 0; b < nb; ++b) {
      for (int c = 0; c < 3; ++c) {
        at(c, b) = out.positions[3 * b + c].value;
        at(3 + c, b) = out.positions[3 * b + c].error;
        at(6 + c, b) = out.velocities[3 * b + c].value;
        at(9 + c, b) = out.velocities[3 * b + c].error;
      }
    }
    if (s == 0) {
      final_time = out.time;
      steps0 = steps;
    }
  }
  time[0] = final_time.value;
  time[1] = final_time.error;
  if (n_steps != nullptr) {
    *n_steps = steps0;
  }
  return kOk;
}

// The large-N all-pairs force for a subset of the targets, in the summation
// order of the product's kernel (principia_b200/csrc/pb_kernels_allpairs.cuh;
// the arithmetic of one interaction is ephemeris_body.hpp:1366-1374): the
// sources are cut into L = clamp(2^23 / n, 8, 64) contiguous sub-ranges of
// `chunk` = ceil(n / L) rounded up to a multiple of 64 sources; each sub-range
// is summed sequentially (the target itself skipped), the sub-ranges are
// folded in order by groups of 8, and the groups are folded in order.  This is
// NOT the reference's pair order (scattered third-law updates), which a subset
// of targets cannot follow; the tests compare that order to a tolerance at
// small n.  q: 3 planes of n; targets: n_targets indices; a: [target][xyz].
int32_t orc_allpairs_target_accelerations(int64_t n, double const* mu,
                                          double const* q, int64_t n_targets,
                                          int64_t const* targets, double* a) {
  int sub_ranges = 64;
  while (sub_ranges > 8 &&
         static_cast<int64_t>(sub_ranges) * n > (int64_t{1} << 23)) {
    sub_ranges /= 2;
  }
  int64_t const per_range = (n + sub_ranges - 1) / sub_ranges;
  int64_t const chunk = (per_range + 63) / 64 * 64;
  double const* const qx = q;
  double const* const qy = q + n;
  double const* const qz = q + 2 * n;
#pragma omp parallel for schedule(dynamic)
  for (int64_t k = 0; k < n_targets; ++k) {
    int64_t const i = targets[k];
    double total[3] = {0.0, 0.0, 0.0};
    for (int group = 0; group < sub_ranges / 8; ++group) {
      double group_sum[3] = {0.0, 0.0, 0.0};
      for (int w = 0; w < 8; ++w) {
        int64_t const j_begin = (static_cast<int64_t>(group) * 8 + w) * chunk;
        int64_t const j_end = std::min(j_begin + chunk, n);
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int64_t j = j_begin; j < j_end; ++j) {
          if (j == i) {
            continue;
          }
          double const dx = qx[j] - qx[i];
          double const dy = qy[j] - qy[i];
          double const dz = qz[j] - qz[i];
          double const d2 = dx * dx + dy * dy + dz * dz;
          double const d = std::sqrt(d2);
          double const one_over_d3 = d / (d2 * d2);
          double const mu_over_d3 = mu[j] * one_over_d3;
          ax += dx * mu_over_d3;
          ay += dy * mu_over_d3;
          az += dz * mu_over_d3;
        }
        if (w == 0) {
          group_sum[0] = ax;
          group_sum[1] = ay;
          group_sum[2] = az;
        } else {
          group_sum[0] = group_sum[0] + ax;
          group_sum[1] = group_sum[1] + ay;
          group_sum[2] = group_sum[2] + az;
        }
      }
      for (int c = 0; c < 3; ++c) {
        total[c] = group == 0 ? group_sum[c] : total[c] + group_sum[c];
      }
    }
    a[3 * k] = total[0];
    a[3 * k + 1] = total[1];
    a[3 * k + 2] = total[2];
  }
  return 0;
}

int32_t orc_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
