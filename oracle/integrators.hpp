// ORACLE -- TEST INFRASTRUCTURE ONLY (see numerics.hpp).
//
// CPU restatement of the three steppers on the hot path.  States are flat
// arrays of scalar components: every DoublePrecision<Vector> operation of the
// reference is component-wise, so a Position is simply three consecutive
// components.
#pragma once

#include <algorithm>
#include <cstdint>
#include <deque>
#include <functional>
#include <vector>

#include "numerics.hpp"

namespace orc {

// integrators/ordinary_differential_equations.hpp:156-220.
struct State {
  DP time;
  std::vector<DP> positions;
  std::vector<DP> velocities;
};
using RightHandSide = std::function<int32_t(
    double t, std::vector<double> const& q, std::vector<double>& g)>;
using AppendState = std::function<void(State const&)>;
// ExplicitSecondOrderOrdinaryDifferentialEquation: the right-hand side also
// takes the velocities (integrators/ordinary_differential_equations.hpp).
using GeneralizedRightHandSide = std::function<int32_t(
    double t, std::vector<double> const& q, std::vector<double> const& v,
    std::vector<double>& g)>;

// ---------------------------------------------------------------------------
// Tableaux, integrators/methods.hpp.

enum Composition { BA, ABA, BAB };

struct SrknMethod {
  Composition composition;
  std::vector<double> a, b, c;
  int evaluations;
  int stages() const { return static_cast<int>(a.size()); }
};

// integrators/symplectic_runge_kutta_nyström_integrator_body.hpp:198-208: the
// nodes are accumulated in double-double.
inline void ComputeNodes(SrknMethod& m) {
  DP c_i;
  m.c.resize(m.a.size());
  for (std::size_t i = 0; i < m.a.size(); ++i) {
    m.c[i] = c_i.value;
    DP a_i;
    a_i.value = m.a[i];
    c_i = c_i + a_i;
  }
}

// methods.hpp:379-419.
inline SrknMethod const& BlanesMoan2002SRKN14A() {
  static SrknMethod const m = [] {
    SrknMethod m;
    m.composition = ABA;
    m.evaluations = 14;
    m.a = {0.037859319840611600,   0.10263563310243500,
           -0.025867888266558700,  0.31424140307144700,
           -0.13014445951741500,   0.10641770036954300,
           -0.0087942431285105800, 0.2073050690568954,
           -0.0087942431285105800, 0.10641770036954300,
           -0.13014445951741500,   0.31424140307144700,
           -0.025867888266558700,  0.10263563310243500,
           0.037859319840611600};
    m.b = {0.0,
           0.091719152624461650,  0.18398317000500600,
           -0.056534365832888270, 0.0049146887747128540,
           0.14376112716835800,   0.32856769374680400,
           -0.19641146648645423,  -0.19641146648645423,
           0.32856769374680400,   0.14376112716835800,
           0.0049146887747128540, -0.056534365832888270,
           0.18398317000500600,   0.091719152624461650};
    ComputeNodes(m);
    return m;
  }();
  return m;
}

// methods.hpp:922-949.
inline SrknMethod const& McLachlanAtela1992Order5Optimal() {
  static SrknMethod const m = [] {
    SrknMethod m;
    m.composition = BA;
    m.evaluations = 6;
    m.a = {0.339839625839110000,  -0.088601336903027329,
           0.5858564768259621188, -0.603039356536491888,
           0.3235807965546976394, 0.4423637942197494587};
    m.b = {0.1193900292875672758,  0.6989273703824752308,
           -0.1713123582716007754, 0.4012695022513534480,
           0.0107050818482359840,  -0.0589796254980311632};
    ComputeNodes(m);
    return m;
  }();
  return m;
}

// methods.hpp:574-597, DormandالمكاوىPrince1986RKN434FM.
struct ErknMethod {
  int stages;
  int lower_order;
  bool first_same_as_last;
  std::vector<double> c;
  std::vector<double> a;  // Strictly lower triangular, (i, j) -> i(i-1)/2 + j.
  std::vector<double> b_hat, b_hat_prime, b, b_prime;
  // The velocity stages of a *generalized* method (empty otherwise).
  std::vector<double> a_prime;
  double A(int const i, int const j) const { return a[i * (i - 1) / 2 + j]; }
  double APrime(int const i, int const j) const {
    return a_prime[i * (i - 1) / 2 + j];
  }
};

inline ErknMethod const& DormandElMikkawyPrince1986RKN434FM() {
  static ErknMethod const m = [] {
    ErknMethod m;
    m.stages = 4;
    m.lower_order = 3;
    m.first_same_as_last = true;
    m.c = {0.0, 1.0 / 4.0, 7.0 / 10.0, 1.0};
    m.a = {1.0 / 32.0,
           7.0 / 1000.0, 119.0 / 500.0,
           1.0 / 14.0,   8.0 / 27.0,    25.0 / 189.0};
    m.b_hat = {1.0 / 14.0, 8.0 / 27.0, 25.0 / 189.0, 0.0};
    m.b_hat_prime = {1.0 / 14.0, 32.0 / 81.0, 250.0 / 567.0, 5.0 / 54.0};
    m.b = {-7.0 / 150.0, 67.0 / 150.0, 3.0 / 20.0, -1.0 / 20.0};
    m.b_prime = {13.0 / 21.0, -20.0 / 27.0, 275.0 / 189.0, -1.0 / 3.0};
    return m;
  }();
  return m;
}

// Fine1987RKNG34, methods.hpp:543-570: an embedded explicit *generalized*
// Runge-Kutta-Nyström method (the right-hand side may depend on the velocity,
// integrators/embedded_explicit_generalized_runge_kutta_nyström_integrator_
// body.hpp:44-300).  With a right-hand side that ignores the velocity -- the
// ephemeris with no or time-only intrinsic acceleration -- the generalized
// Solve performs exactly the operations of the ErknInstance below with this
// tableau (the velocity stages, built with a', feed nothing).  5 stages, no
// FSAL; b = b_hat - delta as the reference forms them.
inline ErknMethod const& Fine1987RKNG34() {
  static ErknMethod const m = [] {
    ErknMethod m;
    m.stages = 5;
    m.lower_order = 3;
    m.first_same_as_last = false;
    m.c = {0.0, 2 / 9.0, 1 / 3.0, 3 / 4.0, 1.0};
    m.a = {2 / 81.0,
           1 / 36.0,   1 / 36.0,
           9 / 128.0,  0.0,        27 / 128.0,
           11 / 60.0,  -3 / 20.0,  9 / 25.0,  8 / 75.0};
    m.a_prime = {2 / 9.0,
                 1 / 12.0,     1 / 4.0,
                 69 / 128.0,   -234 / 128.0,  135 / 64.0,
                 -17 / 12.0,   27 / 4.0,      -27 / 5.0,   16 / 15.0};
    m.b_hat = {19 / 180.0, 0.0, 63 / 200.0, 16 / 225.0, 1 / 120.0};
    m.b_hat_prime = {1 / 9.0, 0.0, 9 / 20.0, 16 / 45.0, 1 / 12.0};
    std::vector<double> const delta = {25 / 1116.0, 0.0, -63 / 1240.0,
                                       64 / 1395.0, -13 / 744.0};
    std::vector<double> const delta_prime = {2 / 125.0, 0.0, -27 / 625.0,
                                             32 / 625.0, -3 / 125.0};
    for (int i = 0; i < 5; ++i) {
      m.b.push_back(m.b_hat[i] - delta[i]);
      m.b_prime.push_back(m.b_hat_prime[i] - delta_prime[i]);
    }
    return m;
  }();
  return m;
}

// methods.hpp:998-1007, :1040-1055 and
// integrators/cohen_hubbard_oesterwinter_body.hpp (orders 8 and 12).
struct SlmsMethod {
  int order;
  std::vector<double> alpha, beta_numerator;
  double beta_denominator;
  std::vector<double> cho_numerators;
  double cho_denominator;
};

inline SlmsMethod const& Quinlan1999Order8A() {
  static SlmsMethod const m = {
      8,
      {1.0, -2.0, 2.0, -2.0, 2.0},
      {0.0, 22081.0, -29418.0, 75183.0, -75212.0},
      15120.0,
      {416173.0, 950684.0, -1025097.0, 1059430.0, -768805.0, 362112.0,
       -99359.0, 12062.0},
      1814400.0};
  return m;
}

inline SlmsMethod const& QuinlanTremaine1990Order12() {
  static SlmsMethod const m = {
      12,
      {1.0, -2.0, 2.0, -1.0, 0.0, 0.0, 0.0},
      {0.0, 90987349.0, -229596838.0, 812627169.0, -1628539944.0,
       2714971338.0, -3041896548.0},
      53222400.0,
      {92158447389.0, 301307140046.0, -554452444015.0, 1035372815340.0,
       -1505150506950.0, 1655690777412.0, -1363696062582.0, 828085590240.0,
       -360089099415.0, 106193749950.0, -19043781851.0, 1569102436.0},
      435891456000.0};
  return m;
}

// ---------------------------------------------------------------------------
// integrators/symplectic_runge_kutta_nyström_integrator_body.hpp:24-144.

class SrknInstance {
 public:
  SrknInstance(SrknMethod const& method, RightHandSide rhs, State state,
               AppendState append_state, double const step)
      : method_(method),
        rhs_(std::move(rhs)),
        state_(std::move(state)),
        append_state_(std::move(append_state)),
        step_(step) {}

  State const& state() const { return state_; }
  std::int64_t evaluations = 0;

  int32_t Solve(double const t_final) {
    auto const& a = method_.a;
    auto const& b = method_.b;
    auto const& c = method_.c;
    int const stages = method_.stages();
    Composition const composition = method_.composition;
    int const dimension = static_cast<int>(state_.positions.size());

    double const h = step_;
    double const abs_h = h < 0 ? -h : h;  // Sign(step) * h.
    DP& t = state_.time;
    std::vector<double> dq(dimension), dv(dimension);
    std::vector<DP>& q = state_.positions;
    std::vector<DP>& v = state_.velocities;
    std::vector<double> q_stage(dimension), g(dimension);
    int const first_stage = composition == BA ? 0 : 1;
    int32_t status = kOk;

    if (composition == BAB) {
      for (int k = 0; k < dimension; ++k) {
        q_stage[k] = q[k].value;
      }
      Update(status, rhs_(t.value, q_stage, g));
      ++evaluations;
    }

    while (abs_h <= std::abs((t_final - t.value) - t.error)) {
      std::fill(dq.begin(), dq.end(), 0.0);
      std::fill(dv.begin(), dv.end(), 0.0);

      if (first_stage == 1) {
        for (int k = 0; k < dimension; ++k) {
          if (composition == BAB) {
            dv[k] += h * b[0] * g[k];
          }
          dq[k] += h * a[0] * (v[k].value + dv[k]);
        }
      }

      for (int i = first_stage; i < stages; ++i) {
        for (int k = 0; k < dimension; ++k) {
          q_stage[k] = q[k].value + dq[k];
        }
        Update(status, rhs_(t.value + (t.error + c[i] * h), q_stage, g));
        ++evaluations;
        for (int k = 0; k < dimension; ++k) {
          dv[k] += h * b[i] * g[k];
          dq[k] += h * a[i] * (v[k].value + dv[k]);
        }
      }

      Increment(t, h);
      for (int k = 0; k < dimension; ++k) {
        Increment(q[k], dq[k]);
        Increment(v[k], dv[k]);
      }
      append_state_(state_);
      if (status == kAborted) {
        return status;
      }
    }
    return status;
  }

 private:
  SrknMethod const& method_;
  RightHandSide rhs_;
  State state_;
  AppendState append_state_;
  double step_;
};

// ---------------------------------------------------------------------------
// integrators/embedded_explicit_runge_kutta_nyström_integrator_body.hpp:44-258.

struct StateError {
  std::vector<double> position_error;
  std::vector<double> velocity_error;
};
using ToleranceToErrorRatio =
    std::function<double(double h, State const&, StateError const&)>;

struct AdaptiveParameters {
  double first_step;
  double safety_factor = 0.9;
  std::int64_t max_steps = std::numeric_limits<std::int64_t>::max();
  bool last_step_is_exact = true;
};

class ErknInstance {
 public:
  ErknInstance(ErknMethod const& method, RightHandSide rhs, State state,
               AppendState append_state, ToleranceToErrorRatio ratio,
               AdaptiveParameters const& parameters)
      : method_(method),
        rhs_(std::move(rhs)),
        state_(std::move(state)),
        append_state_(std::move(append_state)),
        ratio_(std::move(ratio)),
        parameters_(parameters),
        step_(parameters.first_step) {}

  State const& state() const { return state_; }
  std::int64_t evaluations = 0;
  std::int64_t rejections = 0;
  // If set (and the method has velocity stages), the right-hand side of the
  // generalized integrator; rhs_ is then unused.
  GeneralizedRightHandSide generalized_rhs;

  int32_t Solve(double const t_final) {
    auto const& m = method_;
    int const stages = m.stages;
    int const dimension = static_cast<int>(state_.positions.size());
    double const direction = parameters_.first_step < 0 ? -1.0 : 1.0;

    double& h = step_;
    DP& t = state_.time;
    std::vector<double> dq_hat(dimension), dv_hat(dimension);
    std::vector<DP>& q_hat = state_.positions;
    std::vector<DP>& v_hat = state_.velocities;
    StateError error_estimate;
    error_estimate.position_error.resize(dimension);
    error_estimate.velocity_error.resize(dimension);
    std::vector<double> q_stage(dimension);
    std::vector<double> v_stage(dimension);
    std::vector<std::vector<double>> g(stages, std::vector<double>(dimension));

    bool at_end = false;
    double tolerance_to_error_ratio = 0;
    int first_stage = 0;
    std::int64_t step_count = 0;
    int32_t status = kOk;
    int32_t step_status = kOk;
    State final_state = state_;
    bool has_final_state = false;
    bool first_iteration = true;

    while (!at_end) {
      do {
        if (!first_iteration) {
          step_status = kOk;
          h *= parameters_.safety_factor *
               Root(m.lower_order + 1, tolerance_to_error_ratio);
          if (t.value + (t.error + h) == t.value) {
            return kFailedPrecondition;
          }
        }
        first_iteration = false;

        if (parameters_.last_step_is_exact) {
          double const time_to_end = (t_final - t.value) - t.error;
          at_end = direction * h >= direction * time_to_end;
          if (at_end) {
            h = time_to_end;
            final_state = state_;
            has_final_state = true;
          }
        }

        double const h2 = h * h;

        for (int i = first_stage; i < stages; ++i) {
          double const t_stage =
              (parameters_.last_step_is_exact && at_end && m.c[i] == 1.0)
                  ? t_final
                  : t.value + (t.error + m.c[i] * h);
          for (int k = 0; k < dimension; ++k) {
            double sum = 0;
            for (int j = 0; j < i; ++j) {
              sum += m.A(i, j) * g[j][k];
            }
            q_stage[k] = q_hat[k].value + h * m.c[i] * v_hat[k].value +
                         h2 * sum;
          }
          if (generalized_rhs) {
            // embedded_explicit_generalized_runge_kutta_nyström_integrator_
            // body.hpp:186-201: the velocity stage.
            for (int k = 0; k < dimension; ++k) {
              double sum_prime = 0;
              for (int j = 0; j < i; ++j) {
                sum_prime += m.APrime(i, j) * g[j][k];
              }
              v_stage[k] = v_hat[k].value + h * sum_prime;
            }
            Update(step_status, generalized_rhs(t_stage, q_stage, v_stage, g[i]));
          } else {
            Update(step_status, rhs_(t_stage, q_stage, g[i]));
          }
          ++evaluations;
        }

        for (int k = 0; k < dimension; ++k) {
          double sum_b_hat = 0, sum_b = 0, sum_b_hat_prime = 0,
                 sum_b_prime = 0;
          for (int i = 0; i < stages; ++i) {
            sum_b_hat += m.b_hat[i] * g[i][k];
            sum_b += m.b[i] * g[i][k];
            sum_b_hat_prime += m.b_hat_prime[i] * g[i][k];
            sum_b_prime += m.b_prime[i] * g[i][k];
          }
          dq_hat[k] = h * v_hat[k].value + h2 * sum_b_hat;
          double const dq = h * v_hat[k].value + h2 * sum_b;
          dv_hat[k] = h * sum_b_hat_prime;
          double const dv = h * sum_b_prime;
          error_estimate.position_error[k] = dq - dq_hat[k];
          error_estimate.velocity_error[k] = dv - dv_hat[k];
        }
        tolerance_to_error_ratio = ratio_(h, state_, error_estimate);
        if (tolerance_to_error_ratio < 1.0) {
          ++rejections;
        }
      } while (tolerance_to_error_ratio < 1.0);

      Update(status, step_status);

      if (!parameters_.last_step_is_exact &&
          t.value + (t.error + h) > t_final) {
        final_state = state_;
        has_final_state = true;
        break;
      }

      if (m.first_same_as_last) {
        std::swap(g.front(), g.back());
        first_stage = 1;
      }

      Increment(t, h);
      for (int k = 0; k < dimension; ++k) {
        Increment(q_hat[k], dq_hat[k]);
        Increment(v_hat[k], dv_hat[k]);
      }
      append_state_(state_);
      ++step_count;
      if (step_count == parameters_.max_steps && !at_end) {
        return kResourceExhausted;
      }
    }
    if (has_final_state) {
      state_ = final_state;
    }
    return status;
  }

 private:
  ErknMethod const& method_;
  RightHandSide rhs_;
  State state_;
  AppendState append_state_;
  ToleranceToErrorRatio ratio_;
  AdaptiveParameters parameters_;
  double step_;
};

// ---------------------------------------------------------------------------
// integrators/symmetric_linear_multistep_integrator_body.hpp:22-152,281-321 and
// integrators/starter_body.hpp:23-93.

class SlmsInstance {
 public:
  static constexpr int startup_step_divisor = 16;

  struct Step {
    std::vector<DP> displacements;
    std::vector<double> accelerations;
    DP time;
  };

  SlmsInstance(SlmsMethod const& method, RightHandSide rhs, State state,
               AppendState append_state, double const step)
      : method_(method),
        rhs_(std::move(rhs)),
        state_(std::move(state)),
        append_state_(std::move(append_state)),
        step_(step) {}

  State const& state() const { return state_; }
  std::deque<Step> const& previous_steps() const { return previous_steps_; }
  bool started() const {
    return static_cast<int>(previous_steps_.size()) == method_.order;
  }
  std::int64_t evaluations = 0;

  int32_t Solve(double const t_final) {
    auto const& alpha = method_.alpha;
    auto const& beta_numerator = method_.beta_numerator;
    double const beta_denominator = method_.beta_denominator;

    if (!started()) {
      StarterSolve(t_final);
      if (!started()) {
        return kOk;
      }
    }

    int const dimension =
        static_cast<int>(previous_steps_.back().displacements.size());
    double const h = step_;
    DP t = previous_steps_.back().time;
    int const k = method_.order;
    int32_t status = kOk;
    std::vector<double> positions(dimension);
    std::vector<DP> sum_minus_alpha_q(dimension);
    std::vector<double> sum_beta_numerator_a(dimension);

    while (h <= (t_final - t.value) - t.error) {
      // j = 0.
      {
        Step const& front = previous_steps_[0];
        for (int d = 0; d < dimension; ++d) {
          sum_minus_alpha_q[d] = Scale(-alpha[0], front.displacements[d]);
          sum_beta_numerator_a[d] = beta_numerator[0] * front.accelerations[d];
        }
      }
      // Generic j paired with k - j.
      for (int j = 1; j < k / 2; ++j) {
        Step const& front = previous_steps_[j];
        Step const& back = previous_steps_[k - j];
        for (int d = 0; d < dimension; ++d) {
          sum_minus_alpha_q[d] =
              sum_minus_alpha_q[d] - Scale(alpha[j], front.displacements[d]);
          sum_minus_alpha_q[d] =
              sum_minus_alpha_q[d] - Scale(alpha[j], back.displacements[d]);
          sum_beta_numerator_a[d] +=
              beta_numerator[j] *
              (front.accelerations[d] + back.accelerations[d]);
        }
      }
      // j = k / 2.
      {
        Step const& middle = previous_steps_[k / 2];
        for (int d = 0; d < dimension; ++d) {
          sum_minus_alpha_q[d] =
              sum_minus_alpha_q[d] -
              Scale(alpha[k / 2], middle.displacements[d]);
          sum_beta_numerator_a[d] +=
              beta_numerator[k / 2] * middle.accelerations[d];
        }
      }

      Increment(t, h);
      Step current_step;
      current_step.time = t;
      current_step.accelerations.resize(dimension);
      current_step.displacements.reserve(dimension);

      for (int d = 0; d < dimension; ++d) {
        DP& current_displacement = sum_minus_alpha_q[d];
        Increment(current_displacement,
                  h * h * sum_beta_numerator_a[d] / beta_denominator);
        current_step.displacements.push_back(current_displacement);
        DP const current_position = DP() + current_displacement;
        positions[d] = current_position.value;
        state_.positions[d] = current_position;
      }
      Update(status, rhs_(t.value, positions, current_step.accelerations));
      ++evaluations;
      previous_steps_.push_back(std::move(current_step));
      previous_steps_.pop_front();

      ComputeVelocityUsingCohenHubbardOesterwinter();

      state_.time = t;
      append_state_(state_);
      if (status == kAborted) {
        return status;
      }
    }
    return status;
  }

 private:
  // symmetric_linear_multistep_integrator_body.hpp:246-263.
  void FillStepFromState(State const& state, Step& step) {
    std::vector<double> positions;
    step.time = state.time;
    for (auto const& position : state.positions) {
      step.displacements.push_back(position - DP());
      positions.push_back(position.value);
    }
    step.accelerations.resize(step.displacements.size());
    rhs_(step.time.value, positions, step.accelerations);
    ++evaluations;
  }

  // starter_body.hpp:23-74.
  void StarterSolve(double const s_final) {
    int const steps = method_.order;
    if (previous_steps_.empty()) {
      previous_steps_.emplace_back();
      FillStepFromState(state_, previous_steps_.back());
    }
    double const startup_step = step_ / startup_step_divisor;

    auto const startup_append_state = [this, steps](State const& state) {
      if (static_cast<int>(previous_steps_.size()) < steps) {
        state_ = state;
        if (++startup_step_index_ % startup_step_divisor == 0) {
          previous_steps_.emplace_back();
          FillStepFromState(state_, previous_steps_.back());
          append_state_(state);
        }
      }
    };

    SrknInstance startup_instance(
        BlanesMoan2002SRKN14A(),
        [this](double t, std::vector<double> const& q, std::vector<double>& g) {
          ++evaluations;
          return rhs_(t, q, g);
        },
        state_, startup_append_state, startup_step);
    double const target = std::min(
        state_.time.value +
            (steps - static_cast<int>(previous_steps_.size())) * step_ +
            step_ / 2.0,
        s_final);
    startup_instance.Solve(target);
  }

  // symmetric_linear_multistep_integrator_body.hpp:281-321.
  void ComputeVelocityUsingCohenHubbardOesterwinter() {
    int const dimension =
        static_cast<int>(previous_steps_.back().displacements.size());
    int const n = static_cast<int>(previous_steps_.size());
    for (int d = 0; d < dimension; ++d) {
      DP& velocity = state_.velocities[d];
      DP const displacement_change = previous_steps_[n - 1].displacements[d] -
                                     previous_steps_[n - 2].displacements[d];
      velocity = DP();
      velocity.value =
          (displacement_change.value + displacement_change.error) / step_;
      double weighted_accelerations = 0;
      for (std::size_t i = 0; i < method_.cho_numerators.size(); ++i) {
        weighted_accelerations +=
            method_.cho_numerators[i] * previous_steps_[n - 1 - i].accelerations[d];
      }
      velocity.value +=
          weighted_accelerations * step_ / method_.cho_denominator;
    }
  }

  SlmsMethod const& method_;
  RightHandSide rhs_;
  State state_;
  AppendState append_state_;
  double step_;
  std::deque<Step> previous_steps_;
  std::int64_t startup_step_index_ = 0;
};

}  // namespace orc
