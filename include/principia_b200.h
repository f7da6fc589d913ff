/* principia_b200 -- C ABI of the B200-native batched gravitational-flow engine.
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * mockingbirdnest/Principia's physics::Ephemeris.  The reference has no FFI
 * seam on this path (it is C++ templates end to end); the entry points below
 * are what a C-ABI replacement of the path binds, one per reference routine,
 * each citing the reference interface it replaces.  The C++ facade
 * `principia_b200/cpp/ephemeris.hpp` re-creates the reference's class surface
 * on top of this header; INTEGRATION.md shows the reference-side change.
 *
 * Conventions
 *  - Every quantity is one IEEE binary64 in SI units (m, s, m^3/s^2, rad), as
 *    `Quantity<D>` is in the reference (quantities/quantities.hpp).  Instants
 *    are seconds from J2000 (astronomy/epoch.hpp:16-24).
 *  - Return values are absl::StatusCode numbers (pb_status_t); the reference
 *    returns absl::Status on the same routines.
 *  - Plain pointers and sizes only.  Functions without a `_device` suffix take
 *    HOST pointers and perform the host<->device copies themselves; `_device`
 *    variants take pointers into the ephemeris' CUDA device and run on the
 *    given CUDA stream (a `cudaStream_t` cast to void*, NULL = default).
 *  - There is no CPU fallback: every entry point fails with
 *    PB_FAILED_PRECONDITION if no CUDA device is usable.
 *  - Body indices in this API are the CALLER's indices (the order of the
 *    `bodies` array given to pb_ephemeris_create), like `unowned_bodies_` in
 *    the reference (physics/ephemeris.hpp:507-519).  Internally bodies are
 *    reordered like the reference's constructor does
 *    (physics/ephemeris_body.hpp:138-158): that order is the floating-point
 *    summation order, pb_ephemeris_internal_order exposes it.
 */
#ifndef PRINCIPIA_B200_H_
#define PRINCIPIA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t pb_status_t;
enum {
  PB_OK = 0,
  PB_CANCELLED = 1,
  PB_INVALID_ARGUMENT = 3,
  PB_DEADLINE_EXCEEDED = 4,  /* FlowWithAdaptiveStep could not reach t. */
  PB_RESOURCE_EXHAUSTED = 8, /* ReachedMaximalStepCount. */
  PB_FAILED_PRECONDITION = 9,/* VanishingStepSize; also "no CUDA device". */
  PB_ABORTED = 10,
  PB_OUT_OF_RANGE = 11,      /* "Collision detected". */
  PB_INTERNAL = 13           /* CUDA runtime error; see pb_last_error. */
};

/* integrators/methods.hpp; serialization kinds of the same names. */
enum {
  PB_BLANES_MOAN_2002_SRKN_14A = 0,
  PB_MCLACHLAN_ATELA_1992_ORDER_5_OPTIMAL = 1,
  PB_QUINLAN_1999_ORDER_8A = 2,
  PB_QUINLAN_TREMAINE_1990_ORDER_12 = 3,
  PB_DORMAND_ELMIKKAWY_PRINCE_1986_RKN_434FM = 16,
  /* integrators/methods.hpp:543-570, embedded explicit generalized RKN. */
  PB_FINE_1987_RKNG_34 = 17
};

/* Arithmetic of the fixed-step flows and of the all-pairs force.
 * PB_ARITHMETIC_EXACT: IEEE binary64 without contraction, the operations of
 * the reference in the reference's order (its -ffp-contract=off build,
 * Makefile:123-169 there); results are bit-identical to the CPU oracle.
 * PB_ARITHMETIC_CONTRACTED (opt-in): the same formulae compiled with fused
 * multiply-adds (and, in the all-pairs force, r^-3 from a refined reciprocal
 * square root): not bit-identical, within BASELINE.json's tolerance (relative
 * position error <= 1e-12 after 10^4 steps).  The adaptive flows are always
 * exact (identical accepted-step sequence). */
enum { PB_ARITHMETIC_EXACT = 0, PB_ARITHMETIC_CONTRACTED = 1 };

/* One massive body: MassiveBody / RotatingBody / OblateBody parameters
 * (physics/massive_body.hpp, rotating_body.hpp:45-77, oblate_body.hpp:39-75).
 * `cos`/`sin` are the fully normalised coefficients, lower triangular,
 * (n, m) -> n(n+1)/2 + m, (degree+1)(degree+2)/2 entries each
 * (numerics/fixed_arrays_body.hpp:722). */
typedef struct pb_body_t {
  double gravitational_parameter;
  double min_radius;           /* 0 for a non-rotating MassiveBody. */
  double mean_radius;
  double reference_angle;
  double reference_instant;
  double angular_frequency;
  double right_ascension_of_pole;
  double declination_of_pole;
  double reference_radius;     /* Oblate bodies only. */
  int32_t is_rotating;
  int32_t is_oblate;
  int32_t geopotential_degree; /* Oblate bodies only, 2..50. */
  int32_t is_zonal;
  const double* cos;
  const double* sin;
} pb_body_t;

typedef struct pb_ephemeris pb_ephemeris_t;

/* ---- Construction: Ephemeris::Ephemeris, physics/ephemeris.hpp:124-128,
 * ephemeris_body.hpp:94-170 (body ordering) and Geopotential::Geopotential,
 * geopotential_body.hpp:669-733 (damping thresholds from the tolerance). */
pb_status_t pb_ephemeris_create(int32_t n_bodies,
                                const pb_body_t* bodies,
                                double geopotential_tolerance,
                                int32_t cuda_device,
                                pb_ephemeris_t** ephemeris);
/* Refuses (and records an error for pb_last_error) while instances or
 * ensembles created from the ephemeris are alive: destroy those first. */
void pb_ephemeris_destroy(pb_ephemeris_t* ephemeris);
/* The message of the last non-OK status returned on the calling thread. */
const char* pb_last_error(void);

/* ---- Threading (physics/ephemeris.hpp:64-66: "const methods are
 * thread-safe; flows on distinct objects may run concurrently").  Every entry
 * point may be called from any thread.  Calls on one ephemeris run
 * concurrently -- each leases its own scratch, staging buffers and stream --
 * except that pb_ephemeris_append_segments waits for the flows and queries in
 * progress and holds the new ones back while it runs.  One call at a time per
 * pb_fixed_step_instance_t / pb_nbody_ensemble_t / pb_nbody_allpairs_t
 * (enforced by a lock per handle).  The fixed-step flow kernels keep the hot
 * body's harmonics in a __constant__ bank: two fixed-step flows run
 * concurrently when they use the same ephemeris and hot body, otherwise the
 * second waits for the first.
 *
 * ---- Cooperative cancellation (RETURN_IF_STOPPED after every step,
 * integrators/symplectic_runge_kutta_nyström_integrator_body.hpp:137,
 * embedded_explicit_runge_kutta_nyström_integrator_body.hpp:243,
 * symmetric_linear_multistep_integrator_body.hpp:149): after
 * pb_ephemeris_request_stop (any thread, lock-free), the flows in progress on
 * the ephemeris and on its instances and ensembles return PB_CANCELLED with
 * the state, time and sinks of the last completed step -- the fixed-step
 * massless flows at the next boundary between chunks of at most 256 steps, the
 * ensembles after the current step, the adaptive flow after the current
 * accepted step of every trajectory (per-trajectory status PB_CANCELLED).  The
 * request stays in force until pb_ephemeris_clear_stop. */
pb_status_t pb_ephemeris_request_stop(pb_ephemeris_t* ephemeris);
pb_status_t pb_ephemeris_clear_stop(pb_ephemeris_t* ephemeris);

/* order[i] = caller index of the body at internal position i; returns the
 * number of oblate bodies through *n_oblate (may be NULL). */
pb_status_t pb_ephemeris_internal_order(const pb_ephemeris_t* ephemeris,
                                        int32_t* order,
                                        int32_t* n_oblate);

/* Geopotential::degree_damping() / sectoral_damping() of an oblate body:
 * inner thresholds for degrees 0..*n_degrees-1 (geopotential.hpp:57-58).
 * `inner_thresholds` needs room for 51 doubles. */
pb_status_t pb_geopotential_damping(const pb_ephemeris_t* ephemeris,
                                    int32_t body,
                                    int32_t* n_degrees,
                                    double* inner_thresholds,
                                    double* sectoral_inner_threshold);

/* ---- ContinuousTrajectory segments (physics/continuous_trajectory.hpp).
 * Appends `n_segments` polynomials in the monomial basis to `body`'s
 * trajectory; mirrors the growth of `polynomials_` in
 * ContinuousTrajectory::Append (continuous_trajectory_body.hpp:82-127,758-858).
 * Segment s covers ]t_max[s-1], t_max[s]], is evaluated at (t - origin[s]) by
 * the Estrin scheme with FMA (numerics/polynomial_evaluators_body.hpp:72-160);
 * coefficients[s][k][xyz], k = 0..17 (entries above degree[s] are ignored).
 * `t_min` is only used by the first call for a body (first_time_). */
pb_status_t pb_ephemeris_append_segments(pb_ephemeris_t* ephemeris,
                                         int32_t body,
                                         int32_t n_segments,
                                         double t_min,
                                         const double* t_max,
                                         const double* origin,
                                         const int32_t* degree,
                                         const double* coefficients);
double pb_ephemeris_t_min(const pb_ephemeris_t* ephemeris);
double pb_ephemeris_t_max(const pb_ephemeris_t* ephemeris);

/* Ephemeris::EvaluateAllPositions, ephemeris.hpp:154-155: positions[3*body]. */
pb_status_t pb_ephemeris_evaluate_all_positions(pb_ephemeris_t* ephemeris,
                                                double t,
                                                double* positions);

/* ---- Accelerations and potential.
 * Ephemeris::ComputeGravitationalAccelerationOnMasslessBody (batched),
 * ephemeris.hpp:253-264 -> ...ByAllMassiveBodiesOnMasslessBodies,
 * ephemeris_body.hpp:1554-1591.  SoA: q = [x[n] y[n] z[n]], a likewise.
 * *status receives PB_OK or PB_OUT_OF_RANGE (bitwise OR over bodies and
 * probes, ephemeris_body.hpp:1437). */
pb_status_t pb_compute_gravitational_acceleration_on_massless_bodies(
    pb_ephemeris_t* ephemeris, double t, int64_t n, const double* q, double* a,
    pb_status_t* status);
pb_status_t pb_compute_gravitational_acceleration_on_massless_bodies_device(
    pb_ephemeris_t* ephemeris, double t, int64_t n, const double* q, double* a,
    int32_t* status_device, void* stream);

/* Ephemeris::ComputeGravitationalPotential (batched), ephemeris.hpp:283-285. */
pb_status_t pb_compute_gravitational_potential(pb_ephemeris_t* ephemeris,
                                               double t, int64_t n,
                                               const double* q,
                                               double* potential);

/* trajectory(body)->EvaluateVelocity(t) for every body
 * (continuous_trajectory_body.hpp:697-706, with EvaluateAllPositions the
 * EvaluateDegreesOfFreedom the jerk routines use): velocities[3*body]. */
pb_status_t pb_ephemeris_evaluate_all_velocities(pb_ephemeris_t* ephemeris,
                                                 double t,
                                                 double* velocities);

/* Ephemeris::ComputeGravitationalJerkOnMasslessBody (batched),
 * ephemeris.hpp:233-237, ephemeris_body.hpp:524-561: point masses only, like
 * the reference.  SoA: q, v, jerk = [x[n] y[n] z[n]]. */
pb_status_t pb_compute_gravitational_jerk_on_massless_bodies(
    pb_ephemeris_t* ephemeris, double t, int64_t n, const double* q,
    const double* v, double* jerk);
pb_status_t pb_compute_gravitational_jerk_on_massless_bodies_device(
    pb_ephemeris_t* ephemeris, double t, int64_t n, const double* q,
    const double* v, double* jerk, void* stream);

/* Ephemeris::ComputeGravitationalJerkOnMassiveBody(body, t) (positions ==
 * NULL: degrees of freedom evaluated at t from the trajectories) and
 * ...OnMassiveBodies(bodies, degrees of freedom) (positions, velocities AoS
 * [body][xyz], caller order), ephemeris.hpp:239-250,
 * ephemeris_body.hpp:563-605,1317-1340; Ephemeris::ComputeJacobianOnMassiveBody,
 * ephemeris.hpp:226-231, ephemeris_body.hpp:495-522.  For every body:
 * jerks[3*body] and jacobians[9*body] (row-major); either may be NULL. */
pb_status_t pb_compute_gravitational_jerk_and_jacobian_on_massive_bodies(
    pb_ephemeris_t* ephemeris, double t, const double* positions,
    const double* velocities, double* jerks, double* jacobians);

/* Ephemeris::ComputeGravitationalAccelerationOnMassiveBodies,
 * ephemeris.hpp:276-280 (positions given, AoS [body][xyz], caller order) and
 * ...OnMassiveBody, :268-271 (positions == NULL: evaluated at t from the
 * trajectories).  accelerations[3*body]. */
pb_status_t pb_compute_gravitational_acceleration_on_massive_bodies(
    pb_ephemeris_t* ephemeris, double t, const double* positions,
    double* accelerations);

/* ---- Fixed-step massless flow: Ephemeris::NewInstance + FlowWithFixedStep,
 * ephemeris.hpp:181-194,222-225, with
 * SymplecticRungeKuttaNystromIntegrator::Instance::Solve
 * (integrators/symplectic_runge_kutta_nyström_integrator_body.hpp:24-144).
 *
 * `state` is the integrator instance's ODE::State
 * (integrators/ordinary_differential_equations.hpp:156-180) for n probes
 * sharing one time, SoA of 12 planes of n doubles:
 *   planes 0-2 q.value xyz, 3-5 q.error, 6-8 v.value, 9-11 v.error.
 * `time` is the DoublePrecision<Instant> {value, error}; both are updated in
 * place so that the call is restartable like Instance::Solve.
 * Steps are taken while |step| <= |(t_final - time.value) - time.error|.
 * Trajectory sink (AppendMasslessBodiesStateToTrajectories,
 * ephemeris_body.hpp:1118-1131): if `samples` != NULL, every `sample_every`-th
 * step (counted from this call) the time and the 6 value planes are appended:
 * samples[s][6][n], sample_times[s]; at most `max_samples`.
 * *n_steps receives the number of steps taken; *status the first non-OK RHS
 * status (PB_OUT_OF_RANGE on collision; integration continues, as in the
 * reference). */
typedef struct pb_fixed_step_flow_t {
  int32_t method;       /* An SRKN method, or PB_QUINLAN_1999_ORDER_8A /
                         * PB_QUINLAN_TREMAINE_1990_ORDER_12 (step > 0): the
                         * instance, with the history it would keep, then
                         * lives for this call only (the startup runs again
                         * at the next call); use pb_fixed_step_instance_*
                         * to keep it. */
  double step;
  double t_final;
  int64_t n;
  double* state;        /* 12 * n */
  double* time;         /* 2 */
  int64_t sample_every; /* 0: no sampling. */
  int64_t max_samples;
  double* samples;      /* max_samples * 6 * n, or NULL. */
  double* sample_times; /* max_samples, HOST memory in both variants. */
  int64_t* n_samples;   /* out, may be NULL. */
  int64_t* n_steps;     /* out, may be NULL. */
  pb_status_t* status;  /* out, may be NULL. */
  /* The IntrinsicAccelerations of NewInstance (ephemeris.hpp:181-194, added to
   * the gravitational accelerations, ephemeris_body.hpp:376-382) in tabulated
   * form: NULL, or an inertially fixed burn per probe, 8 planes of n laid out
   * as in pb_adaptive_step_flow_t.burns. */
  const double* burns;
  /* PB_ARITHMETIC_EXACT (0, the default of a zero-initialised struct) or
   * PB_ARITHMETIC_CONTRACTED. */
  int32_t arithmetic;
} pb_fixed_step_flow_t;
pb_status_t pb_flow_with_fixed_step(pb_ephemeris_t* ephemeris,
                                    const pb_fixed_step_flow_t* flow);
/* state/samples in device memory; time, sample_times and outs on the host. */
pb_status_t pb_flow_with_fixed_step_device(pb_ephemeris_t* ephemeris,
                                           const pb_fixed_step_flow_t* flow,
                                           void* stream);

/* ---- Ephemeris::NewInstance (ephemeris.hpp:181-194) + FlowWithFixedStep
 * (:222-225) with a persistent integrator instance over n massless
 * trajectories: Integrator<NewtonianMotionEquation>::Instance
 * (integrators/integrators.hpp:40-94).  The ODE::State lives on the device;
 * for the symmetric linear multistep methods
 * (integrators/symmetric_linear_multistep_integrator_body.hpp:22-152; the
 * plugin's history integrator is Quinlan1999Order8A at 10 s,
 * ksp_plugin/integrators.cpp:66-72) the instance keeps previous_steps_, is
 * started by BlanesMoan2002SRKN14A at step / 16 (starter_body.hpp:23-93) and
 * computes velocities with Cohen-Hubbard-Oesterwinter (:281-321).  `state`
 * and `time` as in pb_fixed_step_flow_t. */
typedef struct pb_fixed_step_instance pb_fixed_step_instance_t;
pb_status_t pb_fixed_step_instance_create(pb_ephemeris_t* ephemeris,
                                          int32_t method, double step,
                                          int64_t n, const double* state,
                                          const double* time,
                                          pb_fixed_step_instance_t** instance);
void pb_fixed_step_instance_destroy(pb_fixed_step_instance_t* instance);
/* The intrinsic_accelerations of NewInstance, tabulated: 8 planes of n as in
 * pb_fixed_step_flow_t.burns (host memory, copied), or NULL for none. */
pb_status_t pb_fixed_step_instance_set_intrinsic_accelerations(
    pb_fixed_step_instance_t* instance, const double* burns);
/* PB_ARITHMETIC_EXACT (default) or PB_ARITHMETIC_CONTRACTED for the flows of
 * this instance. */
pb_status_t pb_fixed_step_instance_set_arithmetic(
    pb_fixed_step_instance_t* instance, int32_t arithmetic);
/* FlowWithFixedStep(t_final, instance) = Instance::Solve(t_final).  Every
 * `sample_every`-th state appended by this call goes to samples[s][6][n]
 * (host), sample_times[s]; *n_steps = states appended; *flow_status as in
 * pb_fixed_step_flow_t.  Outs may be NULL. */
pb_status_t pb_fixed_step_instance_flow(pb_fixed_step_instance_t* instance,
                                        double t_final, int64_t sample_every,
                                        int64_t max_samples, double* samples,
                                        double* sample_times,
                                        int64_t* n_steps, int64_t* n_samples,
                                        pb_status_t* flow_status);
/* FlowWithFixedStep(t_final, instance) with a trajectory sink on the device:
 * every appended state goes to `downsampler` (pb_downsampler_*, below) in
 * chunks, nothing travels to the host.  Append the initial state first if the
 * timelines are to start with it. */
typedef struct pb_downsampler pb_downsampler_t;
pb_status_t pb_fixed_step_instance_flow_to_downsampler(
    pb_fixed_step_instance_t* instance, double t_final,
    pb_downsampler_t* downsampler, int64_t* n_steps,
    pb_status_t* flow_status);
/* Instance::WriteToMessage / ReadFromMessage (integrators/integrators.hpp:
 * 66-78) as a flat array of doubles: the state, the time, the intrinsic
 * accelerations and, for a multistep instance, its previous steps and the
 * progress of its startup.  A restored instance continues bit for bit like the
 * original.  `checkpoint` has room for ..._checkpoint_size() doubles. */
int64_t pb_fixed_step_instance_checkpoint_size(
    const pb_fixed_step_instance_t* instance);
pb_status_t pb_fixed_step_instance_write_checkpoint(
    pb_fixed_step_instance_t* instance, double* checkpoint);
pb_status_t pb_fixed_step_instance_read_checkpoint(
    pb_ephemeris_t* ephemeris, const double* checkpoint, int64_t size,
    pb_fixed_step_instance_t** instance);
/* instance.state(): 12 planes of n and the time {value, error}; either may be
 * NULL. */
pb_status_t pb_fixed_step_instance_get_state(
    pb_fixed_step_instance_t* instance, double* state, double* time);

/* ---- Adaptive massless flow: Ephemeris::FlowWithAdaptiveStep,
 * ephemeris.hpp:201-207, ephemeris_body.hpp:415-444,1624-1717, with
 * EmbeddedExplicitRungeKuttaNystromIntegrator::Instance::Solve
 * (integrators/embedded_explicit_runge_kutta_nyström_integrator_body.hpp:
 * 44-258), one independent trajectory per probe, batched.
 * `state` as above (12 planes); `time` holds per-probe DoublePrecision times:
 * planes time.value[n], time.error[n].  first_step = t_final - time.value,
 * safety factor 0.9, last step exact.  Per-probe outputs (n entries each, may
 * be NULL): status (PB_OK, PB_FAILED_PRECONDITION, PB_RESOURCE_EXHAUSTED;
 * collisions are swallowed as in the reference), accepted steps, rejected
 * attempts, RHS evaluations.  If `step_log` != NULL the times of the first
 * `step_log_capacity` accepted steps of each probe are written to
 * step_log[k * n + probe] (parity of the accepted-step sequence). */
typedef struct pb_adaptive_step_flow_t {
  int32_t method;       /* PB_DORMAND_ELMIKKAWY_PRINCE_1986_RKN_434FM, or
                         * PB_FINE_1987_RKNG_34 (the generalized overload,
                         * ephemeris_body.hpp:446-478, with a velocity-
                         * independent intrinsic acceleration). */
  double t_final;
  double length_integration_tolerance;
  double speed_integration_tolerance;
  int64_t max_steps;
  int64_t n;
  double* state;        /* 12 * n */
  double* time;         /* 2 * n */
  int32_t* status;      /* n */
  int64_t* accepted_steps;
  int64_t* rejected_steps;
  int64_t* evaluations;
  int64_t step_log_capacity;
  double* step_log;     /* step_log_capacity * n */
  /* If != NULL, the q and v values of the first step_log_capacity accepted
   * states of each probe (what append_state hands to
   * DiscreteTrajectory::Append): state_log[k][6][n]. */
  double* state_log;    /* step_log_capacity * 6 * n */
  /* If != NULL, every accepted state of probe i is appended to timeline i of
   * this trajectory sink (pb_downsampler_*, n trajectories) inside the kernel:
   * the DiscreteTrajectory the reference flows. */
  struct pb_downsampler* sink;
  /* The IntrinsicAcceleration of the reference (a host function of t,
   * ephemeris.hpp:72-74, added to the gravitational acceleration,
   * ephemeris_body.hpp:430-432) in tabulated form: if != NULL, probe i fires an
   * inertially fixed burn, Manoeuvre::ComputeIntrinsicAcceleration
   * (ksp_plugin/manœuvre_body.hpp:395-406):
   *   direction * thrust / (initial_mass - (t - initial_time) * mass_flow)
   * for initial_time <= t <= final_time, the zero vector elsewhere.
   * 8 planes of n: direction xyz, thrust, initial_mass, mass_flow,
   * initial_time, final_time (SI). */
  const double* burns;
  /* The GeneralizedIntrinsicAcceleration of the generalized overload
   * (ephemeris.hpp:209-215, ephemeris_body.hpp:446-478; needs
   * PB_FINE_1987_RKNG_34, whose velocity stages a' then feed the right-hand
   * side): if != NULL, n body indices (caller order; in device memory for the
   * _device variant); probe i with frenet_centres[i] >= 0 fires a burn whose
   * `direction` is expressed in the Frenet frame (tangent, normal, binormal)
   * of its trajectory in the body-centred non-rotating frame of that body,
   * Manoeuvre::FrenetIntrinsicAcceleration (ksp_plugin/manœuvre_body.hpp:
   * 310-321,383-393; physics/reference_frame_body.hpp:38-54): tangent along
   * the velocity relative to the body, normal along the part of the geometric
   * acceleration (gravitational acceleration of the probe minus that of the
   * body) orthogonal to it.  -1: inertially fixed as above. */
  const int32_t* frenet_centres;
} pb_adaptive_step_flow_t;
enum { PB_BURN_PLANES = 8 };
pb_status_t pb_flow_with_adaptive_step(pb_ephemeris_t* ephemeris,
                                       const pb_adaptive_step_flow_t* flow);
pb_status_t pb_flow_with_adaptive_step_device(
    pb_ephemeris_t* ephemeris, const pb_adaptive_step_flow_t* flow,
    void* stream);

/* ---- Massive-body flow (Ephemeris::Prolong's integration,
 * ephemeris_body.hpp:116-170,312-344,1503-1552) for ensembles of independent
 * small systems sharing one gravity model (astronomy/trappist_dynamics_test.cpp
 * :275-297: one Ephemeris per ensemble member), with
 * SymmetricLinearMultistepIntegrator (…multistep_integrator_body.hpp:22-152),
 * its SRKN14A starter (starter_body.hpp:23-93) and the
 * Cohen-Hubbard-Oesterwinter velocities (:281-321), or plain SRKN.
 *
 * The ensemble uses the ephemeris' bodies (internal order, masses and
 * geopotentials; trajectories are not used).  `state`: 12 planes as above,
 * each plane [body (internal order)][system]: index (plane * n_bodies + body)
 * * n_systems + system.  `time` {value, error} is shared.  Runs the startup if
 * `history` is not yet started, then steps while
 * step <= (t_final - time.value) - time.error. */
typedef struct pb_nbody_ensemble pb_nbody_ensemble_t;
pb_status_t pb_nbody_ensemble_create(pb_ephemeris_t* ephemeris,
                                     int32_t method, double step,
                                     int64_t n_systems,
                                     const double* state, /* 12*n_bodies*n_sys */
                                     double initial_time,
                                     pb_nbody_ensemble_t** ensemble);
void pb_nbody_ensemble_destroy(pb_nbody_ensemble_t* ensemble);
/* Instance::Solve(t_final); *n_steps = steps appended by this call. */
pb_status_t pb_nbody_ensemble_solve(pb_nbody_ensemble_t* ensemble,
                                    double t_final, int64_t* n_steps);
/* Same, with the instance's append_state callback (integrators.hpp:60-66)
 * materialised: the time and the 6 value planes (q, v) of every appended state
 * are written to times[r], states[r][6][n_bodies][n_systems] (host).  Returns
 * once `capacity` states have been recorded or t_final is reached; call again
 * to continue.  This is how the facade's Prolong feeds
 * ContinuousTrajectory::Append (ephemeris_body.hpp:1071-1116). */
pb_status_t pb_nbody_ensemble_solve_recording(pb_nbody_ensemble_t* ensemble,
                                              double t_final, int64_t capacity,
                                              double* times, double* states,
                                              int64_t* n_recorded);
/* Current ODE::State (host copy): state 12*n_bodies*n_systems, time[2]. */
pb_status_t pb_nbody_ensemble_get_state(pb_nbody_ensemble_t* ensemble,
                                        double* state, double* time);

/* ---- ContinuousTrajectory production on the device.
 * A batch of n ContinuousTrajectory objects appended in lockstep with equally
 * spaced degrees of freedom -- the bodies of one ephemeris, or of every member
 * of an ensemble: ContinuousTrajectory::Append with
 * ComputeBestNewhallApproximation (physics/continuous_trajectory_body.hpp
 * :82-127,758-858) and NewhallApproximationInMonomialBasis
 * (numerics/newhall_body.hpp:52-117,259-287) run in one kernel every 8 points,
 * one thread per trajectory; the polynomials stay in device memory
 * ([segment][xyz][18][n], sized for `max_segments`) and are evaluated there
 * (EvaluatePosition / EvaluateVelocity, :686-722).
 * `step` and `tolerance` are the constructor's (:60-80). */
typedef struct pb_continuous_trajectories pb_continuous_trajectories_t;
pb_status_t pb_continuous_trajectories_create(
    int32_t device, int64_t n_trajectories, double step, double tolerance,
    int64_t max_segments, pb_continuous_trajectories_t** trajectories);
void pb_continuous_trajectories_destroy(
    pb_continuous_trajectories_t* trajectories);
/* Append(time, degrees of freedom) for every trajectory; q, v: 3 planes of n.
 * The host variant returns PB_INVALID_ARGUMENT if any trajectory hit the
 * "apocalypse" (:851-857); the device variant is asynchronous on `stream`
 * (q, v in device memory) and leaves that to ..._status. */
pb_status_t pb_continuous_trajectories_append(
    pb_continuous_trajectories_t* trajectories, double time, const double* q,
    const double* v);
pb_status_t pb_continuous_trajectories_append_device(
    pb_continuous_trajectories_t* trajectories, double time, const double* q,
    const double* v, void* stream);
pb_status_t pb_continuous_trajectories_status(
    pb_continuous_trajectories_t* trajectories);
double pb_continuous_trajectories_t_min(
    const pb_continuous_trajectories_t* trajectories);
double pb_continuous_trajectories_t_max(
    const pb_continuous_trajectories_t* trajectories);
int64_t pb_continuous_trajectories_segment_count(
    const pb_continuous_trajectories_t* trajectories);
/* The polynomials of one trajectory, in the layout of
 * pb_ephemeris_append_segments: t_max[s], origin[s], degree[s],
 * coefficients[s][18][3] (zero above the degree); any may be NULL. */
pb_status_t pb_continuous_trajectories_get_segments(
    pb_continuous_trajectories_t* trajectories, int64_t trajectory,
    double* t_min, double* t_max, double* origin, int32_t* degree,
    double* coefficients);
/* EvaluatePosition / EvaluateVelocity of every trajectory at t: 3 planes of n
 * each; either may be NULL. */
pb_status_t pb_continuous_trajectories_evaluate(
    pb_continuous_trajectories_t* trajectories, double t, double* positions,
    double* velocities);
pb_status_t pb_continuous_trajectories_evaluate_device(
    pb_continuous_trajectories_t* trajectories, double t, double* positions,
    double* velocities, void* stream);
/* Instance::Solve(t_final) of the ensemble with append_state =
 * Ephemeris::AppendMassiveBodiesState (ephemeris_body.hpp:1071-1116) on the
 * device: every appended state goes to `trajectories` (n = n_bodies *
 * n_systems, trajectory = body (internal order) * n_systems + system) without
 * leaving the GPU.  If the trajectories are empty the current state is
 * appended first (the Ephemeris constructor, :136). */
pb_status_t pb_nbody_ensemble_solve_to_trajectories(
    pb_nbody_ensemble_t* ensemble, double t_final,
    pb_continuous_trajectories_t* trajectories, int64_t* n_steps);

/* ---- Large-N all-pairs point-mass system (synthetic; the same
 * ComputeGravitationalAccelerationBetweenAllMassiveBodies with only spherical
 * bodies, ephemeris_body.hpp:1342-1380,1503-1552), Quinlan-Tremaine / SRKN
 * (integrators/symmetric_linear_multistep_integrator_body.hpp:22-152).
 * Sharded by target block across ranks: this rank owns targets
 * [i_begin, i_end) of n.  Every force evaluation needs all positions, so after
 * every position update the owned slices are exchanged:
 *  - pb_nbody_allpairs_set_nccl: ONE in-place ncclAllGather on the force
 *    stream, issued by the library (equal blocks in rank order);
 *  - or the `exchange` callback (may be NULL when the rank owns every target):
 *    it receives the device pointer of the source array, [n][4] doubles
 *    {x, y, z, mu} whose rows [i_begin, i_end) this rank has just written, and
 *    must return once every other rank's rows are visible in that array (work
 *    enqueued on `stream` counts).  A non-OK return aborts the solve.
 * A target's acceleration has the same bits whichever rank owns it (the
 * summation order depends on n only, pb_kernels_allpairs.cuh). */
typedef pb_status_t (*pb_exchange_positions_t)(void* user,
                                               double* sources_device,
                                               int64_t n, int64_t i_begin,
                                               int64_t i_end, void* stream);
typedef struct pb_nbody_allpairs pb_nbody_allpairs_t;
pb_status_t pb_nbody_allpairs_create(int32_t cuda_device, int32_t method,
                                     double step, int64_t n, int64_t i_begin,
                                     int64_t i_end,
                                     const double* gravitational_parameters,
                                     const double* q, /* 3 planes of n */
                                     const double* v, /* 3 planes of n */
                                     double initial_time,
                                     pb_exchange_positions_t exchange,
                                     void* exchange_user,
                                     pb_nbody_allpairs_t** system);
void pb_nbody_allpairs_destroy(pb_nbody_allpairs_t* system);
/* `nccl_comm` is the caller's ncclComm_t (not owned; NULL detaches): its size
 * and this rank's index must match the sharding, n = size x (i_end - i_begin),
 * i_begin = rank x (i_end - i_begin).  NCCL is resolved at run time from the
 * copy already loaded in the process (else libnccl.so.2 on the library path). */
pb_status_t pb_nbody_allpairs_set_nccl(pb_nbody_allpairs_t* system,
                                       void* nccl_comm);
/* For hosts that have no communicator of their own: ncclGetUniqueId /
 * ncclCommInitRank / ncclCommDestroy through the same run-time binding.
 * `unique_id` is 128 bytes, created on one rank and handed to the others by the
 * host (the Python mirror broadcasts it with torch.distributed). */
pb_status_t pb_nccl_get_unique_id(char* unique_id);
pb_status_t pb_nccl_comm_create(const char* unique_id, int32_t world,
                                int32_t rank, int32_t cuda_device,
                                void** nccl_comm);
void pb_nccl_comm_destroy(void* nccl_comm);
/* PB_ARITHMETIC_EXACT (default) or PB_ARITHMETIC_CONTRACTED, see the enum. */
pb_status_t pb_nbody_allpairs_set_arithmetic(pb_nbody_allpairs_t* system,
                                             int32_t arithmetic);
pb_status_t pb_nbody_allpairs_solve(pb_nbody_allpairs_t* system, double t_final,
                                    int64_t* n_steps,
                                    int64_t* n_force_evaluations);
/* Owned slice: q, v as 3 planes of (i_end - i_begin) values; time[2]. */
pb_status_t pb_nbody_allpairs_get_state(pb_nbody_allpairs_t* system, double* q,
                                        double* v, double* time);
/* One force evaluation on the current positions (for measurement): the
 * accelerations of the owned targets, 3 planes of (i_end - i_begin). */
pb_status_t pb_nbody_allpairs_accelerations(pb_nbody_allpairs_t* system,
                                            double* a);
/* Milliseconds spent in the force kernels, resp. in the position exchanges,
 * of the last solve / accelerations call (CUDA events on the force stream). */
double pb_nbody_allpairs_last_force_ms(const pb_nbody_allpairs_t* system);
double pb_nbody_allpairs_last_exchange_ms(const pb_nbody_allpairs_t* system);

/* ---- The trajectory sink: DiscreteTrajectorySegment::SetDownsampling +
 * Append (physics/discrete_trajectory_segment_body.hpp:414-479; tolerances in
 * ksp_plugin/integrators.cpp:23-36) for n trajectories on the device, so that
 * the states appended by a flow need not all travel to the host.  A new point
 * replaces the last one of a timeline as long as the accumulated bound on the
 * error of the Hermite interpolation (numerics/hermite3_body.hpp:154-194)
 * stays below `tolerance` (negative: no downsampling).  Each timeline holds at
 * most `capacity` points. */
pb_status_t pb_downsampler_create(int32_t cuda_device, int64_t n,
                                  double tolerance, int64_t capacity,
                                  pb_downsampler_t** downsampler);
void pb_downsampler_destroy(pb_downsampler_t* downsampler);
/* Appends n_samples states, in order, to every trajectory:
 * samples[s][6][n] (q xyz, v xyz: the `samples` of the fixed-step flows) at
 * sample_times[s] (host).  `samples` on the host, or on the device. */
pb_status_t pb_downsampler_append(pb_downsampler_t* downsampler,
                                  int64_t n_samples,
                                  const double* sample_times,
                                  const double* samples);
pb_status_t pb_downsampler_append_device(pb_downsampler_t* downsampler,
                                         int64_t n_samples,
                                         const double* sample_times,
                                         const double* samples, void* stream);
/* Appends the logs of an adaptive flow (device memory, as filled by
 * pb_flow_with_adaptive_step_device): trajectory i gets its first
 * min(accepted_steps[i], log_capacity) states, step_log[s][n],
 * state_log[s][6][n]. */
pb_status_t pb_downsampler_append_logs_device(pb_downsampler_t* downsampler,
                                              int64_t log_capacity,
                                              const double* step_log,
                                              const int64_t* accepted_steps,
                                              const double* state_log,
                                              void* stream);
/* The timelines: counts[n]; times[capacity][n]; points[capacity][6][n];
 * errors[capacity][n] (the downsampling error recorded with each
 * interpolation).  Any may be NULL.  PB_RESOURCE_EXHAUSTED if a timeline
 * outgrew `capacity` (its later points were dropped). */
pb_status_t pb_downsampler_get(pb_downsampler_t* downsampler, int64_t* counts,
                               double* times, double* points, double* errors);
/* DiscreteTrajectorySegment::EvaluatePosition / EvaluateVelocity
 * (physics/discrete_trajectory_segment_body.hpp:132-155) of every timeline at
 * t: the point at t if there is one, else the Hermite interpolation between
 * its neighbours (numerics/hermite3_body.hpp).  q, v: 3 planes of n (either may
 * be NULL); in_range[n] (may be NULL) is 0 where the timeline does not cover t
 * (the reference CHECKs), and the outputs are NaN there. */
pb_status_t pb_downsampler_evaluate(pb_downsampler_t* downsampler, double t,
                                    double* q, double* v, int32_t* in_range);
pb_status_t pb_downsampler_evaluate_device(pb_downsampler_t* downsampler,
                                           double t, double* q, double* v,
                                           int32_t* in_range, void* stream);

/* ---- Series in the Чебышёв basis.
 * PolynomialInЧебышёвBasis<Value, Instant, degree>::operator() and
 * EvaluateDerivative, numerics/polynomial_in_чебышёв_basis_body.hpp:151-202
 * (Clenshaw; the representation of pre-Cardano ContinuousTrajectory segments),
 * batched: series i has `dimension` components, degree[i] <=
 * PB_MAX_CHEBYSHEV_DEGREE and bounds [lower_bound[i], upper_bound[i]];
 * coefficients[i][k][c] with PB_MAX_CHEBYSHEV_DEGREE + 1 slots of k.
 * Evaluation j is of series series_index[j] at arguments[j].  values and
 * derivatives are [n_evaluations][dimension]; either may be NULL. */
#define PB_MAX_CHEBYSHEV_DEGREE 20
pb_status_t pb_chebyshev_series_evaluate(
    int32_t cuda_device, int64_t n_series, int32_t dimension,
    const int32_t* degree, const double* lower_bound,
    const double* upper_bound, const double* coefficients,
    int64_t n_evaluations, const int32_t* series_index,
    const double* arguments, double* values, double* derivatives);

/* ---- Measurement helpers. */
/* Device time (CUDA events on the launching stream) spent in the flow kernels
 * of the last pb_flow_with_*_step[_device] call on this ephemeris, and their
 * number. */
double pb_ephemeris_last_flow_kernel_ms(const pb_ephemeris_t* ephemeris,
                                        int64_t* launches);
/* Which code path the last fixed-step flow on this ephemeris took.  *hot_body:
 * the caller index of the oblate body whose harmonics were served from the
 * constant bank (-1: none) -- the body whose harmonics of degree >= 3 the most
 * probes of the first flow needed (every probe votes), kept for the later
 * flows; a second, "warm" body (harmonics up to degree 3, e.g. the damped J2 of
 * the Sun) is chosen the same way.  *generic_harmonics != 0: some probe
 * evaluated harmonics that neither serves (same bits, several times slower),
 * in which case the next flow chooses afresh. */
pb_status_t pb_ephemeris_last_flow_path(const pb_ephemeris_t* ephemeris,
                                        int32_t* hot_body,
                                        int32_t* generic_harmonics);
/* Register-resident DFMA chain: returns the achieved FP64 FLOP/s on the
 * device (2 flop per FMA) and the kernel time through *ms. */
pb_status_t pb_measure_fp64_peak(int32_t cuda_device, double* flops_per_second,
                                 double* ms);
/* The SM clock (MHz) of the device under its current load, measured ON the
 * device: one warp reads the cycle counter against the nanosecond global timer
 * for `microseconds`.  For benchmarks: the `clocks.sm` that nvidia-smi / NVML
 * report (/opt/skills/guides/B200_PROFILING.md), without a driver query in
 * the timed region.  Any thread; runs on a stream of its own. */
pb_status_t pb_measure_sm_clock(int32_t cuda_device, double microseconds,
                                double* mhz);
/* The same in two halves, for a caller that wants the measurement to overlap
 * work it is about to launch from the same thread: _start enqueues the probe
 * (it waits `delay_microseconds`, then measures over `microseconds`) and
 * returns at once, _read waits for it and returns its figure.  One probe in
 * flight per thread. */
pb_status_t pb_sm_clock_probe_start(int32_t cuda_device,
                                    double delay_microseconds,
                                    double microseconds);
pb_status_t pb_sm_clock_probe_read(int32_t cuda_device, double* mhz);
/* Kernels launched by this library since load (all entry points). */
int64_t pb_kernel_launch_count(void);

#ifdef __cplusplus
}  /* extern "C" */
#endif

#endif  /* PRINCIPIA_B200_H_ */
