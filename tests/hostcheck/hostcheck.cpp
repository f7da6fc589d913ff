// TEST HARNESS (not shipped, not a fallback): compiles the product's
// host/device headers for the HOST with g++ -ffp-contract=off so that the
// arithmetic of the device functions can be compared bit for bit with the
// oracle here, on the CPU-only build box, before GPU time is spent.  The
// product library never links this file.
#include <cstring>
#include <vector>

#include "../../principia_b200/csrc/pb_accel.cuh"
#include "../../principia_b200/csrc/pb_downsampling.cuh"
#include "../../principia_b200/csrc/pb_host_model.hpp"

using namespace pb;

struct hc_model {
  HostModel model;
  std::vector<double> normalization;
  std::vector<int32_t> segment_offset;
  EphemerisView view;
};

extern "C" {

void hc_cr_sincos(double x, double* s, double* c) { CrSinCos(x, *s, *c); }
double hc_cr_root4(double x) { return CrRoot4(x); }
double hc_estrin(double const* c, int32_t degree, double x) {
  return EstrinEvaluate<1>(c, degree, x);
}
double hc_estrin_derivative(double const* c, int32_t degree, double x) {
  return EstrinEvaluateDerivative<1>(c, degree, x);
}
double hc_pow_int(double x, int32_t k) { return PowInt(x, k); }

// DownsamplerAppend on one trajectory (q, v: [n_points][3]); returns the size
// of the timeline, the retained points in kept_* (room for n_points).
int64_t hc_downsample(int64_t n_points, double const* times, double const* q,
                      double const* v, double tolerance, double* kept_times,
                      double* kept_q, double* kept_v, double* kept_errors) {
  std::vector<double> timeline_times(n_points), points(6 * n_points),
      errors(n_points), a2a3(6);
  long long count = 0;
  double downsampling_error = 0;
  int32_t overflow = 0;
  DownsamplerView const view{1,
                             n_points,
                             tolerance,
                             timeline_times.data(),
                             points.data(),
                             errors.data(),
                             &count,
                             &downsampling_error,
                             a2a3.data(),
                             &overflow};
  for (int64_t i = 0; i < n_points; ++i) {
    DownsamplerAppend(view, 0, times[i],
                      Vec3{q[3 * i], q[3 * i + 1], q[3 * i + 2]},
                      Vec3{v[3 * i], v[3 * i + 1], v[3 * i + 2]});
  }
  for (long long k = 0; k < count; ++k) {
    kept_times[k] = timeline_times[k];
    for (int c = 0; c < 3; ++c) {
      kept_q[3 * k + c] = points[6 * k + c];
      kept_v[3 * k + c] = points[6 * k + 3 + c];
    }
    kept_errors[k] = errors[k];
  }
  return overflow != 0 ? -1 : count;
}
void hc_chebyshev(double const* c, int32_t degree, double lower, double upper,
                  double t, double* value, double* derivative) {
  double const one_over_width = 1 / (upper - lower);
  *value = ChebyshevEvaluate<1>(c, degree, lower, upper, one_over_width, t);
  *derivative =
      ChebyshevEvaluateDerivative<1>(c, degree, lower, upper, one_over_width, t);
}

// One segment [18][3] of the given degree, evaluated by the run-time tree and by
// the padded tree (whose slots above the degree must hold -0.0).
void hc_estrin_padded(double const* c, int32_t degree, double origin, double t,
                      double* reference, double* padded) {
  EphemerisView view{};
  int32_t segment_degree = degree;
  view.segment_origin = &origin;
  view.segment_degree = &segment_degree;
  view.segment_coefficients = c;
  Vec3 const a = EvaluateSegment(view, 0, t);
  Vec3 const b = EvaluateSegmentPadded(view, 0, t);
  reference[0] = a.x; reference[1] = a.y; reference[2] = a.z;
  padded[0] = b.x; padded[1] = b.y; padded[2] = b.z;
}

int32_t hc_model_create(int32_t n, pb_body_t const* bodies, double tolerance,
                        hc_model** out) {
  auto* m = new hc_model;
  int32_t const status = BuildHostModel(n, bodies, tolerance, m->model);
  if (status != PB_OK) {
    delete m;
    return status;
  }
  m->normalization.assign(pb_legendre_normalization_factor,
                          pb_legendre_normalization_factor + 51 * 52 / 2);
  m->segment_offset.assign(n + 1, 0);
  EphemerisView& v = m->view;
  std::memset(&v, 0, sizeof(v));
  v.n_bodies = n;
  v.n_oblate = m->model.n_oblate;
  v.n_frames = m->model.n_frames;
  v.stage_stride = StageStride(n, v.n_frames);
  v.bodies = m->model.bodies.data();
  v.dampings = m->model.dampings.data();
  v.cos = m->model.cos.data();
  v.sin = m->model.sin.data();
  v.normalization = m->normalization.data();
  *out = m;
  return PB_OK;
}
void hc_model_destroy(hc_model* m) { delete m; }

int32_t hc_model_order(hc_model const* m, int32_t* order, int32_t* n_oblate) {
  std::memcpy(order, m->model.order.data(),
              m->model.order.size() * sizeof(int32_t));
  *n_oblate = m->model.n_oblate;
  return PB_OK;
}

int32_t hc_model_damping(hc_model const* m, int32_t body, int32_t* n_degrees,
                         double* inner, double* sectoral_inner) {
  DeviceBody const& b = m->model.bodies[m->model.internal_of_caller[body]];
  if (!b.is_oblate) {
    return PB_INVALID_ARGUMENT;
  }
  *n_degrees = b.n_damping;
  for (int i = 0; i < b.n_damping; ++i) {
    inner[i] = m->model.dampings[b.damping_offset + i].inner_threshold;
  }
  *sectoral_inner = b.sectoral.inner_threshold;
  return PB_OK;
}

// Geopotential acceleration of caller body `body` on displacement r at time t.
int32_t hc_geopotential_acceleration(hc_model const* m, int32_t body, double t,
                                     double const* r, int32_t unrolled,
                                     double* out) {
  int const b = m->model.internal_of_caller[body];
  GeopotentialBody const g = MakeGeopotentialBody(m->view, b);
  Vec3 x_hat, y_hat;
  double frame[6] = {};
  if (g.body->frame_slot >= 0) {
    SurfaceFrameAxes(*g.body, t, x_hat, y_hat);
    frame[0] = x_hat.x; frame[1] = x_hat.y; frame[2] = x_hat.z;
    frame[3] = y_hat.x; frame[4] = y_hat.y; frame[5] = y_hat.z;
  }
  Vec3 const rv{r[0], r[1], r[2]};
  double const r2 = Norm2(rv);
  double const r_norm = sqrt(r2);
  double const one_over_r3 = r_norm / (r2 * r2);
  if (unrolled == 3) {
    // The capped unrolled evaluation of the adaptive flow's deep body
    // (pb_kernels_adaptive.cuh AdaptiveHarmonics): the body of degree
    // kMaxUnrolledDegree left at the limiting degree of this distance.
    int const max_degree =
        LimitingDegree(g.degree_damping, g.body->n_damping, r_norm) - 1;
    Vec3 a{0.0, 0.0, 0.0};
    if (max_degree > kMaxUnrolledDegree) {
      return PB_INVALID_ARGUMENT;
    }
    if (max_degree != 1) {
      bool const is_zonal =
          g.body->is_zonal || r_norm > g.body->sectoral.outer_threshold;
      if (is_zonal) {
        x_hat = g.body->equatorial;
        y_hat = g.body->biequatorial;
      } else {
        SurfaceFrameAxes(*g.body, t, x_hat, y_hat);
      }
      GeopotentialSetup s;
      InitializeSetup(*g.body, x_hat, y_hat, rv, r_norm, r2, one_over_r3, s);
      PointerTables const tables{g};
      LocalScratch<kMaxUnrolledDegree> w;
      a = is_zonal ? GeopotentialUnrolledAcceleration<kMaxUnrolledDegree, true,
                                                      true>(tables, s, w,
                                                            max_degree)
                   : GeopotentialUnrolledAcceleration<kMaxUnrolledDegree, false,
                                                      true>(tables, s, w,
                                                            max_degree);
    }
    out[0] = a.x; out[1] = a.y; out[2] = a.z;
    return PB_OK;
  }
  Vec3 const a =
      unrolled == 2
          ? GeneralSphericalHarmonicsAccelerationRolled(g, t, rv, r_norm, r2,
                                                        one_over_r3)
          : unrolled
                ? GeneralSphericalHarmonicsAcceleration<true>(
                      g, FrameSource{frame, 0.0}, rv, r_norm, r2, one_over_r3)
                : GeneralSphericalHarmonicsAcceleration<false>(
                      g, FrameSource{nullptr, t}, rv, r_norm, r2,
                      one_over_r3);
  out[0] = a.x; out[1] = a.y; out[2] = a.z;
  return PB_OK;
}

// Massless acceleration + potential given the massive positions (caller order)
// at time t (frames computed here).
int32_t hc_acceleration_on_massless_body(hc_model const* m, double t,
                                         double const* positions,
                                         double const* q, int32_t unrolled,
                                         double* a, double* potential) {
  EphemerisView const& v = m->view;
  std::vector<double> stage(v.stage_stride);
  for (int b = 0; b < v.n_bodies; ++b) {
    int const caller = m->model.order[b];
    for (int c = 0; c < 3; ++c) {
      stage[3 * b + c] = positions[3 * caller + c];
    }
    if (v.bodies[b].frame_slot >= 0) {
      Vec3 x_hat, y_hat;
      SurfaceFrameAxes(v.bodies[b], t, x_hat, y_hat);
      double* frame = stage.data() + 3 * v.n_bodies + 6 * v.bodies[b].frame_slot;
      frame[0] = x_hat.x; frame[1] = x_hat.y; frame[2] = x_hat.z;
      frame[3] = y_hat.x; frame[4] = y_hat.y; frame[5] = y_hat.z;
    }
  }
  Vec3 acc;
  Vec3 const qv{q[0], q[1], q[2]};
  int32_t const error =
      unrolled ? AccelerationOnMasslessBody<true>(v, stage.data(), qv, acc)
               : AccelerationOnMasslessBody<false>(v, stage.data(), qv, acc);
  a[0] = acc.x; a[1] = acc.y; a[2] = acc.z;
  *potential = PotentialOnMasslessBody(v, stage.data(), qv);
  return error;
}

}  // extern "C"
