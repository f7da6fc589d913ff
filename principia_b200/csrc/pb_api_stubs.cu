// Entry points whose kernels are not written yet: they fail loudly.
#include "../../include/principia_b200.h"
#include "pb_runtime.cuh"

using namespace pb;

extern "C" {

pb_status_t pb_flow_with_adaptive_step(pb_ephemeris_t*,
                                       const pb_adaptive_step_flow_t*) {
  return Fail(PB_INTERNAL, "pb_flow_with_adaptive_step: not implemented");
}
pb_status_t pb_flow_with_adaptive_step_device(pb_ephemeris_t*,
                                              const pb_adaptive_step_flow_t*,
                                              void*) {
  return Fail(PB_INTERNAL, "pb_flow_with_adaptive_step_device: not implemented");
}
pb_status_t pb_nbody_ensemble_create(pb_ephemeris_t*, int32_t, double, int64_t,
                                     const double*, double,
                                     pb_nbody_ensemble_t**) {
  return Fail(PB_INTERNAL, "pb_nbody_ensemble_create: not implemented");
}
void pb_nbody_ensemble_destroy(pb_nbody_ensemble_t*) {}
pb_status_t pb_nbody_ensemble_solve(pb_nbody_ensemble_t*, double, int64_t*) {
  return Fail(PB_INTERNAL, "pb_nbody_ensemble_solve: not implemented");
}
pb_status_t pb_nbody_ensemble_get_state(pb_nbody_ensemble_t*, double*,
                                        double*) {
  return Fail(PB_INTERNAL, "pb_nbody_ensemble_get_state: not implemented");
}
pb_status_t pb_nbody_allpairs_create(int32_t, int32_t, double, int64_t, int64_t,
                                     int64_t, const double*, const double*,
                                     const double*, double,
                                     pb_exchange_positions_t, void*,
                                     pb_nbody_allpairs_t**) {
  return Fail(PB_INTERNAL, "pb_nbody_allpairs_create: not implemented");
}
void pb_nbody_allpairs_destroy(pb_nbody_allpairs_t*) {}
pb_status_t pb_nbody_allpairs_solve(pb_nbody_allpairs_t*, double, int64_t*,
                                    int64_t*) {
  return Fail(PB_INTERNAL, "pb_nbody_allpairs_solve: not implemented");
}
pb_status_t pb_nbody_allpairs_get_state(pb_nbody_allpairs_t*, double*, double*,
                                        double*) {
  return Fail(PB_INTERNAL, "pb_nbody_allpairs_get_state: not implemented");
}
pb_status_t pb_nbody_allpairs_accelerations(pb_nbody_allpairs_t*, double*) {
  return Fail(PB_INTERNAL, "pb_nbody_allpairs_accelerations: not implemented");
}
double pb_nbody_allpairs_last_force_ms(const pb_nbody_allpairs_t*) { return 0; }

}  // extern "C"
