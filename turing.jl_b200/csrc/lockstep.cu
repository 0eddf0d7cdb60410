// lockstep.cu -- lock-step strategy (placeholder until implemented)
#include "engine.cuh"
namespace tnuts {
int lockstep_create(Engine *e) { return set_error(e, TNUTS_E_UNSUPPORTED, "lock-step strategy not built yet"); }
void lockstep_destroy(Engine *) {}
int lockstep_init(Engine *e, const double *) { return set_error(e, TNUTS_E_UNSUPPORTED, "lock-step"); }
int lockstep_run(Engine *e) { return set_error(e, TNUTS_E_UNSUPPORTED, "lock-step"); }
int lockstep_point_lpgrad(Engine *e, const double *, long long, double *, double *) { return set_error(e, TNUTS_E_UNSUPPORTED, "lock-step"); }
int lockstep_point_constrain(Engine *e, const double *, long long, double *, double *) { return set_error(e, TNUTS_E_UNSUPPORTED, "lock-step"); }
int lockstep_point_leapfrog(Engine *e, const double *, const double *, const double *, double, int, long long, double *, double *, double *) { return set_error(e, TNUTS_E_UNSUPPORTED, "lock-step"); }
}
