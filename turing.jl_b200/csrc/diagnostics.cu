// diagnostics.cu -- placeholder
#include "engine.cuh"
namespace tnuts {
int diagnostics_run(Engine *e, double *, double *) { return set_error(e, TNUTS_E_UNSUPPORTED, "diagnostics not built yet"); }
void diagnostics_destroy(Engine *) {}
}
