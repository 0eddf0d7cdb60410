// comm.cu -- placeholder
#include "engine.cuh"
extern "C" {
int tnuts_comm_unique_id(void *) { return TNUTS_E_UNSUPPORTED; }
int tnuts_comm_init(tnuts_engine *h, const void *, int32_t, int32_t) { return tnuts::set_error((tnuts::Engine *)h, TNUTS_E_UNSUPPORTED, "comm not built yet"); }
}
