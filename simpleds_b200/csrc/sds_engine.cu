// Fused engine (placeholder).
#include "sds_common.cuh"
static thread_local int g_last_launches = 0;
extern "C" int64_t sds_engine_workspace_bytes(const sds_plan *, const sds_engine_desc *) { return 0; }
extern "C" int sds_engine_run(const sds_plan *, const sds_engine_desc *, const void *,
                              const uint8_t *, const float *, const float *, const int32_t *,
                              const float *, const double *, const double *, const double *,
                              const void *, double *, double *, double *, double *, double *,
                              double *, void *, sds_stream) {
  sds::set_error("sds_engine_run: not built yet");
  return SDS_ENOTSUP;
}
extern "C" int sds_engine_last_launches(void) { return g_last_launches; }
