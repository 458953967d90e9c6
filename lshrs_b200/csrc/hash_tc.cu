// placeholder until the tcgen05 kernel lands
#include "lshx_common.cuh"
namespace lshx {
struct TcPlan {};
bool tc_shape_supported(const HashShape&) { return false; }
int tc_plan_create(const HashShape&, const float*, TcPlan** out) { *out = nullptr; return LSHX_OK; }
void tc_plan_destroy(TcPlan*) {}
int launch_hash_tc(const HashShape&, TcPlan*, const float*, int64_t, uint8_t*, uint8_t*, cudaStream_t) {
  set_error("tcgen05 kernel not built");
  return LSHX_ERR_INVALID_ARG;
}
}  // namespace lshx
