// tcgen05 engine (placeholder until the fused kernel lands).
#include "kernels.cuh"
namespace ciaosr {
bool tc_shapes_ok(const ciaosr_head_desc*) { return false; }
size_t tc_blob_bytes(const ciaosr_head_desc*) { return 0; }
int tc_pack(const ciaosr_head_desc*, const PlanLayout&, float*, cudaStream_t) { return CIAOSR_OK; }
size_t head_tc_workspace(const PlanLayout&, int, int, int, int) { return 0; }
int run_head_tc(const PlanLayout&, const float*, const HeadArgs&, void*, size_t, cudaStream_t) {
  set_error("tcgen05 engine not built");
  return CIAOSR_E_INVALID;
}
}  // namespace ciaosr
