// Parts of the PSN_PREC_TC path that still run on the fp32 FFMA kernels (explicit forwarding, never a CPU path).
// Replaced one by one by tcgen05 kernels: see DESIGN.md "Kernels" for the current split.
#include "internal.cuh"
extern "C" int psn_has_tensor_path(void) { return 1; }
namespace psn {
int s2_visibility_simt(const psn_mlp* vis_net, int nf, const float* pts, long long Ns, const float* lights, int L, float* vis,
                       cudaStream_t st);
int tc_infer_occ(const psn_mlp* geo, const PointGen& gen, long long M, float* out, cudaStream_t st) {
  return simt_occupancy(geo, gen, M, nullptr, PSN_OUT_LOGIT, out, 1, st);  // full 257-wide query: API completeness only
}
}  // namespace psn
