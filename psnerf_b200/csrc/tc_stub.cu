// Placeholder for the tcgen05 path while it is being brought up: every entry reports "unavailable".
#include "internal.cuh"
extern "C" int psn_has_tensor_path(void) { return 0; }
namespace psn {
static int na() { set_error("tensor-core (PSN_PREC_TC) path is not available in this build"); return PSN_ERR_SHAPE; }
int tc_pack_bytes(const psn_mlp*) { return 0; }
int tc_pack_fill(psn_mlp*, const float* const*, const float* const*, char*, size_t, cudaStream_t) { return PSN_ERR_SHAPE; }
int tc_occupancy(const psn_mlp*, const PointGen&, long long, const int*, int, float*, cudaStream_t) { return na(); }
int tc_infer_occ(const psn_mlp*, const PointGen&, long long, float*, cudaStream_t) { return na(); }
int tc_gradient(const psn_mlp*, const PointGen&, long long, const int*, float*, void*, cudaStream_t) { return na(); }
int tc_radiance(const psn_mlp*, const psn_mlp*, const PointGen&, long long, float*, float*, void*, cudaStream_t) { return na(); }
int tc_shadow(const psn_mlp*, const PointGen&, long long, float, float*, cudaStream_t) { return na(); }
size_t tc_stash_bytes() { return 0; }
int tc_s2_visibility(const psn_mlp*, int, const float*, long long, const float*, int, float*, void*, size_t, cudaStream_t) { return na(); }
}
