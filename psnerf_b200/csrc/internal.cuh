// Cross-translation-unit declarations (host side).
#pragma once
#include "stage1_simt.cuh"

namespace psn {

struct SecantState {
  int* count;    // number of active (masked) rays
  int* ray;      // [N] ray id per slot
  float *d_low, *d_high, *f_low, *f_high, *d_pred;  // [N] per slot
};
struct SurfList { int* count; int* ray; float* depth; };
// Two-level march: the proposal points the full program has to re-evaluate.  count[0] = entries (clamped to cap), count[1] = raw
// number of selected points, count[2] = points of the whole-march fallback (N * S if the list overflowed, else 0).
struct RefineList { int* count; int* ray; float* depth; int* pos; int cap; };

// Box-culled shadow pass: the (pair, step) entries of the in-box samples of one light chunk (stage1_aux.cu:k_shadow_plan*).
// Two lists: A = the first `lead` in-box steps of every pair, B = the remaining in-box steps of the pairs whose transmittance after
// list A is still above the death threshold.  total[0] / total[1] = entries of A / B (the MLP kernels read them as their device-side
// row counts); evaluated = running 64-bit sum over the chunks of a call; meta[pair] = offset in A | first step << 32 | in-box steps
// << 40; off_b[pair] = offset in B or 0xffffffff (no B entries); entry = A region [0, cap_a) followed by the B region.
struct ShadowList {
  unsigned* total;
  unsigned long long* evaluated;
  unsigned long long* meta;
  unsigned* off_b;
  unsigned* entry;
  unsigned cap_a;
  int lead;
};

// stage1_simt.cu
int simt_occupancy(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, int out_kind, float* out,
                   int with_feat, cudaStream_t st);
int simt_secant(const psn_mlp* geo, const PointGen& gen, const SecantState& sec, int n_iter, float tau, cudaStream_t st);
int simt_gradient(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, float* grad, void* stash,
                  cudaStream_t st);
int simt_radiance(const psn_mlp* geo, const psn_mlp* app, const PointGen& gen, long long M, float* rgb, float* alpha,
                  void* stash, cudaStream_t st);
size_t simt_stash_bytes();
int make_geo_dev(const psn_mlp* net, GeoDev* g);

// tc_*.cu (tcgen05 path)
int tc_occupancy(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, int out_kind, float* out,
                 cudaStream_t st);
int tc_occupancy_cheap(const psn_mlp* geo, const PointGen& gen, long long M, int out_kind, float* out, cudaStream_t st);
int tc_infer_occ(const psn_mlp* geo, const PointGen& gen, long long M, float* out, cudaStream_t st);
int tc_gradient(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, float* grad, void* stash,
                cudaStream_t st);
int tc_radiance(const psn_mlp* geo, const psn_mlp* app, const PointGen& gen, long long M, float* rgb, float* alpha,
                void* stash, int mixed, cudaStream_t st);
// PSN_PREC_TC, _TC_MIXED and _TC_TWOLEVEL all select the tcgen05 kernels; MIXED and TWOLEVEL run the mixed radiance program
// (tc_rad.cu), TWOLEVEL additionally the two-level surface march (api_stage1.cu raymarch_impl)
inline bool prec_is_tc(int precision) { return precision >= PSN_PREC_TC && precision <= PSN_PREC_TC_TWOLEVEL; }
inline bool prec_is_mixed(int precision) { return precision == PSN_PREC_TC_MIXED || precision == PSN_PREC_TC_TWOLEVEL; }
int tc_shadow(const psn_mlp* geo, const PointGen& gen, long long pairs, float box, float* vis, cudaStream_t st);
size_t tc_stash_bytes();

// stage1_aux.cu
int launch_rays(const float* pix, long long N, const float* cam, int stage2, float* dirs, cudaStream_t st);
int launch_sphere_far(const float* dirs, long long N, const float* o, float r, float* far, cudaStream_t st);
int launch_march_scan(const float* occ, const float* far, long long N, int S, float near_, float tau, SecantState s,
                      float* depth, cudaStream_t st);
int launch_secant_update(SecantState s, const float* occ_mid, float tau, long long N, cudaStream_t st);
int launch_march_refine_select(const float* occ, const float* far, long long N, int S, float near_, float tau, float margin,
                               RefineList rl, cudaStream_t st);
int launch_march_refine_scatter(RefineList rl, const float* refined, float* occ, cudaStream_t st);
int launch_march_finalize(SecantState s, float* depth, long long N, cudaStream_t st);
int launch_sample_plan(const float* d_i, const float* far, long long N, const psn_unisurf_params& prm, const float* noise,
                       float* sample_depth, uint8_t* mask, SurfList sl, cudaStream_t st);
int launch_composite(const float* rgb_s, const float* alpha, long long N, int S, int white, float* rgb, float* acc,
                     cudaStream_t st);
int launch_shadow_composite(const float* occ, const float* surf, const float* lights, long long Ns, long long pairs, int S,
                            float lnear, float lfar, float box, float* vis, cudaStream_t st);
int launch_shadow_plan(const float* surf, const float* lights, long long Ns, long long pairs, int S, float lnear, float lfar,
                       float box, ShadowList sl, cudaStream_t st);
int launch_shadow_plan_b(const float* occ, long long pairs, ShadowList sl, cudaStream_t st);
int launch_shadow_composite_list(const float* occ, ShadowList sl, const float* surf, const float* lights, long long Ns,
                                 long long pairs, int S, float lnear, float lfar, float box, float* vis, cudaStream_t st);
int launch_scatter_normals(const float* grad, SurfList sl, float* normal, long long N, cudaStream_t st);

// stage2_simt.cu
size_t s2_workspace_bytes(long long Ns, long long L);

}  // namespace psn
