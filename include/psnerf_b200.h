/* psnerf_b200 — C ABI of the B200-native PS-NeRF render / shading hot path.
 *
 * The reference (ywq/psnerf) has no FFI on this path: its boundary is the Python nn.Module surface
 * (SURVEY.md §8b).  This header is the C-ABI a maintainer binds underneath those modules; every
 * entry point names the reference code it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - All tensors are contiguous fp32 DEVICE pointers unless marked "host".  The caller owns every
 *    input / output / workspace buffer; the library owns only psn_mlp handles.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no implicit sync.
 *  - Return 0 on success, <0 on error (PSN_ERR_*); psn_last_error() gives the message (thread local).
 *  - `precision`: PSN_PREC_FP32 = fp32 FFMA kernels; PSN_PREC_TC = tcgen05 tensor-core kernels with
 *    error-compensated fp16 split operands (hi+lo, 3 MMAs per product), fp32 accumulate in TMEM;
 *    PSN_PREC_TC_MIXED = the same kernels, except that the radiance program (psn_radiance, psn_render_unisurf) keeps the split
 *    product only for the eight softplus layers that decide alpha and runs the feature head, the reverse sweep that feeds the
 *    appearance MLP and the appearance MLP itself as single fp16 passes (rgb within 1e-5 rel-L2 of PSN_PREC_TC; alpha, depth,
 *    masks and the surface-normal output are bit-identical to PSN_PREC_TC).  Opt-in (the library default is PSN_PREC_FP32; bench.py runs it).
 *    PSN_PREC_TC_TWOLEVEL = PSN_PREC_TC_MIXED plus a two-level surface march (psn_raymarch, psn_render_unisurf): all proposal
 *    points go through a single-pass copy of the occupancy program, and the three-pass program re-evaluates only the points the
 *    scan can tell apart (within 0.02 of the threshold, next to a sign change, and their neighbours: < 1 % of them), so the
 *    crossing index, the bracket values and therefore every output are those of PSN_PREC_TC_MIXED.  EXPERIMENTAL: written after
 *    the round's GPU budget was spent; validated by CPU emulation only (tests/precision_study.py march_refine_study).
 *  - No CPU fallback exists: every entry point fails with PSN_ERR_CUDA without a sm_100 device.
 */
#ifndef PSNERF_B200_H
#define PSNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSN_OK 0
#define PSN_ERR_ARG (-1)
#define PSN_ERR_SHAPE (-2)
#define PSN_ERR_CUDA (-3)
#define PSN_ERR_WORKSPACE (-4)

#define PSN_PREC_FP32 0
#define PSN_PREC_TC 1
#define PSN_PREC_TC_MIXED 2
#define PSN_PREC_TC_TWOLEVEL 3

#define PSN_NET_GEO 0 /* stage1/model/network.py:37-66 lin0..lin8 (softplus beta=100, skip concat /sqrt2) */
#define PSN_NET_APP 1 /* stage1/model/network.py:71-79 lina0..lina4 (ReLU, tanh*0.5+0.5)               */
#define PSN_NET_S2 2  /* stage2/model/renderer.py:17-49 Network / Normal_Network (ReLU, cat[y,x] skip)   */

#define PSN_OUT_ALPHA 0     /* sigmoid(-10*logit)  network.py:124-125 */
#define PSN_OUT_NEG_LOGIT 1 /* -logit              network.py:137-138 */
#define PSN_OUT_LOGIT 2     /* raw logit (infer_occ[...,0]) */

typedef struct psn_mlp psn_mlp;

typedef struct {
  int kind;      /* PSN_NET_* */
  int n_layers;  /* number of Linear layers */
  int octaves;   /* GEO: octaves_pe of the point encoding; APP: octaves_pe_views; S2: unused */
  int skip;      /* GEO: index of the layer whose INPUT is cat[x, pe]/sqrt(2) (network.py:90-91), -1 = none.
                    S2: index of the layer AFTER whose output cat[y, x0] is formed (renderer.py:30-31), -1 = none */
  int final_act; /* S2: 0 = linear output (Normal_Network), 1 = sigmoid (Network) */
  float rescale; /* GEO: points are divided by this before encoding (network.py:86) */
} psn_mlp_desc;

int psn_version(void);
const char* psn_last_error(void);
/* Number of CUDA kernels this library has launched in this process (bench.py reports the per-step delta). */
int64_t psn_launch_count(void);
/* Per-kernel timing (CUDA events on the launching stream) for bench.py's roofline numbers. Tags: */
#define PSN_PROF_OCC_MARCH 0  /* occupancy MLP over ray-march proposals (rendering.py:457-462) */
#define PSN_PROF_OCC_SECANT 1 /* occupancy MLP inside the secant loop (rendering.py:540-546)     */
#define PSN_PROF_RADIANCE 2   /* geo fwd + analytic normal + app MLP per sample (network.py:122-136) */
#define PSN_PROF_GRADIENT 3   /* surface normals (rendering.py:208)                                */
#define PSN_PROF_SHADOW 4     /* shadow-ray occupancy (+ transmittance) (rendering.py:391-408)      */
#define PSN_PROF_S2_VIS 5     /* stage-2 visibility MLP over (light, point) pairs (renderer.py:193)  */
#define PSN_PROF_S2_POINT 6   /* stage-2 per-point MLPs (renderer.py:130,166,169)                   */
#define PSN_PROF_OCC_OTHER 7  /* explicit-point occupancy / infer_occ calls                          */
#define PSN_PROF_NTAGS 8
int psn_profile_enable(int on); /* clears previous records */
/* Sums per tag since enable: launches, elapsed ms, rows (samples) processed; waits for the recorded events. */
int psn_profile_collect(int n_tags, int64_t* launches, double* ms, double* rows);
/* 1 when this build contains the tcgen05 (PSN_PREC_TC) kernels. */
int psn_has_tensor_path(void);
/* Device probe: fails with PSN_ERR_CUDA unless a compute-capability-10.x GPU is current. */
int psn_device_check(int* sm_count);

/* Pack one network.  W[l] is the EFFECTIVE weight [out_dims[l], in_dims[l]] row-major (weight-norm already
 * folded: W = g*v/||v||, network.py:64,77), b[l] is [out_dims[l]].  Builds the fp32 k-major copies, the
 * transposed copies for the analytic-normal reverse pass and the swizzled fp16 hi/lo tiles for tcgen05. */
int psn_mlp_create(const psn_mlp_desc* desc, const int* in_dims, const int* out_dims,
                   const float* const* W, const float* const* b, void* stream, psn_mlp** out);
int psn_mlp_free(psn_mlp* net);

/* Workspace (bytes) needed by the calls below for the given problem size; query once, allocate, pass in. */
int64_t psn_workspace_bytes(const char* op, int64_t n_rays, int64_t n_samples, int64_t n_lights);

/* NeuralNetwork.forward(p, only_occupancy=True) / (return_logits=True): network.py:122-125,137-138. */
int psn_occupancy(const psn_mlp* geo, const float* pts /*[M,3]*/, int64_t M, int out_kind, float* out /*[M]*/,
                  int precision, void* stream);
/* NeuralNetwork.infer_occ: network.py:85-95.  out is [M, 1+feat]: logit then feature vector. */
int psn_infer_occ(const psn_mlp* geo, const float* pts, int64_t M, float* out, int precision, void* stream);
/* NeuralNetwork.gradient (d logit / d p, un-normalised): network.py:108-120, evaluated analytically. */
int psn_gradient(const psn_mlp* geo, const float* pts, int64_t M, float* grad /*[M,3]*/, void* ws, int64_t ws_bytes,
                 int precision, void* stream);
/* NeuralNetwork.forward(p, ray_d, return_addocc=True): network.py:126-134 -> rgb [M,3], alpha [M]. */
int psn_radiance(const psn_mlp* geo, const psn_mlp* app, const float* pts, const float* view_dirs /*[M,3]*/,
                 int64_t M, float* rgb, float* alpha, void* ws, int64_t ws_bytes, int precision, void* stream);

/* image_points_to_ray + origin_to_world + normalise: stage1/model/common.py:205-226, rendering.py:67-71.
 * cam (host, 16 floats) = R row-major[9], origin[3], fx, fy, cx, cy.  Stage 1 passes fy = fx (common.py:220).
 * pixels are float (x, y).  With normalize_like_stage2 != 0 it is get_camera_params/lift of
 * stage2/utils/rend_util.py:90-147 (F.normalize, separate fx/fy). */
int psn_rays_from_pixels(const float* pixels /*[N,2]*/, int64_t N, const float* cam /*host[16]*/,
                         int normalize_like_stage2, float* dirs /*[N,3]*/, void* stream);

/* Renderer.ray_marching + secant: rendering.py:410-555.  depth[N]: +inf = no surface, 0 = first proposal
 * occupied, else the refined depth.  Sphere far depth from get_sphere_intersection (rendering.py:576-596). */
int psn_raymarch(const psn_mlp* geo, const float* origin /*host[3]*/, const float* dirs /*[N,3]*/, int64_t N,
                 float near, float radius, int n_steps, int n_secant, float tau, float* depth /*[N]*/,
                 void* ws, int64_t ws_bytes, int precision, void* stream);

typedef struct {
  float near_, radius, delta, tau;
  int march_steps, secant_steps;
  int steps_in;  /* num_points_in  (samples inside the surface interval) */
  int steps_out; /* num_points_out, or 0 when the reference uses full_steps == steps (rendering.py:124-127) */
  int white_background;
} psn_unisurf_params;

/* Renderer.unisurf (eval path, optional externally supplied jitter): rendering.py:50-226.
 * noise: nullable [N, steps_in+steps_out] uniform(0,1) samples replacing torch.rand (rendering.py:139,163).
 * Outputs: rgb[N,3], acc[N], normal[N,3] (g/(|g|+1e-5) on hit rays, 0 elsewhere), mask[N] (uint8),
 * depth[N] (raw ray_marching result), sample_depth (nullable) [N, steps_in+steps_out]. */
int psn_render_unisurf(const psn_mlp* geo, const psn_mlp* app, const float* origin /*host[3]*/,
                       const float* dirs /*[N,3]*/, int64_t N, const psn_unisurf_params* prm,
                       const float* noise, float* rgb, float* acc, float* normal, uint8_t* mask, float* depth,
                       float* sample_depth, void* ws, int64_t ws_bytes, int precision, void* stream);

/* Renderer.light_visibility: rendering.py:378-408.  vis[L, Ns] light-major = 1 - sum_i a_i prod_{j<i}(1-a_j+1e-6),
 * occupancy zeroed outside the [-box, box]^3 cube. */
int psn_shadow_visibility(const psn_mlp* geo, const float* surf /*[Ns,3]*/, const float* lights /*[L,3]*/,
                          int64_t Ns, int L, float lnear, float lfar, int n_steps, float box, float* vis,
                          void* ws, int64_t ws_bytes, int precision, void* stream);

typedef struct {
  int n_freqs_xyz;    /* brdf.net.n_freqs_xyz (embedder.py:39-54) */
  int n_freqs_normal; /* normal.net.n_freqs_xyz */
  int nbasis;         /* lobes per channel (sgbasis.py:7-14) */
  int specular_rgb;   /* train.specular_rgb: weights are [3, nbasis] */
  int intensity_kind; /* 0 scalar, 1 per-light [L,1], 2 per-light rgb [L,3] (renderer.py:188-190) */
  float intensity;    /* scalar intensity when intensity_kind == 0 */
  int render_model;   /* train.render_model: 0 = sgbasis (sgbasis.py), 1 = microfacet (GGX, stage2/model/microfacet.py:35-114):
                         rough_net then has ONE sigmoid output (the roughness), sgw is not written and spec is the per-PIXEL
                         roughness image [N,3] (renderer.py:140-141,204-207) instead of [L,N,3] */
  float fresnel_f0;   /* brdf.fresnel_f0 (Schlick), microfacet only */
} psn_shade_params;

/* PSNetwork.forward, eval path: stage2/model/renderer.py:110-266 with sgbasis.py:16-32, embedder.py:36.
 * Inputs are per SURFACE point (already gathered by surface_mask): pts, view (= -ray_dir), normal_in
 * (used when normal_net is NULL), pix[Ns] = pixel index of the point in [0,N).
 * Outputs are image shaped and fully written by the library, non-surface pixels pre-filled exactly as the
 * reference does (1.0; sg weights 0: renderer.py:145-152):
 *   rgb[L,N,3], spec[L,N,3], vis[L,N,3] (nullable when vis_net is NULL), normal[N,3], albedo[N,3], sgw[N,nbt]. */
int psn_shade_stage2(const psn_mlp* normal_net, const psn_mlp* albedo_net, const psn_mlp* rough_net,
                     const psn_mlp* vis_net, const float* lobe, const psn_shade_params* prm,
                     const float* pts, const float* view, const float* normal_in, const int32_t* pix,
                     int64_t Ns, int64_t N, const float* lights /*[L,3]*/, int L, const float* intensity,
                     float* rgb, float* spec, float* vis, float* normal, float* albedo, float* sgw,
                     void* ws, int64_t ws_bytes, int precision, void* stream);

/* Material editing (stage2/eval.py:116-132, renderer.py:167-168,175-181): psn_shade_stage2 with the albedo of every surface
 * point replaced by albedo_new[3] and / or its SG weights by weights_new[nbt] (device pointers, either may be NULL). */
int psn_shade_stage2_edit(const psn_mlp* normal_net, const psn_mlp* albedo_net, const psn_mlp* rough_net,
                          const psn_mlp* vis_net, const float* lobe, const psn_shade_params* prm,
                          const float* pts, const float* view, const float* normal_in, const int32_t* pix,
                          int64_t Ns, int64_t N, const float* lights /*[L,3]*/, int L, const float* intensity,
                          const float* albedo_new, const float* weights_new,
                          float* rgb, float* spec, float* vis, float* normal, float* albedo, float* sgw,
                          void* ws, int64_t ws_bytes, int precision, void* stream);

/* Per-point stage-2 nets only (albedo / rough evaluated at jittered points, renderer.py:211-231): outputs are
 * per surface point: albedo[Ns,3] (sigmoid), weights[Ns,nbt] (relu). */
int psn_s2_point_nets(const psn_mlp* albedo_net, const psn_mlp* rough_net, int n_freqs, const float* pts,
                      int64_t Ns, float* albedo, float* weights, int nbt, int precision, void* stream);
/* visibility_net over (light, point) pairs, light-major [L, Ns] raw (unclamped) outputs (renderer.py:193, :251-262). */
int psn_s2_visibility(const psn_mlp* vis_net, int n_freqs, const float* pts, int64_t Ns, const float* lights, int L,
                      float* vis /*[L,Ns]*/, void* ws, int64_t ws_bytes, int precision, void* stream);

/* Bring-up / test hook of the tensor-core path: activations handed from geo layer `layer` (0..7) to the next one,
 * fp32 before the fp16 hi/lo split, out[M,256]; logits[M] receives the resulting logit. */
int psn_tc_debug_layer(const psn_mlp* geo, const float* pts, int64_t M, int layer, float* out, float* logits, void* stream);

/* Bring-up tool: clock64() timeline of one tile of the tensor-core occupancy kernel; trace is int64[256] on the device. */
int psn_tc_debug_trace(const psn_mlp* geo, const float* pts, int64_t M, float* out, long long* trace, void* stream);
/* same, epilogue time stamps taken by the warps of TMEM lane quadrant `quadrant` (0..3; each quadrant lives on its own scheduler) */
int psn_tc_debug_trace_q(const psn_mlp* geo, const float* pts, int64_t M, float* out, long long* trace, int quadrant, void* stream);

/* Same for the radiance kernel (23 steps per tile): MMA-lane slots step*8 + {0..3 a_ready, 4 wait-activations, 5 wait-weights,
 * 7 last commit}; trace[192 + step] = epilogue of the step finished (row 0).  stash: psn_workspace-style scratch of at least
 * 148 * 640 KB. */
int psn_tc_debug_trace_rad(const psn_mlp* geo, const psn_mlp* app, const float* pts, const float* views, int64_t M,
                           float* rgb, float* alpha, void* stash, long long* trace, int mixed, void* stream);

/* Test hook for the GEMM of the train steps (csrc/tc_gemm.cu: tcgen05 kind::tf32, three split passes, fp32 accumulate):
 * form 0: C[M,N] = A[M,K] B[N,K]^T (+bias, epi 1 / relu 2 / sigmoid 3); form 1: C[M,N] = A[M,K] B[K,N];
 * form 2: C[M,N] += A[K,M]^T B[K,N].  Row-major fp32 device matrices with leading dimensions lda / ldb / ldc. */
int psn_tc_gemm_debug(int form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias,
                      int64_t M, int N, int64_t K, int epi, void* stream);
/* Same with the fused element-wise epilogues of the stage-1 train step (form 0 / 1; C2 / E1 / E2 are [M,N] with leading dimension lde):
 * epi 4: C = softplus_100(acc + bias), C2 = sigmoid(100 (acc + bias));  5: C = acc E1, E2 := acc E2 100 E1 (1 - E1);
 * 6: C = acc scale E1 + E2;  7: C = acc, C2 = acc E1.  (stage1/model/network.py:88-92,108-120 forward / double backward pieces) */
int psn_tc_gemm_debug_fused(int form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias,
                            int64_t M, int N, int64_t K, int epi, float* C2, const float* E1, float* E2, int64_t lde, float scale,
                            void* stream);

/* ---- stage-2 train step (BASELINE config 5): PSNetwork.forward + backward, stage2/trainer.py:394-410 ------------------------
 * Gradient-carrying parts (renderer.py:193-199,211-231,251-262 with light_vis_detach = vis_rgb_detach = True): the per-point nets
 * through the SG shading of all L lights and their jittered re-evaluation, light directions / intensities, and visibility_net
 * through the Lt vis-train lights.  Networks are given as views of the live torch parameters. */
typedef struct {
  int n_layers, skip, final_act;   /* as psn_mlp_desc for PSN_NET_S2 */
  const int* in_dims;              /* host */
  const int* out_dims;             /* host */
  const float* const* W;           /* host array of device pointers: [out,in] row-major */
  const float* const* b;
  float* const* dW;                /* backward only: gradients are ACCUMULATED into these (caller zero-fills) */
  float* const* db;
} psn_train_net;

int64_t psn_s2_train_tape_bytes(const psn_train_net* normal_net, const psn_train_net* albedo_net, const psn_train_net* rough_net,
                                const psn_train_net* vis_net, int64_t Ns, int L, int Lt);
/* Forward: image-shaped outputs exactly as psn_shade_stage2, plus per-surface-point albedo_j[Ns,3], weights_j[Ns,nbt]
 * (networks re-evaluated at jitter_pts, nullable) and vis_train[Lt,Ns] (raw visibility_net outputs for lights_vt, nullable).
 * vis_packed: the packed visibility net for the detached L-light pass.  tape keeps the activations for the backward call. */
int psn_s2_train_forward(const psn_train_net* normal_net, const psn_train_net* albedo_net, const psn_train_net* rough_net,
                         const psn_train_net* vis_net, const psn_mlp* vis_packed, const float* lobe, const psn_shade_params* prm,
                         const float* pts, const float* view, const int32_t* pix, int64_t Ns, int64_t N, const float* lights, int L,
                         const float* intensity, const float* jitter_pts, const float* lights_vt, int Lt, float* rgb, float* spec,
                         float* vis, float* normal, float* albedo, float* sgw, float* albedo_j, float* weights_j, float* vis_train,
                         void* tape, int64_t tape_bytes, void* ws, int64_t ws_bytes, int precision, void* stream);
/* Backward: g_* are gradients w.r.t. the forward outputs (any may be NULL = zero): g_rgb/g_spec [L,N,3], g_normal [N,3],
 * g_albedo [N,3], g_sgw [N,nbt], g_albedo_j [Ns,3], g_weights_j [Ns,nbt], g_vis_train [Lt,Ns].  Writes (accumulates) parameter
 * gradients through the psn_train_net views, d_lights [L,3] and d_intensity (scalar / [L] / [L,3] per prm->intensity_kind). */
int psn_s2_train_backward(const psn_train_net* normal_net, const psn_train_net* albedo_net, const psn_train_net* rough_net,
                          const psn_train_net* vis_net, const float* lobe, const psn_shade_params* prm, const float* view,
                          const int32_t* pix, int64_t Ns, int64_t N, const float* lights, int L, const float* intensity, int Lt,
                          const float* g_rgb, const float* g_spec, const float* g_normal, const float* g_albedo, const float* g_sgw,
                          const float* g_albedo_j, const float* g_weights_j, const float* g_vis_train, float* d_lights,
                          float* d_intensity, void* tape, int64_t tape_bytes, void* ws, int64_t ws_bytes, void* stream);

/* ---- stage-1 train step: the differentiable field (stage1/model/network.py:85-136 under autograd, create_graph normals) ----------
 * psn_train_net views carry the EFFECTIVE weights (weight norm folded: W = g v / |v|, network.py:64,77) of the geo net
 * (n_layers = 9, skip = index of the layer whose input is cat[x, pe]/sqrt2) and of the appearance net (nullable: gradient-only
 * evaluation, e.g. the surface normals of rendering.py:203-211).  Forward outputs per sample: rgb [M,3] (app only), logit [M]
 * (nullable), grad [M,3] = d logit / d p.  The backward takes the cotangents of those three (any may be NULL) and ACCUMULATES
 * the weight / bias gradients, including the double-backward route through grad.  tape: psn_s1_train_tape_bytes, written by
 * the forward, consumed (and clobbered) by ONE backward; ws: psn_s1_train_ws_bytes. */
int64_t psn_s1_train_tape_bytes(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, int64_t M);
int64_t psn_s1_train_ws_bytes(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, int64_t M);
int psn_s1_train_forward(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, float rescale,
                         const float* pts, const float* views, int64_t M, float* rgb, float* logit, float* grad,
                         void* tape, int64_t tape_bytes, void* ws, int64_t ws_bytes, void* stream);
int psn_s1_train_backward(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, float rescale,
                          int64_t M, const float* g_rgb, const float* g_logit, const float* g_grad,
                          void* tape, int64_t tape_bytes, void* ws, int64_t ws_bytes, void* stream);
/* Backward of psn_composite: d_rgb_s [N,S,3], d_alpha [N,S] from g_rgb [N,3] / g_acc [N] (either may be NULL), S <= 256. */
int psn_composite_bwd(const float* rgb_s, const float* alpha, int64_t N, int S, int white_background,
                      const float* g_rgb, const float* g_acc, float* d_rgb_s, float* d_alpha, void* stream);

/* ---- fused optimizer steps of the train loops ---------------------------------------------------------------
 * psn_adam_step replaces torch.optim.Adam.step() (stage1/train.py:62, stage2/trainer.py:116; amsgrad = False): ONE launch
 * updates every listed tensor in place (param, exp_avg, exp_avg_sq; grad is read).  `step` is the 1-based update count AFTER this
 * call's increment (torch's state['step']); tensors whose counts differ need separate calls.  weight_decay is torch's L2 term
 * (grad + weight_decay * param).  All pointers are contiguous fp32 device memory; the tensor list itself is HOST memory.
 * psn_sparse_adam_step replaces torch.optim.SparseAdam.step() for one [R, D] embedding (stage2/trainer.py:165: the per-light
 * direction [llen,3] and intensity [llen,1] tables): only the rows named in `rows` [K] (int64, device; duplicates are summed
 * in entry order like coalesce(), rows outside [0, R) are ignored) move, by grad_values [K, D].  K <= 65536. */
typedef struct psn_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} psn_adam_tensor;
typedef struct psn_adam_hyper { /* doubles, as torch holds them: (float)(1 - beta2) of the DOUBLE 0.999 is what torch multiplies with */
  double lr, beta1, beta2, eps, weight_decay;
  int64_t step;
} psn_adam_hyper;
int psn_adam_step(const psn_adam_tensor* tensors, int n_tensors, const psn_adam_hyper* hyper, void* stream);
int psn_sparse_adam_step(float* param, float* exp_avg, float* exp_avg_sq, int64_t R, int D, const int64_t* rows,
                         const float* grad_values, int64_t K, const psn_adam_hyper* hyper, void* stream);

/* alpha compositing of per-sample (rgb, alpha): rendering.py:196-197,214-216. */
int psn_composite(const float* rgb_s /*[N,S,3]*/, const float* alpha /*[N,S]*/, int64_t N, int S,
                  int white_background, float* rgb /*[N,3]*/, float* acc /*[N]*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSNERF_B200_H */
