// C-ABI entry points of the stage-1 path (see include/psnerf_b200.h for the reference lines each replaces).
#include <stdlib.h>
#include "stage1_simt.cuh"
#include "launch.cuh"
#include "internal.cuh"
#include "prof.cuh"

using namespace psn;

namespace psn {

// Evaluate occupancy probabilities for generated points with the requested arithmetic.
static int occupancy_any(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, int out_kind, float* out,
                         int precision, cudaStream_t st) {
  const int tag = gen.kind == GEN_MARCH ? PSN_PROF_OCC_MARCH : gen.kind == GEN_INDEXED_DEPTH ? PSN_PROF_OCC_SECANT
                : gen.kind == GEN_SHADOW ? PSN_PROF_SHADOW : PSN_PROF_OCC_OTHER;
  ProfScope prof(tag, M, st);
  if (prec_is_tc(precision)) return tc_occupancy(geo, gen, M, M_dev, out_kind, out, st);
  return simt_occupancy(geo, gen, M, M_dev, out_kind, out, 0, st);
}
static int gradient_any(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, float* grad, void* stash,
                        int precision, cudaStream_t st) {
  ProfScope prof(PSN_PROF_GRADIENT, M, st);
  if (prec_is_tc(precision)) return tc_gradient(geo, gen, M, M_dev, grad, stash, st);
  return simt_gradient(geo, gen, M, M_dev, grad, stash, st);
}
static int radiance_any(const psn_mlp* geo, const psn_mlp* app, const PointGen& gen, long long M, float* rgb, float* alpha,
                        void* stash, int precision, cudaStream_t st) {
  ProfScope prof(PSN_PROF_RADIANCE, M, st);
  if (prec_is_tc(precision)) return tc_radiance(geo, app, gen, M, rgb, alpha, stash, prec_is_mixed(precision) ? 1 : 0, st);
  return simt_radiance(geo, app, gen, M, rgb, alpha, stash, st);
}

static size_t stash_bytes(int precision) { return prec_is_tc(precision) ? tc_stash_bytes() : simt_stash_bytes(); }
static size_t stash_bytes_any() { return tc_stash_bytes() > simt_stash_bytes() ? tc_stash_bytes() : simt_stash_bytes(); }

// Accuracy policy of the tensor-core precisions.  The O(rays x samples) launches (march proposals, radiance samples, shadow samples)
// run on tcgen05; two O(rays) computations decide quantities every later stage amplifies and run on the fp32 FFMA kernels instead:
//  * the secant refinement - its result is the root of the occupancy AS EVALUATED; the split-operand logit carries ~1e-5 of rounding
//    (48 truncating tensor-core accumulations per layer), which near the root is as large as the bracket values themselves: wrong-signed
//    updates push the bracket ~3e-5 off the fp32 root (measured, profiles/r2_parity_errlog_*.jsonl), and the shadow visibilities
//    (x 100) and the 2^9-frequency encodings of stage 2 multiply that;
//  * the surface-normal OUTPUT (one gradient evaluation per hit ray).
// Cost: 8 x 1.3 ms + 3 ms per 512 x 512 view (2 % of the relit-view step), in ONE launch each.
static int normal_precision(int precision) { return prec_is_tc(precision) ? PSN_PREC_FP32 : precision; }

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// Two-level march: at most 1/8 of the proposal points on the refine list (measured: < 1 %); beyond that the whole march falls back
// to the full program on the device (RefineList::count[2]), so the cap bounds memory, not correctness.
static long long refine_cap(long long N, int S) {
  const long long c = N * S / 8;
  return c < 4096 ? 4096 : c;
}
static const float kRefineMargin = 0.02f;  // >= 20 x the largest |cheap - full| occupancy difference seen (tests/precision_study.py)

static size_t march_ws_bytes(long long N, int S) {
  return align256((size_t)N * S * 4) + 8 * align256((size_t)N * 4) + 1024 + 4 * align256((size_t)refine_cap(N, S) * 4) + 512;
}
static const int kShadowLead = 16;  // in-box steps of every shadow ray evaluated before its transmittance decides about the rest (<= 32)
static const long long kShadowChunkPairs = 1 << 21;  // (light, point) pairs per shadow chunk (1 GB of occupancies at S=128)
// A/B switch for measurements: PSNERF_B200_SHADOW_UNCULLED=1 evaluates every step of every shadow ray like the reference does
// (fused k_tc_occ<MODE_SHADOW> on the tensor path) instead of the box-culled list.  Results agree to rounding of the product order.
static bool secant_unfused_requested() {
  const char* e = getenv("PSNERF_B200_SECANT_UNFUSED");
  return e && e[0] == '1';
}
static bool shadow_unculled_requested() {
  const char* e = getenv("PSNERF_B200_SHADOW_UNCULLED");
  return e && e[0] == '1';
}

}  // namespace psn

extern "C" int64_t psn_workspace_bytes(const char* op, int64_t n_rays, int64_t n_samples, int64_t n_lights) {
  if (!op) return -1;
  const size_t stash = (tc_stash_bytes() > simt_stash_bytes() ? tc_stash_bytes() : simt_stash_bytes()) + 4096;
  const long long N = n_rays < 1 ? 1 : n_rays, S = n_samples < 1 ? 1 : n_samples;
  if (!strcmp(op, "gradient") || !strcmp(op, "radiance")) return (int64_t)stash;
  if (!strcmp(op, "raymarch")) return (int64_t)march_ws_bytes(N, (int)S);
  if (!strcmp(op, "unisurf")) {
    // n_samples = max(march_steps, samples per ray); both regions are sized with it.
    return (int64_t)(march_ws_bytes(N, (int)S) + align256((size_t)N * S * 4) * 5 + 6 * align256((size_t)N * 4) +
                     align256((size_t)N * 12) + stash + 4096);
  }
  if (!strcmp(op, "shadow")) {
    const long long pairs = N * (n_lights < 1 ? 1 : n_lights);
    const long long chunk = pairs < kShadowChunkPairs + N ? pairs : kShadowChunkPairs + N;
    return (int64_t)(align256((size_t)chunk * S * 4) + align256((size_t)chunk * 8) + align256((size_t)chunk * 4) + 4096);
  }
  if (!strcmp(op, "shade") || !strcmp(op, "s2_vis")) return (int64_t)s2_workspace_bytes(N, n_lights);
  if (!strcmp(op, "s2_train")) {  // n_rays = max(pixels, surface points), n_samples = vis-train lights, n_lights = L
    const long long rows = N * (S > 1 ? S : 1);
    return (int64_t)(s2_workspace_bytes(N, n_lights) + (size_t)rows * (2 * 512 + 32) * 4 + (size_t)N * 64 * 4 + 8192);
  }
  psn::set_error("psn_workspace_bytes: unknown op '%s'", op);
  return -1;
}

extern "C" int psn_occupancy(const psn_mlp* geo, const float* pts, int64_t M, int out_kind, float* out, int precision,
                             void* stream) {
  PSN_REQUIRE(geo && (M == 0 || (pts && out)), PSN_ERR_ARG, "psn_occupancy: null argument");
  PSN_REQUIRE(out_kind >= 0 && out_kind <= 2, PSN_ERR_ARG, "psn_occupancy: bad out_kind %d", out_kind);
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_EXPLICIT;
  gen.pts = pts;
  return occupancy_any(geo, gen, M, nullptr, out_kind, out, precision, (cudaStream_t)stream);
}

extern "C" int psn_infer_occ(const psn_mlp* geo, const float* pts, int64_t M, float* out, int precision, void* stream) {
  PSN_REQUIRE(geo && (M == 0 || (pts && out)), PSN_ERR_ARG, "psn_infer_occ: null argument");
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_EXPLICIT;
  gen.pts = pts;
  if (prec_is_tc(precision)) return tc_infer_occ(geo, gen, M, out, (cudaStream_t)stream);
  return simt_occupancy(geo, gen, M, nullptr, PSN_OUT_LOGIT, out, 1, (cudaStream_t)stream);
}

extern "C" int psn_gradient(const psn_mlp* geo, const float* pts, int64_t M, float* grad, void* ws, int64_t ws_bytes,
                            int precision, void* stream) {
  PSN_REQUIRE(geo && (M == 0 || (pts && grad)), PSN_ERR_ARG, "psn_gradient: null argument");
  Workspace w(ws, ws_bytes);
  void* stash = w.take<char>(stash_bytes(precision));
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_gradient: workspace too small (%lld < %zu)", (long long)ws_bytes, w.used);
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_EXPLICIT;
  gen.pts = pts;
  return gradient_any(geo, gen, M, nullptr, grad, stash, precision, (cudaStream_t)stream);
}

extern "C" int psn_radiance(const psn_mlp* geo, const psn_mlp* app, const float* pts, const float* view_dirs, int64_t M,
                            float* rgb, float* alpha, void* ws, int64_t ws_bytes, int precision, void* stream) {
  PSN_REQUIRE(geo && app && (M == 0 || (pts && view_dirs && rgb && alpha)), PSN_ERR_ARG, "psn_radiance: null argument");
  Workspace w(ws, ws_bytes);
  void* stash = w.take<char>(stash_bytes(precision));
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_radiance: workspace too small (%lld < %zu)", (long long)ws_bytes, w.used);
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_EXPLICIT;
  gen.pts = pts;
  gen.views = view_dirs;
  return radiance_any(geo, app, gen, M, rgb, alpha, stash, precision, (cudaStream_t)stream);
}

extern "C" int psn_rays_from_pixels(const float* pixels, int64_t N, const float* cam, int normalize_like_stage2, float* dirs,
                                    void* stream) {
  PSN_REQUIRE(cam && (N == 0 || (pixels && dirs)), PSN_ERR_ARG, "psn_rays_from_pixels: null argument");
  return launch_rays(pixels, N, cam, normalize_like_stage2, dirs, (cudaStream_t)stream);
}

namespace psn {
// ray_marching + secant on device; leaves far[N] and the per-ray depth.  rendering.py:410-555.
static int raymarch_impl(const psn_mlp* geo, const float* origin, const float* dirs, long long N, float near_, float radius,
                         int n_steps, int n_secant, float tau, float* depth, float* far, Workspace& w, int precision,
                         cudaStream_t st) {
  float* occ = w.take<float>((size_t)N * n_steps);
  SecantState s;
  s.count = w.take<int>(64);
  s.ray = w.take<int>(N);
  s.d_low = w.take<float>(N); s.d_high = w.take<float>(N); s.f_low = w.take<float>(N); s.f_high = w.take<float>(N);
  s.d_pred = w.take<float>(N);
  float* occ_mid = w.take<float>(N);
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "ray marching: workspace too small (need %zu bytes)", w.used);
  int rc;
  if ((rc = launch_sphere_far(dirs, N, origin, radius, far, st))) return rc;
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_MARCH;
  gen.dirs = dirs;
  gen.far = far;
  gen.near_ = near_;
  gen.S = n_steps;
  gen.o[0] = origin[0]; gen.o[1] = origin[1]; gen.o[2] = origin[2];
  if (precision == PSN_PREC_TC_TWOLEVEL && n_steps >= 4 && N * n_steps < (1LL << 31)) {
    // Level 1: every proposal point through the single-pass program.  Level 2: the full program on the points the scan below can
    // tell apart (k_march_refine_select), written back over the level-1 values.  The scan reads signs everywhere and values only
    // at the crossing, so crossing index and bracket values - hence every output - equal those of the full evaluation.
    RefineList rl;
    rl.cap = (int)refine_cap(N, n_steps);
    rl.count = w.take<int>(64);
    rl.ray = w.take<int>(rl.cap);
    rl.depth = w.take<float>(rl.cap);
    rl.pos = w.take<int>(rl.cap);
    float* refined = w.take<float>(rl.cap);
    PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "two-level ray marching: workspace too small (need %zu bytes)", w.used);
    {
      ProfScope prof(PSN_PROF_OCC_MARCH, N * n_steps, st);
      if ((rc = tc_occupancy_cheap(geo, gen, N * n_steps, PSN_OUT_ALPHA, occ, st))) return rc;
    }
    PSN_CUDA_CHECK(cudaMemsetAsync(rl.count, 0, 4 * sizeof(int), st));
    if ((rc = launch_march_refine_select(occ, far, N, n_steps, near_, tau, kRefineMargin, rl, st))) return rc;
    PointGen g3;
    memset(&g3, 0, sizeof(g3));
    g3.kind = GEN_INDEXED_DEPTH;
    g3.dirs = dirs;
    g3.index = rl.ray;
    g3.depth = rl.depth;
    g3.o[0] = origin[0]; g3.o[1] = origin[1]; g3.o[2] = origin[2];
    {
      ProfScope prof(PSN_PROF_OCC_OTHER, 0, st);
      if ((rc = tc_occupancy(geo, g3, 0, rl.count, PSN_OUT_ALPHA, refined, st))) return rc;
      if ((rc = launch_march_refine_scatter(rl, refined, occ, st))) return rc;
      if ((rc = tc_occupancy(geo, gen, 0, rl.count + 2, PSN_OUT_ALPHA, occ, st))) return rc;  // list overflow: the whole march again (0 points normally)
    }
  } else {
    if ((rc = occupancy_any(geo, gen, N * n_steps, nullptr, PSN_OUT_ALPHA, occ, precision, st))) return rc;
  }
  PSN_CUDA_CHECK(cudaMemsetAsync(s.count, 0, sizeof(int), st));
  if ((rc = launch_march_scan(occ, far, N, n_steps, near_, tau, s, depth, st))) return rc;
  PointGen g2;
  memset(&g2, 0, sizeof(g2));
  g2.kind = GEN_INDEXED_DEPTH;
  g2.dirs = dirs;
  g2.index = s.ray;
  g2.depth = s.d_pred;
  g2.o[0] = origin[0]; g2.o[1] = origin[1]; g2.o[2] = origin[2];
  // All n_secant iterations in ONE launch of the fp32 kernel, the bracket kept on chip (k_geo_secant); the tensor precisions use it
  // too (accuracy policy above).  PSNERF_B200_SECANT_UNFUSED=1 runs one evaluation launch + one update kernel per iteration
  // instead (bit-identical: A/B tests).
  if (!secant_unfused_requested()) {
    ProfScope prof(PSN_PROF_OCC_SECANT, 0, st);
    if ((rc = simt_secant(geo, g2, s, n_secant, tau, st))) return rc;
  } else {
    for (int it = 0; it < n_secant; ++it) {
      if ((rc = occupancy_any(geo, g2, 0, s.count, PSN_OUT_ALPHA, occ_mid, PSN_PREC_FP32, st))) return rc;
      if ((rc = launch_secant_update(s, occ_mid, tau, N, st))) return rc;
    }
  }
  return launch_march_finalize(s, depth, N, st);
}
}  // namespace psn

extern "C" int psn_raymarch(const psn_mlp* geo, const float* origin, const float* dirs, int64_t N, float near_, float radius,
                            int n_steps, int n_secant, float tau, float* depth, void* ws, int64_t ws_bytes, int precision,
                            void* stream) {
  PSN_REQUIRE(geo && origin && (N == 0 || (dirs && depth)), PSN_ERR_ARG, "psn_raymarch: null argument");
  PSN_REQUIRE(n_steps >= 2 && n_secant >= 0, PSN_ERR_ARG, "psn_raymarch: n_steps=%d n_secant=%d", n_steps, n_secant);
  if (N == 0) return PSN_OK;
  Workspace w(ws, ws_bytes);
  float* far = w.take<float>(N);
  return raymarch_impl(geo, origin, dirs, N, near_, radius, n_steps, n_secant, tau, depth, far, w, precision,
                       (cudaStream_t)stream);
}

extern "C" int psn_render_unisurf(const psn_mlp* geo, const psn_mlp* app, const float* origin, const float* dirs, int64_t N,
                                  const psn_unisurf_params* prm, const float* noise, float* rgb, float* acc, float* normal,
                                  uint8_t* mask, float* depth, float* sample_depth, void* ws, int64_t ws_bytes, int precision,
                                  void* stream) {
  PSN_REQUIRE(geo && app && origin && prm && (N == 0 || (dirs && rgb && acc && normal && mask && depth)), PSN_ERR_ARG,
              "psn_render_unisurf: null argument");
  if (N == 0) return PSN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = prm->steps_in + prm->steps_out;
  Workspace w(ws, ws_bytes);
  float* far = w.take<float>(N);
  int rc;
  {
    Workspace wm = w;  // the march region is dead after the surface search: the sample buffers alias it
    if ((rc = raymarch_impl(geo, origin, dirs, N, prm->near_, prm->radius, prm->march_steps, prm->secant_steps, prm->tau,
                            depth, far, wm, precision, st)))
      return rc;
  }
  // NOTE: kernels run in stream order, so aliasing the march scratch with the sample buffers is safe.
  float* sdepth = sample_depth ? sample_depth : w.take<float>((size_t)N * S);
  float* rgb_s = w.take<float>((size_t)N * S * 3);
  float* alpha = w.take<float>((size_t)N * S);
  SurfList sl;
  sl.count = w.take<int>(64);
  sl.ray = w.take<int>(N);
  sl.depth = w.take<float>(N);
  float* grads = w.take<float>((size_t)N * 3);
  void* stash = w.take<char>(stash_bytes_any());
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_render_unisurf: workspace too small (need %zu bytes, have %lld)", w.used,
              (long long)ws_bytes);
  PSN_CUDA_CHECK(cudaMemsetAsync(sl.count, 0, sizeof(int), st));
  if ((rc = launch_sample_plan(depth, far, N, *prm, noise, sdepth, mask, sl, st))) return rc;
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_RAY_DEPTH;
  gen.dirs = dirs;
  gen.depth = sdepth;
  gen.S = S;
  gen.o[0] = origin[0]; gen.o[1] = origin[1]; gen.o[2] = origin[2];
  if ((rc = radiance_any(geo, app, gen, (long long)N * S, rgb_s, alpha, stash, precision, st))) return rc;
  if ((rc = launch_composite(rgb_s, alpha, N, S, prm->white_background, rgb, acc, st))) return rc;
  // surface normals g/(|g|+1e-5) on object rays, zeros elsewhere (rendering.py:199-211)
  PSN_CUDA_CHECK(cudaMemsetAsync(normal, 0, (size_t)N * 3 * sizeof(float), st));
  PointGen g2;
  memset(&g2, 0, sizeof(g2));
  g2.kind = GEN_INDEXED_DEPTH;
  g2.dirs = dirs;
  g2.index = sl.ray;
  g2.depth = sl.depth;
  g2.o[0] = origin[0]; g2.o[1] = origin[1]; g2.o[2] = origin[2];
  if ((rc = gradient_any(geo, g2, 0, sl.count, grads, stash, normal_precision(precision), st))) return rc;
  return launch_scatter_normals(grads, sl, normal, N, st);
}

extern "C" int psn_shadow_visibility(const psn_mlp* geo, const float* surf, const float* lights, int64_t Ns, int L, float lnear,
                                     float lfar, int n_steps, float box, float* vis, void* ws, int64_t ws_bytes, int precision,
                                     void* stream) {
  PSN_REQUIRE(geo && (Ns == 0 || L == 0 || (surf && lights && vis)), PSN_ERR_ARG, "psn_shadow_visibility: null argument");
  PSN_REQUIRE(n_steps >= 2, PSN_ERR_ARG, "psn_shadow_visibility: n_steps=%d", n_steps);
  if (Ns == 0 || L == 0) return PSN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  long long lights_per_chunk = kShadowChunkPairs / Ns;
  if (lights_per_chunk < 1) lights_per_chunk = 1;
  if (lights_per_chunk > L) lights_per_chunk = L;
  const long long chunk_pairs = lights_per_chunk * Ns;
  Workspace w(ws, ws_bytes);
  // ws[0..3] = list length of the current chunk, ws[8..15] = in-box samples evaluated by this call (64-bit), ws[16..19] = 1 when the
  // box-culled pass ran: readable by the caller after the call (psnerf_b200.engine.shadow_visibility(..., return_stats=True))
  unsigned* counters = w.take<unsigned>(64);
  float* occ = w.take<float>((size_t)chunk_pairs * n_steps);  // culled: packed entries, overwritten in place by their alpha
  unsigned long long* meta = w.take<unsigned long long>((size_t)chunk_pairs);
  unsigned* off_b = w.take<unsigned>((size_t)chunk_pairs);
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_shadow_visibility: workspace too small (need %zu bytes, have %lld)", w.used,
              (long long)ws_bytes);
  // The box-culled pass needs the step index in SHADOW_LIST_STEP_BITS bits and the pair index in the rest of an entry.
  const bool culled = n_steps <= (1 << SHADOW_LIST_STEP_BITS) && chunk_pairs < (1LL << (32 - SHADOW_LIST_STEP_BITS)) &&
                      !shadow_unculled_requested();
  PSN_CUDA_CHECK(cudaMemsetAsync(counters, 0, 64 * sizeof(unsigned), st));
  ShadowList sl;
  sl.total = counters;
  sl.evaluated = reinterpret_cast<unsigned long long*>(counters + 2);
  sl.meta = meta;
  sl.off_b = off_b;
  sl.entry = reinterpret_cast<unsigned*>(occ);
  sl.lead = n_steps < kShadowLead ? n_steps : kShadowLead;
  sl.cap_a = (unsigned)(chunk_pairs * sl.lead);  // list A never exceeds lead entries per pair; B takes the rest of the buffer
  for (long long l0 = 0; l0 < L; l0 += lights_per_chunk) {
    const long long nl = (L - l0 < lights_per_chunk) ? (L - l0) : lights_per_chunk;
    PointGen gen;
    memset(&gen, 0, sizeof(gen));
    gen.kind = GEN_SHADOW;
    gen.surf = surf;
    gen.lights = lights + l0 * 3;
    gen.Ns = Ns;
    gen.S = n_steps;
    gen.lnear = lnear;
    gen.lfar = lfar;
    int rc;
    if (culled) {
      // plan A -> MLP on the first `lead` in-box steps of every pair -> plan B -> MLP on the remaining steps of the pairs that are
      // still alive -> transmittance over all steps; the row counts of both MLP launches are read on the device
      PSN_CUDA_CHECK(cudaMemsetAsync(sl.total, 0, 2 * sizeof(unsigned), st));
      if ((rc = launch_shadow_plan(surf, gen.lights, Ns, nl * Ns, n_steps, lnear, lfar, box, sl, st))) return rc;
      gen.kind = GEN_SHADOW_LIST;
      for (int which = 0; which < 2; ++which) {
        float* occ_w = occ + (which ? sl.cap_a : 0u);
        gen.index = reinterpret_cast<const int*>(occ_w);
        const int* rows = reinterpret_cast<const int*>(sl.total + which);
        {
          ProfScope prof(PSN_PROF_SHADOW, which ? 0 : nl * Ns * n_steps, st);
          if (prec_is_tc(precision)) rc = tc_occupancy(geo, gen, 0, rows, PSN_OUT_ALPHA, occ_w, st);
          else rc = simt_occupancy(geo, gen, 0, rows, PSN_OUT_ALPHA, occ_w, 0, st);
          if (rc) return rc;
        }
        if (which == 0 && (rc = launch_shadow_plan_b(occ, nl * Ns, sl, st))) return rc;
      }
      if ((rc = launch_shadow_composite_list(occ, sl, surf, gen.lights, Ns, nl * Ns, n_steps, lnear, lfar, box, vis + l0 * Ns, st)))
        return rc;
      continue;
    }
    if (prec_is_tc(precision) && n_steps == 128) {
      ProfScope prof(PSN_PROF_SHADOW, nl * Ns * n_steps, st);
      rc = tc_shadow(geo, gen, nl * Ns, box, vis + l0 * Ns, st);  // fused march + transmittance, no HBM round trip
      if (rc) return rc;
      continue;
    }
    if ((rc = occupancy_any(geo, gen, nl * Ns * n_steps, nullptr, PSN_OUT_ALPHA, occ, precision, st))) return rc;
    if ((rc = launch_shadow_composite(occ, surf, lights + l0 * 3, Ns, nl * Ns, n_steps, lnear, lfar, box, vis + l0 * Ns, st)))
      return rc;
  }
  return PSN_OK;
}

extern "C" int psn_composite(const float* rgb_s, const float* alpha, int64_t N, int S, int white_background, float* rgb,
                             float* acc, void* stream) {
  PSN_REQUIRE(N == 0 || (rgb_s && alpha && rgb && acc), PSN_ERR_ARG, "psn_composite: null argument");
  return launch_composite(rgb_s, alpha, N, S, white_background, rgb, acc, (cudaStream_t)stream);
}
