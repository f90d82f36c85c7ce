// Stage-2 train step, forward + backward (BASELINE config 5; stage2/trainer.py:394-410 around PSNetwork.forward,
// stage2/model/renderer.py:110-266 with train.light_vis_detach = train.vis_rgb_detach = True as in every shipped conf).
//
// What carries gradient (renderer.py:193-199, 211-231, 251-262): the three per-point nets (normal, albedo, SG weights) through
// the SG shading of all L lights, their jittered re-evaluation, the light directions and intensities, and visibility_net through
// the L' "vis-train" lights only.  The L-light visibility pass is detached: it runs on the inference kernels (tcgen05 or fp32).
// Gradient-carrying MLPs run as plain fp32 GEMMs with saved post-activations (rows <= 8 x 8192: < 2 % of the step's FLOPs);
// weight gradients are reduced with split-K atomics.  Hand-derived shading backward: see k_shade_bwd.
#include "simt_mlp.cuh"
#include "launch.cuh"
#include "internal.cuh"
#include "prof.cuh"
#include "train_gemm.cuh"

namespace psn {

// ---- element-wise pieces ------------------------------------------------------------------------------------------------
// out[r, col0 + j] = embed(x[src(r)])  with src(r) = (r / div) % mod   (div = 1, mod = M: identity; tiling of points / lights)
__global__ void k_embed(const float* __restrict__ x, long long rows, long long div, long long mod, int nf, float* __restrict__ out,
                        int ld, int col0) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const long long s = (r / div) % mod;
  const float v[3] = {x[s * 3], x[s * 3 + 1], x[s * 3 + 2]};
  float* o = out + r * ld + col0;
  o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
  for (int i = 0; i < nf; ++i) {
    const float f = (float)(1 << i);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sn, cs;
      sincosf(v[c] * f, &sn, &cs);
      o[3 + 6 * i + c] = sn;
      o[6 + 6 * i + c] = cs;
    }
  }
}
__global__ void k_copy_cols(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, long long rows, int ncols, int col0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ncols) return;
  const long long r = i / ncols;
  const int c = (int)(i - r * ncols);
  dst[r * ldd + col0 + c] = src[r * lds + c];
}
__global__ void k_normalize_fwd(const float* __restrict__ y, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = y[i * 3], b = y[i * 3 + 1], c = y[i * 3 + 2];
  const float nn = fmaxf(sqrtf(a * a + b * b + c * c), 1e-12f);
  out[i * 3] = a / nn; out[i * 3 + 1] = b / nn; out[i * 3 + 2] = c / nn;
}
// d y = (d n - n (n . d n)) / |y|   (F.normalize backward; eps branch ignored: |y| > 1e-12 in practice)
__global__ void k_normalize_bwd(const float* __restrict__ dn, const float* __restrict__ y, long long n, float* __restrict__ dy) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = y[i * 3], b = y[i * 3 + 1], c = y[i * 3 + 2];
  const float nn = fmaxf(sqrtf(a * a + b * b + c * c), 1e-12f);
  const float u[3] = {a / nn, b / nn, c / nn};
  const float g[3] = {dn[i * 3], dn[i * 3 + 1], dn[i * 3 + 2]};
  const float d = u[0] * g[0] + u[1] * g[1] + u[2] * g[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) dy[i * 3 + k] = (g[k] - u[k] * d) / nn;
}

// ---- shading backward -----------------------------------------------------------------------------------------------------
// Per surface point (one thread), loop over lights.  Forward (renderer.py:183-197, sgbasis.py:24-31):
//   h = (l+v)/|l+v| ; t = h.n - 1 ; D_k = exp(lam_k t) ; S_c = sum_k w[c,k] D_k ; spec_c = max(S_c,0) ; brdf_c = a_c + spec_c
//   cos = l.n ; u_c = brdf_c I_c cos visc ; rgb_c = clamp(u_c, 0, 1)          (visc = clamp(vis,0,1) is a constant: detached)
// Outputs: d a, d w (pre-relu mask applied by the caller), d n per point; d l, d I per light (block reduction + atomics).
struct ShadeBwdArgs {
  const float *g_rgb, *g_spec;  // [L,N,3] image shaped (g_spec may be null)
  const float *normal, *albedo, *weights, *view, *vis, *lights, *lobe, *intensity;
  const int* pix;
  float *d_albedo, *d_weights, *d_normal, *d_lights, *d_intensity;
  long long N, Ns;
  int L, nbasis, specular_rgb, nbt, intensity_kind;
  float intensity_scalar;
};
__global__ void __launch_bounds__(128)
k_shade_bwd(ShadeBwdArgs a) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = n < a.Ns;
  const long long s = live ? n : 0;
  const long long px = a.pix[s];
  const float nn[3] = {a.normal[s * 3], a.normal[s * 3 + 1], a.normal[s * 3 + 2]};
  const float vv[3] = {a.view[s * 3], a.view[s * 3 + 1], a.view[s * 3 + 2]};
  const float al[3] = {a.albedo[s * 3], a.albedo[s * 3 + 1], a.albedo[s * 3 + 2]};
  float w[27], dw[27];
  for (int k = 0; k < a.nbt; ++k) { w[k] = a.weights[s * a.nbt + k]; dw[k] = 0.f; }
  float da[3] = {0.f, 0.f, 0.f}, dn[3] = {0.f, 0.f, 0.f};
  __shared__ float red[128 / 32][6];
  for (int l = 0; l < a.L; ++l) {
    const float ll[3] = {a.lights[l * 3], a.lights[l * 3 + 1], a.lights[l * 3 + 2]};
    float dl[3] = {0.f, 0.f, 0.f}, dI[3] = {0.f, 0.f, 0.f};
    if (live) {
      const float hx = ll[0] + vv[0], hy = ll[1] + vv[1], hz = ll[2] + vv[2];
      const float hn = fmaxf(sqrtf(hx * hx + hy * hy + hz * hz), 1e-12f);
      const float h[3] = {hx / hn, hy / hn, hz / hn};
      const float t = (h[0] * nn[0] + h[1] * nn[1] + h[2] * nn[2]) - 1.f;
      float D[9], S[3] = {0.f, 0.f, 0.f};
      for (int k = 0; k < a.nbasis; ++k) {
        D[k] = expf(fmaxf(a.lobe[k], 0.f) * t);
        if (a.specular_rgb) { S[0] += w[k] * D[k]; S[1] += w[a.nbasis + k] * D[k]; S[2] += w[2 * a.nbasis + k] * D[k]; }
        else S[0] += w[k] * D[k];
      }
      if (!a.specular_rgb) S[1] = S[2] = S[0];
      const float cosv = ll[0] * nn[0] + ll[1] * nn[1] + ll[2] * nn[2];
      const float visc = a.vis ? fminf(fmaxf(a.vis[(long long)l * a.Ns + s], 0.f), 1.f) : 1.f;
      const long long o = ((long long)l * a.N + px) * 3;
      float dS[3], dcos = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float spec = fmaxf(S[c], 0.f);
        const float brdf = al[c] + spec;
        const float I = a.intensity_kind == 0 ? a.intensity_scalar : a.intensity_kind == 1 ? a.intensity[l] : a.intensity[l * 3 + c];
        const float u = brdf * I * cosv * visc;
        const float gu = (u > 0.f && u < 1.f) ? a.g_rgb[o + c] : 0.f;  // clamp(0,1) passes gradient strictly inside
        const float dbrdf = gu * I * cosv * visc;
        dI[c] = gu * brdf * cosv * visc;
        dcos += gu * brdf * I * visc;
        da[c] += dbrdf;
        float gs = dbrdf;                                              // d spec_c
        if (a.g_spec) gs += a.g_spec[o + c];
        dS[c] = S[c] > 0.f ? gs : 0.f;
      }
      float dt = 0.f;
      for (int k = 0; k < a.nbasis; ++k) {
        float dD;
        if (a.specular_rgb) {
          dw[k] += dS[0] * D[k]; dw[a.nbasis + k] += dS[1] * D[k]; dw[2 * a.nbasis + k] += dS[2] * D[k];
          dD = dS[0] * w[k] + dS[1] * w[a.nbasis + k] + dS[2] * w[2 * a.nbasis + k];
        } else {
          const float ds = dS[0] + dS[1] + dS[2];
          dw[k] += ds * D[k];
          dD = ds * w[k];
        }
        dt += dD * fmaxf(a.lobe[k], 0.f) * D[k];
      }
      // t = h.n - 1 ; cos = l.n
      float dh[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dh[c] = dt * nn[c];
        dn[c] += dt * h[c] + dcos * ll[c];
        dl[c] = dcos * nn[c];
      }
      const float hd = h[0] * dh[0] + h[1] * dh[1] + h[2] * dh[2];
#pragma unroll
      for (int c = 0; c < 3; ++c) dl[c] += (dh[c] - h[c] * hd) / hn;  // h = normalize(l + v)
    }
    // per-light reductions over the block
    float r6[6] = {dl[0], dl[1], dl[2], dI[0], dI[1], dI[2]};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) r6[q] += __shfl_xor_sync(0xffffffffu, r6[q], off);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int q = 0; q < 6; ++q) red[threadIdx.x >> 5][q] = r6[q];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      float t = 0.f;
      for (int wv = 0; wv < 128 / 32; ++wv) t += red[wv][threadIdx.x];
      if (threadIdx.x < 3) {
        if (a.d_lights) atomicAdd(&a.d_lights[l * 3 + threadIdx.x], t);
      } else if (a.d_intensity) {
        const int c = threadIdx.x - 3;
        if (a.intensity_kind == 2) atomicAdd(&a.d_intensity[l * 3 + c], t);
        else if (a.intensity_kind == 1) atomicAdd(&a.d_intensity[l], t);
        else atomicAdd(&a.d_intensity[0], t);
      }
    }
    __syncthreads();
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { a.d_albedo[s * 3 + c] += da[c]; a.d_normal[s * 3 + c] += dn[c]; }
    for (int k = 0; k < a.nbt; ++k) a.d_weights[s * a.nbt + k] += dw[k];
  }
}
// gather image-shaped gradients of per-pixel outputs to per-slot buffers: dst[s, c] = src[pix[s], c]  (or 0 when src is null)
__global__ void k_gather_rows(const float* __restrict__ src, const int* __restrict__ pix, long long Ns, int ncols, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ns * ncols) return;
  const long long s = i / ncols;
  const int c = (int)(i - s * ncols);
  dst[i] = src ? src[(long long)pix[s] * ncols + c] : 0.f;
}
__global__ void k_relu_inplace(float* x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], 0.f);
}
__global__ void k_sum3(const float* __restrict__ src, long long n, float* __restrict__ dst) {  // dst[i] = src[i,0]+src[i,1]+src[i,2]
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i * 3] + src[i * 3 + 1] + src[i * 3 + 2];
}

// ---- host orchestration ------------------------------------------------------------------------------------------------
// Saved activations of one MLP evaluation: act[l] = output of layer l ([rows, ld[l]], post activation, with the skip input
// appended when l == skip); x0 = network input [rows, in0].
struct MlpTape {
  const float* x0;
  float* act[kMaxLayers];
  int ld[kMaxLayers];
  long long rows;
};

// dz_last: gradient w.r.t. the PRE-activation of the last layer, [rows, out_last] (ld = out_last).  s0 / s1: scratch [rows, 512].
// dW[l] += dz^T x_l ; db[l] += colsum(dz) ; dx = dz W[l] (-> s1) ; dz_{l-1} = dx[:, :out_{l-1}] * relu'(act_{l-1}) (-> s0).
static int mlp_backward(const psn_train_net* n, const MlpTape* t, const float* dz_last, float* s0, float* s1, cudaStream_t st) {
  const long long rows = t->rows;
  if (rows == 0) return PSN_OK;
  const float* dz = dz_last;
  int lddz = n->out_dims[n->n_layers - 1];
  for (int l = n->n_layers - 1; l >= 0; --l) {
    const float* xin = (l == 0) ? t->x0 : t->act[l - 1];
    const int ldx = (l == 0) ? n->in_dims[0] : t->ld[l - 1];
    int rc;
    if ((rc = gemm(2, dz, lddz, xin, ldx, n->dW[l], n->in_dims[l], nullptr, n->out_dims[l], n->in_dims[l], rows, 0, st))) return rc;
    count_launch();
    colsum(dz, lddz, rows, n->out_dims[l], n->db[l], st);
    if (l == 0) break;
    if ((rc = gemm(1, dz, lddz, n->W[l], n->in_dims[l], s1, n->in_dims[l], nullptr, rows, n->in_dims[l], n->out_dims[l], 0, st))) return rc;
    // only the first out_{l-1} columns flow on (the appended skip input carries no gradient: points / detached lights)
    count_launch();
    k_act_bwd<<<nblk(rows * n->out_dims[l - 1]), 256, 0, st>>>(s1, n->in_dims[l], t->act[l - 1], t->ld[l - 1], s0, n->out_dims[l - 1], rows,
                                                                n->out_dims[l - 1], 2);
    dz = s0;
    lddz = n->out_dims[l - 1];
  }
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

// ---- tape layout (identical carve in forward and backward) -----------------------------------------------------------------
struct S2Tape {
  float *E_p, *E_n, *E_j, *X_v;               // embeddings: points (brdf freqs), points (normal freqs), jittered points, vis-train pairs
  MlpTape normal, albedo, rough, albedo_j, rough_j, vis_t;
  float *n_s, *a_s, *w_s, *v_raw;             // per slot: unit normal, albedo, relu'd SG weights; raw L-light visibility [L,Ns]
  size_t floats;
};
static size_t al64(size_t f) { return (f + 63) / 64 * 64; }

static void tape_mlp(const psn_train_net* n, long long rows, MlpTape* t, float*& cur) {
  t->rows = rows;
  for (int l = 0; l < n->n_layers; ++l) {
    t->ld[l] = n->out_dims[l] + (l == n->skip ? n->in_dims[0] : 0);
    t->act[l] = cur;
    cur += al64((size_t)rows * t->ld[l]);
  }
}
static void carve_tape(float* base, const psn_train_net* nn, const psn_train_net* an, const psn_train_net* rn, const psn_train_net* vn,
                       long long Ns, int L, int Lt, int nbt, S2Tape* t) {
  float* cur = base;
  const int e = an->in_dims[0];
  const int en = nn ? nn->in_dims[0] : e;
  t->E_p = cur; cur += al64((size_t)Ns * e);
  t->E_n = cur; cur += al64((size_t)Ns * en);
  t->E_j = cur; cur += al64((size_t)Ns * e);
  t->X_v = cur; cur += al64((size_t)Lt * Ns * (vn ? vn->in_dims[0] : 1));
  if (nn) tape_mlp(nn, Ns, &t->normal, cur);
  tape_mlp(an, Ns, &t->albedo, cur);
  tape_mlp(rn, Ns, &t->rough, cur);
  tape_mlp(an, Ns, &t->albedo_j, cur);
  tape_mlp(rn, Ns, &t->rough_j, cur);
  if (vn) tape_mlp(vn, (long long)Lt * Ns, &t->vis_t, cur);
  t->n_s = cur; cur += al64((size_t)Ns * 3);
  t->a_s = cur; cur += al64((size_t)Ns * 3);
  t->w_s = cur; cur += al64((size_t)Ns * nbt);
  t->v_raw = cur; cur += al64((size_t)Ns * (L > 0 ? L : 1));
  t->floats = (size_t)(cur - base);
}

// run one MLP over its (already carved) tape
static int mlp_forward_tape(const psn_train_net* n, const float* x0, MlpTape* t, cudaStream_t st) {
  t->x0 = x0;
  const long long rows = t->rows;
  if (rows == 0) return PSN_OK;
  const float* in = x0;
  int ld_in = n->in_dims[0];
  for (int l = 0; l < n->n_layers; ++l) {
    const bool last = (l == n->n_layers - 1);
    const int epi = last ? (n->final_act == 1 ? 3 : 1) : 2;
    int rc = gemm(0, in, ld_in, n->W[l], n->in_dims[l], t->act[l], t->ld[l], n->b[l], rows, n->out_dims[l], n->in_dims[l], epi, st);
    if (rc) return rc;
    if (l == n->skip) {
      count_launch();
      k_copy_cols<<<nblk(rows * n->in_dims[0]), 256, 0, st>>>(x0, n->in_dims[0], t->act[l], t->ld[l], rows, n->in_dims[0], n->out_dims[l]);
    }
    in = t->act[l];
    ld_in = t->ld[l];
  }
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

static int check_net(const psn_train_net* n, const char* what, bool need_grads) {
  PSN_REQUIRE(n && n->n_layers >= 1 && n->n_layers < kMaxLayers && n->in_dims && n->out_dims && n->W && n->b, PSN_ERR_ARG,
              "%s: bad psn_train_net", what);
  PSN_REQUIRE(!need_grads || (n->dW && n->db), PSN_ERR_ARG, "%s: gradient buffers missing", what);
  PSN_REQUIRE(n->skip < n->n_layers - 1, PSN_ERR_SHAPE, "%s: skip after the last layer is unsupported", what);
  for (int l = 1; l < n->n_layers; ++l)
    PSN_REQUIRE(n->in_dims[l] == n->out_dims[l - 1] + (l - 1 == n->skip ? n->in_dims[0] : 0) && n->in_dims[l] <= 512, PSN_ERR_SHAPE,
                "%s: layer %d input width %d inconsistent", what, l, n->in_dims[l]);
  return PSN_OK;
}

int s2_visibility_simt(const psn_mlp* vis_net, int nf, const float* pts, long long Ns, const float* lights, int L, float* vis,
                       cudaStream_t st);
int tc_s2_visibility(const psn_mlp* net, int nf, const float* pts, long long Ns, const float* lights, int L, float* vis, void* ws,
                     size_t ws_bytes, cudaStream_t st);
int s2_shade_images(const float* n_s, const float* a_s, const float* w_s, const float* view, const float* v_raw, const float* lights,
                    const float* lobe, const float* intensity, const psn_shade_params* prm, const int32_t* pix, long long Ns, long long N,
                    int L, float* rgb, float* spec, float* vis, float* normal, float* albedo, float* sgw, int write_normal, int* sop,
                    cudaStream_t st);  // stage2_simt.cu

}  // namespace psn

using namespace psn;

extern "C" int64_t psn_s2_train_tape_bytes(const psn_train_net* nn, const psn_train_net* an, const psn_train_net* rn,
                                           const psn_train_net* vn, int64_t Ns, int L, int Lt) {
  if (!an || !rn) return -1;
  S2Tape t;
  carve_tape(nullptr, nn, an, rn, vn, Ns, L, Lt, rn->out_dims[rn->n_layers - 1], &t);
  return (int64_t)(t.floats * sizeof(float) + 1024);
}

extern "C" int psn_s2_train_forward(const psn_train_net* nn, const psn_train_net* an, const psn_train_net* rn, const psn_train_net* vn,
                                    const psn_mlp* vis_packed, const float* lobe, const psn_shade_params* prm, const float* pts,
                                    const float* view, const int32_t* pix, int64_t Ns, int64_t N, const float* lights, int L,
                                    const float* intensity, const float* jitter_pts, const float* lights_vt, int Lt, float* rgb,
                                    float* spec, float* vis, float* normal, float* albedo, float* sgw, float* albedo_j, float* weights_j,
                                    float* vis_train, void* tape, int64_t tape_bytes, void* ws, int64_t ws_bytes, int precision,
                                    void* stream) {
  PSN_REQUIRE(prm && lobe && lights && rgb && spec && albedo && sgw && tape, PSN_ERR_ARG, "psn_s2_train_forward: null argument");
  int rc;
  if ((rc = check_net(an, "albedo_net", false)) || (rc = check_net(rn, "rough_net", false))) return rc;
  if (nn && (rc = check_net(nn, "normal_net", false))) return rc;
  if (vn && (rc = check_net(vn, "visibility_net", false))) return rc;
  PSN_REQUIRE(nn && normal, PSN_ERR_ARG, "psn_s2_train_forward: the train step needs normal_net (train.normal_mlp)");
  PSN_REQUIRE(!vn || (vis_packed && vis), PSN_ERR_ARG, "psn_s2_train_forward: visibility needs the packed net and the vis output");
  PSN_REQUIRE(!lights_vt || (vn && vis_train && Lt > 0), PSN_ERR_ARG, "psn_s2_train_forward: vis-train lights need visibility_net");
  cudaStream_t st = (cudaStream_t)stream;
  PSN_REQUIRE(prm->render_model == 0, PSN_ERR_SHAPE, "stage-2 train step: only render_model = sgbasis is implemented");
  const int nbt = prm->specular_rgb ? 3 * prm->nbasis : prm->nbasis;
  PSN_REQUIRE(rn->out_dims[rn->n_layers - 1] == nbt && nbt <= 27, PSN_ERR_SHAPE, "rough_net output %d != %d", rn->out_dims[rn->n_layers - 1], nbt);
  S2Tape t;
  carve_tape((float*)tape, nn, an, rn, vn, Ns, L, lights_vt ? Lt : 0, nbt, &t);
  PSN_REQUIRE((int64_t)(t.floats * sizeof(float)) <= tape_bytes, PSN_ERR_WORKSPACE, "psn_s2_train_forward: tape too small (%zu > %lld)",
              t.floats * sizeof(float), (long long)tape_bytes);
  const int e = an->in_dims[0], nf = prm->n_freqs_xyz, nfn = prm->n_freqs_normal;
  PSN_REQUIRE(e == 3 + 6 * nf && nn->in_dims[0] == 3 + 6 * nfn && rn->in_dims[0] == e, PSN_ERR_SHAPE, "embedding widths do not match the nets");
  Workspace w(ws, ws_bytes);
  int* sop = w.take<int>((size_t)N + 4);
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_s2_train_forward: workspace too small");
  if (Ns > 0) {
    count_launch();
    k_embed<<<nblk(Ns), 256, 0, st>>>(pts, Ns, 1, Ns, nf, t.E_p, e, 0);
    count_launch();
    k_embed<<<nblk(Ns), 256, 0, st>>>(pts, Ns, 1, Ns, nfn, t.E_n, nn->in_dims[0], 0);
    if ((rc = mlp_forward_tape(nn, t.E_n, &t.normal, st))) return rc;
    count_launch();
    k_normalize_fwd<<<nblk(Ns), 256, 0, st>>>(t.normal.act[nn->n_layers - 1], Ns, t.n_s);
    if ((rc = mlp_forward_tape(an, t.E_p, &t.albedo, st))) return rc;
    if ((rc = mlp_forward_tape(rn, t.E_p, &t.rough, st))) return rc;
    PSN_CUDA_CHECK(cudaMemcpyAsync(t.a_s, t.albedo.act[an->n_layers - 1], (size_t)Ns * 3 * 4, cudaMemcpyDeviceToDevice, st));
    PSN_CUDA_CHECK(cudaMemcpyAsync(t.w_s, t.rough.act[rn->n_layers - 1], (size_t)Ns * nbt * 4, cudaMemcpyDeviceToDevice, st));
    count_launch();
    k_relu_inplace<<<nblk(Ns * nbt), 256, 0, st>>>(t.w_s, Ns * nbt);
    if (jitter_pts) {
      PSN_REQUIRE(albedo_j && weights_j, PSN_ERR_ARG, "psn_s2_train_forward: jitter outputs missing");
      count_launch();
      k_embed<<<nblk(Ns), 256, 0, st>>>(jitter_pts, Ns, 1, Ns, nf, t.E_j, e, 0);
      if ((rc = mlp_forward_tape(an, t.E_j, &t.albedo_j, st))) return rc;
      if ((rc = mlp_forward_tape(rn, t.E_j, &t.rough_j, st))) return rc;
      PSN_CUDA_CHECK(cudaMemcpyAsync(albedo_j, t.albedo_j.act[an->n_layers - 1], (size_t)Ns * 3 * 4, cudaMemcpyDeviceToDevice, st));
      PSN_CUDA_CHECK(cudaMemcpyAsync(weights_j, t.rough_j.act[rn->n_layers - 1], (size_t)Ns * nbt * 4, cudaMemcpyDeviceToDevice, st));
      count_launch();
      k_relu_inplace<<<nblk(Ns * nbt), 256, 0, st>>>(weights_j, Ns * nbt);
    }
    if (vn) {  // detached L-light pass on the inference kernels
      ProfScope prof(PSN_PROF_S2_VIS, (long long)Ns * L, st);
      if (prec_is_tc(precision)) {
        const size_t off = (w.used + 255) / 256 * 256;
        rc = tc_s2_visibility(vis_packed, nf, pts, Ns, lights, L, t.v_raw, (char*)ws + off, (size_t)ws_bytes > off ? (size_t)ws_bytes - off : 0, st);
      } else {
        rc = s2_visibility_simt(vis_packed, nf, pts, Ns, lights, L, t.v_raw, st);
      }
      if (rc) return rc;
    }
    if (lights_vt) {
      const long long rows = (long long)Lt * Ns;
      const int ev = vn->in_dims[0];
      PSN_REQUIRE(ev == 2 * e, PSN_ERR_SHAPE, "visibility_net input %d != 2 x %d", ev, e);
      count_launch();
      k_embed<<<nblk(rows), 256, 0, st>>>(pts, rows, 1, Ns, nf, t.X_v, ev, 0);          // row r = t*Ns + n -> point n
      count_launch();
      k_embed<<<nblk(rows), 256, 0, st>>>(lights_vt, rows, Ns, Lt, nf, t.X_v, ev, e);    //                 -> light t
      if ((rc = mlp_forward_tape(vn, t.X_v, &t.vis_t, st))) return rc;
      PSN_CUDA_CHECK(cudaMemcpyAsync(vis_train, t.vis_t.act[vn->n_layers - 1], (size_t)rows * 4, cudaMemcpyDeviceToDevice, st));
    }
  }
  return s2_shade_images(t.n_s, t.a_s, t.w_s, view, vn ? t.v_raw : nullptr, lights, lobe, intensity, prm, pix, Ns, N, L, rgb, spec, vis,
                         normal, albedo, sgw, 1, sop, st);
}

extern "C" int psn_s2_train_backward(const psn_train_net* nn, const psn_train_net* an, const psn_train_net* rn, const psn_train_net* vn,
                                     const float* lobe, const psn_shade_params* prm, const float* view, const int32_t* pix, int64_t Ns,
                                     int64_t N, const float* lights, int L, const float* intensity, int Lt, const float* g_rgb,
                                     const float* g_spec, const float* g_normal, const float* g_albedo, const float* g_sgw,
                                     const float* g_albedo_j, const float* g_weights_j, const float* g_vis_train, float* d_lights,
                                     float* d_intensity, void* tape, int64_t tape_bytes, void* ws, int64_t ws_bytes, void* stream) {
  PSN_REQUIRE(prm && lobe && lights && tape, PSN_ERR_ARG, "psn_s2_train_backward: null argument");
  int rc;
  if ((rc = check_net(an, "albedo_net", true)) || (rc = check_net(rn, "rough_net", true)) || (rc = check_net(nn, "normal_net", true))) return rc;
  if (vn && (rc = check_net(vn, "visibility_net", g_vis_train != nullptr))) return rc;
  if (Ns == 0) return PSN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PSN_REQUIRE(prm->render_model == 0, PSN_ERR_SHAPE, "stage-2 train step: only render_model = sgbasis is implemented");
  const int nbt = prm->specular_rgb ? 3 * prm->nbasis : prm->nbasis;
  S2Tape t;
  carve_tape((float*)tape, nn, an, rn, vn, Ns, L, g_vis_train ? Lt : (Lt > 0 ? Lt : 0), nbt, &t);
  PSN_REQUIRE((int64_t)(t.floats * sizeof(float)) <= tape_bytes, PSN_ERR_WORKSPACE, "psn_s2_train_backward: tape size mismatch");
  // re-attach the network inputs (carve_tape only lays out the buffers)
  t.normal.x0 = t.E_n; t.albedo.x0 = t.E_p; t.rough.x0 = t.E_p; t.albedo_j.x0 = t.E_j; t.rough_j.x0 = t.E_j; t.vis_t.x0 = t.X_v;
  const long long rows_v = (long long)Lt * Ns;
  const long long max_rows = (g_vis_train && rows_v > Ns) ? rows_v : Ns;
  Workspace w(ws, ws_bytes);
  float* s0 = w.take<float>((size_t)max_rows * 512);
  float* s1 = w.take<float>((size_t)max_rows * 512);
  float* d_a = w.take<float>((size_t)Ns * 3 + 4);
  float* d_w = w.take<float>((size_t)Ns * nbt + 4);
  float* d_n = w.take<float>((size_t)Ns * 3 + 4);
  float* dz = w.take<float>((size_t)max_rows * 32 + 4);
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_s2_train_backward: workspace too small (need %zu bytes, have %lld)", w.used, (long long)ws_bytes);
  // (a) visibility_net through the vis-train lights (renderer.py:251-262): dz_last = sum over the 3 expanded channels, done by the caller
  if (g_vis_train && vn) {
    if ((rc = mlp_backward(vn, &t.vis_t, g_vis_train, s0, s1, st))) return rc;
  }
  // (b) shading backward: per-point d albedo / d weights / d normal start from the direct output gradients
  count_launch();
  k_gather_rows<<<nblk(Ns * 3), 256, 0, st>>>(g_albedo, pix, Ns, 3, d_a);
  count_launch();
  k_gather_rows<<<nblk(Ns * nbt), 256, 0, st>>>(g_sgw, pix, Ns, nbt, d_w);
  count_launch();
  k_gather_rows<<<nblk(Ns * 3), 256, 0, st>>>(g_normal, pix, Ns, 3, d_n);
  if (g_rgb || g_spec) {
    ShadeBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.g_rgb = g_rgb; a.g_spec = g_spec;
    a.normal = t.n_s; a.albedo = t.a_s; a.weights = t.w_s; a.view = view; a.vis = vn ? t.v_raw : nullptr; a.lights = lights; a.lobe = lobe;
    a.intensity = intensity; a.pix = pix;
    a.d_albedo = d_a; a.d_weights = d_w; a.d_normal = d_n; a.d_lights = d_lights; a.d_intensity = d_intensity;
    a.N = N; a.Ns = Ns; a.L = L; a.nbasis = prm->nbasis; a.specular_rgb = prm->specular_rgb; a.nbt = nbt;
    a.intensity_kind = prm->intensity_kind; a.intensity_scalar = prm->intensity;
    PSN_REQUIRE(g_rgb, PSN_ERR_ARG, "psn_s2_train_backward: g_spec without g_rgb is unsupported");
    count_launch();
    k_shade_bwd<<<nblk(Ns, 128), 128, 0, st>>>(a);
  }
  // (c) normal net: through F.normalize
  count_launch();
  k_normalize_bwd<<<nblk(Ns), 256, 0, st>>>(d_n, t.normal.act[nn->n_layers - 1], Ns, dz);
  if ((rc = mlp_backward(nn, &t.normal, dz, s0, s1, st))) return rc;
  // (d) albedo net (sigmoid output), main and jittered evaluation
  count_launch();
  k_act_bwd<<<nblk(Ns * 3), 256, 0, st>>>(d_a, 3, t.albedo.act[an->n_layers - 1], 3, dz, 3, Ns, 3, 3);
  if ((rc = mlp_backward(an, &t.albedo, dz, s0, s1, st))) return rc;
  if (g_albedo_j) {
    count_launch();
    k_act_bwd<<<nblk(Ns * 3), 256, 0, st>>>(g_albedo_j, 3, t.albedo_j.act[an->n_layers - 1], 3, dz, 3, Ns, 3, 3);
    if ((rc = mlp_backward(an, &t.albedo_j, dz, s0, s1, st))) return rc;
  }
  // (e) rough net: weights = relu(raw output)
  count_launch();
  k_act_bwd<<<nblk(Ns * nbt), 256, 0, st>>>(d_w, nbt, t.rough.act[rn->n_layers - 1], nbt, dz, nbt, Ns, nbt, 2);
  if ((rc = mlp_backward(rn, &t.rough, dz, s0, s1, st))) return rc;
  if (g_weights_j) {
    count_launch();
    k_act_bwd<<<nblk(Ns * nbt), 256, 0, st>>>(g_weights_j, nbt, t.rough_j.act[rn->n_layers - 1], nbt, dz, nbt, Ns, nbt, 2);
    if ((rc = mlp_backward(rn, &t.rough_j, dz, s0, s1, st))) return rc;
  }
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
