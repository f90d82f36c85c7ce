// Stage-1 field, differentiable: forward + hand-derived backward of NeuralNetwork.forward(p, ray_d, return_addocc=True)
// under autograd (stage1/model/network.py:85-136), the inner step of the stage-1 train loop
// (stage1/model/training.py:141-198 -> rendering.py:50-226 with eval_=False).
//
// Per sample the forward computes   z_l = x_l W_l^T + b_l,  h_l = softplus_100(z_l)   (l = 0..7; x_0 = pe(p / rescale),
// x_skip = cat[h, pe] / sqrt2),   out = h_7 W_8^T + b_8  (logit = out[0], feature = out[1:]),   the analytic normal
//     a'_8 = W_8[0,:] ;  dz_l = a'_{l+1} * s_l  (s_l = sigmoid(100 z_l)) ;  a_l = dz_l W_l ;  g = J_pe^T a_0 / rescale
// and the appearance MLP on u = [p, pe(view / |view|), g, feature].  The reference feeds g into the appearance MLP with
// create_graph=True (network.py:117,130-132), so the loss reaches the geometry weights through THREE routes: the logit
// (alpha), the feature vector and g.  The third one is a double backward through the softplus stack; here it is the
// hand-written reverse-mode sweep over the a-chain ("second-order pass" below):
//     abar_0 = J_pe gbar / rescale ;  for l = 0..7:  Wbar_l += dz_l^T abar_l ;  dzbar = abar_l W_l^T ;
//     zbar2_l = dzbar * a'_{l+1} * 100 s_l (1 - s_l) ;  abar'_{l+1} = dzbar * s_l ;  Wbar_8[0,:] += colsum(abar'_8)
// followed by the ordinary backward of the forward stack with zbar_l = hbar_l * s_l + zbar2_l.
// Everything is plain fp32 (k_gemm, train_gemm.cuh) with the activations of all M samples saved on a caller-owned tape
// (30.5 KB per sample); weight gradients are reduced with split-K atomics.  Weight normalisation (W = g v / |v|) stays with
// the caller: the entry points take and return EFFECTIVE weights / gradients.
#include "launch.cuh"
#include "internal.cuh"
#include "train_gemm.cuh"

namespace psn {

#define PSN_S1_INV_SQRT2 0.70710678118654752440f

// ---- element-wise kernels ------------------------------------------------------------------------------------------------
// pe[m, :] = [x, sin(2^0 x), cos(2^0 x), ...], x = p / rescale (network.py:141-150)
__global__ void k_s1_pe(const float* __restrict__ p, long long M, int octaves, float rescale, float* __restrict__ pe, int ld) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float* o = pe + m * ld;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float x = p[m * 3 + c] / rescale;
    o[c] = x;
    for (int i = 0; i < octaves; ++i) {
      float sn, cs;
      sincosf((float)(1 << i) * x, &sn, &cs);
      o[3 + 6 * i + c] = sn;
      o[6 + 6 * i + c] = cs;
    }
  }
}
// h (in: pre-activation z, out: softplus_100(z) with PyTorch's threshold 20), s = sigmoid(100 z)
__global__ void k_s1_softplus(float* __restrict__ h, float* __restrict__ s, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float z = h[i], t = 100.f * z;
  h[i] = t > 20.f ? z : log1pf(expf(t)) / 100.f;
  s[i] = 1.f / (1.f + expf(-t));
}
// x4[m, :] = cat[h[m, :nh], pe[m, :npe]] / sqrt2 (network.py:90-91)
__global__ void k_s1_skip_cat(const float* __restrict__ h, int nh, const float* __restrict__ pe, int npe, long long M, float* __restrict__ x) {
  const int w = nh + npe;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * w) return;
  const long long m = i / w;
  const int c = (int)(i - m * w);
  x[i] = (c < nh ? h[m * nh + c] : pe[m * npe + (c - nh)]) * PSN_S1_INV_SQRT2;
}
__global__ void k_s1_bcast_row(const float* __restrict__ row, int w, long long M, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M * w) dst[i] = row[i % w];
}
__global__ void k_s1_mul(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ c) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = a[i] * b[i];
}
// reverse pass at the skip layer: a [M, nh + npe] (gradient w.r.t. x_skip) -> ap_prev [M, nh] = a[:, :nh] / sqrt2,
// gpe [M, npe] = a[:, nh:] / sqrt2 (assigned: the skip layer is visited before layer 0)
__global__ void k_s1_skip_split(const float* __restrict__ a, int nh, int npe, long long M, float* __restrict__ ap_prev, float* __restrict__ gpe) {
  const int w = nh + npe;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * w) return;
  const long long m = i / w;
  const int c = (int)(i - m * w);
  const float v = a[i] * PSN_S1_INV_SQRT2;
  if (c < nh) ap_prev[m * nh + c] = v;
  else gpe[m * npe + (c - nh)] = v;
}
__global__ void k_s1_axpy(const float* __restrict__ x, long long n, int accumulate, float* __restrict__ y) {  // y (+)= x
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = accumulate ? y[i] + x[i] : x[i];
}
// g = J_pe^T gpe / rescale with d pe / d x = [1, f cos(f x), -f sin(f x)] read back from the saved encoding
__global__ void k_s1_pe_jt(const float* __restrict__ gpe, const float* __restrict__ pe, int ld, long long M, int octaves, float rescale,
                           float* __restrict__ g, float* __restrict__ g2, int ld2) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* q = gpe + m * ld;
  const float* e = pe + m * ld;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = q[c];
    for (int i = 0; i < octaves; ++i) {
      const float f = (float)(1 << i);
      acc += f * e[6 + 6 * i + c] * q[3 + 6 * i + c] - f * e[3 + 6 * i + c] * q[6 + 6 * i + c];
    }
    acc /= rescale;
    g[m * 3 + c] = acc;
    if (g2) g2[m * ld2 + c] = acc;
  }
}
// gpe_bar = J_pe g_bar / rescale
__global__ void k_s1_pe_j(const float* __restrict__ gbar, const float* __restrict__ pe, int ld, long long M, int octaves, float rescale,
                          float* __restrict__ gpe_bar) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* e = pe + m * ld;
  float* o = gpe_bar + m * ld;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float gb = gbar[m * 3 + c] / rescale;
    o[c] = gb;
    for (int i = 0; i < octaves; ++i) {
      const float f = (float)(1 << i);
      o[3 + 6 * i + c] = f * e[6 + 6 * i + c] * gb;
      o[6 + 6 * i + c] = -f * e[3 + 6 * i + c] * gb;
    }
  }
}
// u[m, 0:3] = p, u[m, 3:3+pv] = pe(view / |view|) (network.py:98,127-129)
__global__ void k_s1_app_input(const float* __restrict__ p, const float* __restrict__ view, long long M, int octaves_view, float* __restrict__ u,
                               int ld) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float* o = u + m * ld;
  const float v0 = view[m * 3], v1 = view[m * 3 + 1], v2 = view[m * 3 + 2];
  const float nv = sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
  const float vn[3] = {v0 / nv, v1 / nv, v2 / nv};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = p[m * 3 + c];
    o[3 + c] = vn[c];
    for (int i = 0; i < octaves_view; ++i) {
      float sn, cs;
      sincosf((float)(1 << i) * vn[c], &sn, &cs);
      o[3 + 3 + 6 * i + c] = sn;
      o[3 + 6 + 6 * i + c] = cs;
    }
  }
}
__global__ void k_s1_copy_col(const float* __restrict__ src, int ld, int col, long long M, float* __restrict__ dst) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) dst[m] = src[m * ld + col];
}
// rgb = tanh(pre) * 0.5 + 0.5 ; t = tanh(pre) kept for the backward
__global__ void k_s1_rgb(const float* __restrict__ pre, long long n, float* __restrict__ t, float* __restrict__ rgb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = tanhf(pre[i]);
  t[i] = v;
  rgb[i] = v * 0.5f + 0.5f;
}
__global__ void k_s1_rgb_bwd(const float* __restrict__ g_rgb, const float* __restrict__ t, long long n, float* __restrict__ d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = g_rgb[i] * 0.5f * (1.f - t[i] * t[i]);
}
// gbar = g_grad (nullable) + ubar[:, col0 : col0+3] (nullable)
__global__ void k_s1_gbar(const float* __restrict__ g_grad, const float* __restrict__ ubar, int ld, int col0, long long M, float* __restrict__ gbar) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * 3) return;
  const long long m = i / 3;
  const int c = (int)(i - m * 3);
  gbar[i] = (g_grad ? g_grad[i] : 0.f) + (ubar ? ubar[m * ld + col0 + c] : 0.f);
}
// second-order pass, element-wise part: dzbar (in t2) ->  ap := zbar2 = dzbar * ap * 100 s (1 - s) ;  abar' = dzbar * s
__global__ void k_s1_second(const float* __restrict__ t2, float* __restrict__ ap, const float* __restrict__ s, long long n, float* __restrict__ abar) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = t2[i], sv = s[i];
  abar[i] = d * sv;
  ap[i] = d * ap[i] * (100.f * sv * (1.f - sv));
}
// abar_skip [M, nh + npe] = cat[abar' [M, nh], gpe_bar [M, npe]] / sqrt2
__global__ void k_s1_skip_merge(const float* __restrict__ abar_p, int nh, const float* __restrict__ gpe_bar, int npe, long long M,
                                float* __restrict__ out) {
  const int w = nh + npe;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * w) return;
  const long long m = i / w;
  const int c = (int)(i - m * w);
  out[i] = (c < nh ? abar_p[m * nh + c] : gpe_bar[m * npe + (c - nh)]) * PSN_S1_INV_SQRT2;
}
// zbar = hbar (nullable, ld ldh, scaled) * s + zbar2
__global__ void k_s1_zbar(const float* __restrict__ hbar, int ldh, float hscale, const float* __restrict__ s, const float* __restrict__ z2, int w,
                          long long M, float* __restrict__ zbar) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * w) return;
  const long long m = i / w;
  const int c = (int)(i - m * w);
  zbar[i] = (hbar ? hbar[m * ldh + c] * hscale * s[i] : 0.f) + z2[i];
}
// outbar [M, 1 + nf] = [g_logit (nullable), ubar[:, col0 : col0 + nf] (nullable)]
__global__ void k_s1_outbar(const float* __restrict__ g_logit, const float* __restrict__ ubar, int ld, int col0, int nf, long long M,
                            float* __restrict__ out) {
  const int w = 1 + nf;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * w) return;
  const long long m = i / w;
  const int c = (int)(i - m * w);
  out[i] = c == 0 ? (g_logit ? g_logit[m] : 0.f) : (ubar ? ubar[m * ld + col0 + c - 1] : 0.f);
}
__global__ void k_s1_colsum_row0(const float* __restrict__ x, int w, long long M, float* __restrict__ dst) {  // dst[c] += sum_m x[m, c]
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= w) return;
  const long long r0 = (long long)blockIdx.y * 1024, r1 = min(M, r0 + 1024);
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += x[r * w + c];
  atomicAdd(&dst[c], acc);
}

// ---- alpha compositing backward (rendering.py:196-197,214-216) --------------------------------------------------------------
// w_i = a_i T_i, T_i = prod_{j<i} om_j, om_j = 1 - a_j + 1e-6; rgb = sum w_i c_i (+ 1 - acc with a white background); acc = sum w_i.
// With wbar_i = d L / d w_i:   d a_i = T_i (wbar_i - R_i),   R_i = sum_{k>i} wbar_k a_k prod_{i<j<k} om_j = wbar_{i+1} a_{i+1} + om_{i+1} R_{i+1}
// (no divisions by the possibly tiny om_i).  One thread per ray; T_i is kept in a per-thread array (S <= 256).
constexpr int kCompositeMaxS = 256;
__global__ void k_composite_bwd(const float* __restrict__ rgb_s, const float* __restrict__ alpha, long long N, int S, int white,
                                const float* __restrict__ g_rgb, const float* __restrict__ g_acc, float* __restrict__ d_rgb_s,
                                float* __restrict__ d_alpha) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float gr[3] = {g_rgb ? g_rgb[r * 3] : 0.f, g_rgb ? g_rgb[r * 3 + 1] : 0.f, g_rgb ? g_rgb[r * 3 + 2] : 0.f};
  const float ga = g_acc ? g_acc[r] : 0.f;
  const float* a = alpha + r * S;
  const float* c = rgb_s + r * S * 3;
  float Tl[kCompositeMaxS];
  float T = 1.f;
  for (int i = 0; i < S; ++i) {  // d c_i = w_i g_rgb ; wbar_i parked in d_alpha
    Tl[i] = T;
    const float w = a[i] * T;
    float wb = ga;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      d_rgb_s[(r * S + i) * 3 + k] = w * gr[k];
      wb += (c[i * 3 + k] - (white ? 1.f : 0.f)) * gr[k];
    }
    d_alpha[r * S + i] = wb;
    T *= (1.f - a[i]) + 1e-6f;
  }
  float R = 0.f;
  for (int i = S - 1; i >= 0; --i) {
    const float wb = d_alpha[r * S + i];
    d_alpha[r * S + i] = Tl[i] * (wb - R);
    R = wb * a[i] + ((1.f - a[i]) + 1e-6f) * R;
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
struct S1Shape {
  int nh;            // hidden layers (8)
  int in[kMaxLayers], out[kMaxLayers];
  int pe_dim, pv_dim, skip, app_in, feat, has_app, nla;
  long long h_total;  // sum of out[l], l < nh
};
struct S1Tape {
  float *pe, *x_skip, *u, *t3;
  float *h[kMaxLayers], *s[kMaxLayers], *ap[kMaxLayers];
  float* y[kMaxLayers];
};

static int s1_shape(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, S1Shape* sh) {
  PSN_REQUIRE(geo && geo->n_layers >= 3 && geo->n_layers <= kMaxLayers, PSN_ERR_ARG, "stage-1 train: bad geo net");
  memset(sh, 0, sizeof(*sh));
  sh->nh = geo->n_layers - 1;
  sh->pe_dim = 3 + 6 * octaves;
  sh->pv_dim = 3 + 6 * octaves_view;
  sh->skip = geo->skip;
  PSN_REQUIRE(geo->in_dims[0] == sh->pe_dim, PSN_ERR_SHAPE, "stage-1 train: geo input width %d != encoding width %d", geo->in_dims[0],
              sh->pe_dim);
  for (int l = 0; l < geo->n_layers; ++l) {
    sh->in[l] = geo->in_dims[l];
    sh->out[l] = geo->out_dims[l];
    if (l > 0) {
      const int want = (l == sh->skip) ? sh->out[l - 1] + sh->pe_dim : sh->out[l - 1];
      PSN_REQUIRE(sh->in[l] == want, PSN_ERR_SHAPE, "stage-1 train: geo layer %d input width %d, expected %d", l, sh->in[l], want);
    }
    if (l < sh->nh) sh->h_total += sh->out[l];
  }
  PSN_REQUIRE(sh->skip != 0 && sh->skip < sh->nh + 1, PSN_ERR_SHAPE, "stage-1 train: skip layer %d", sh->skip);
  sh->feat = sh->out[sh->nh] - 1;
  sh->has_app = app ? 1 : 0;
  if (app) {
    sh->nla = app->n_layers;
    sh->app_in = 3 + sh->pv_dim + 3 + sh->feat;
    PSN_REQUIRE(app->n_layers >= 2 && app->n_layers <= kMaxLayers && app->in_dims[0] == sh->app_in && app->out_dims[app->n_layers - 1] == 3,
                PSN_ERR_SHAPE, "stage-1 train: app net layout (input %d, expected %d)", app->in_dims[0], sh->app_in);
  }
  return PSN_OK;
}

static size_t s1_tape_floats(const S1Shape& sh, long long M) {
  size_t f = (size_t)M * sh.pe_dim + 64;
  f += 3 * ((size_t)M * sh.h_total + 64 * sh.nh);
  if (sh.skip > 0) f += (size_t)M * sh.in[sh.skip] + 64;
  if (sh.has_app) {
    f += (size_t)M * sh.app_in + 64 + (size_t)M * 3 + 64;
    // hidden activations of the appearance MLP; widths come from the net at carve time (<= 256 assumed for sizing below)
  }
  return f;
}

static int s1_carve(const S1Shape& sh, const psn_train_net* app, long long M, void* tape, long long tape_bytes, S1Tape* t) {
  Workspace w(tape, tape_bytes);
  memset(t, 0, sizeof(*t));
  t->pe = w.take<float>((size_t)M * sh.pe_dim);
  for (int l = 0; l < sh.nh; ++l) {
    t->h[l] = w.take<float>((size_t)M * sh.out[l]);
    t->s[l] = w.take<float>((size_t)M * sh.out[l]);
    t->ap[l] = w.take<float>((size_t)M * sh.out[l]);
  }
  if (sh.skip > 0) t->x_skip = w.take<float>((size_t)M * sh.in[sh.skip]);
  if (sh.has_app) {
    t->u = w.take<float>((size_t)M * sh.app_in);
    t->t3 = w.take<float>((size_t)M * 3);
    for (int l = 0; l + 1 < app->n_layers; ++l) t->y[l] = w.take<float>((size_t)M * app->out_dims[l]);
  }
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "stage-1 train: tape too small (need %zu bytes, have %lld)", w.used, tape_bytes);
  return PSN_OK;
}

}  // namespace psn

using namespace psn;

extern "C" int64_t psn_s1_train_tape_bytes(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, int64_t M) {
  S1Shape sh;
  if (s1_shape(geo, app, octaves, octaves_view, &sh)) return -1;
  size_t f = s1_tape_floats(sh, M);
  if (app)
    for (int l = 0; l + 1 < app->n_layers; ++l) f += (size_t)M * app->out_dims[l] + 64;
  return (int64_t)(f * sizeof(float) + 256 * (8 + 3 * sh.nh + (app ? app->n_layers : 0)));
}

// scratch of the forward / backward calls: a handful of [M, wmax] buffers
extern "C" int64_t psn_s1_train_ws_bytes(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, int64_t M) {
  S1Shape sh;
  if (s1_shape(geo, app, octaves, octaves_view, &sh)) return -1;
  int wmax = sh.out[sh.nh];
  for (int l = 0; l < sh.nh; ++l) wmax = sh.in[l] > wmax ? sh.in[l] : wmax;
  if (app) {
    wmax = sh.app_in > wmax ? sh.app_in : wmax;
    for (int l = 0; l < app->n_layers; ++l) wmax = app->out_dims[l] > wmax ? app->out_dims[l] : wmax;
  }
  return (int64_t)((size_t)M * (6 * (size_t)wmax + 2 * sh.pe_dim + 16) * sizeof(float) + 16 * 256);
}

extern "C" int psn_s1_train_forward(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, float rescale,
                                    const float* pts, const float* views, int64_t M, float* rgb, float* logit, float* grad, void* tape,
                                    int64_t tape_bytes, void* ws, int64_t ws_bytes, void* stream) {
  S1Shape sh;
  int rc = s1_shape(geo, app, octaves, octaves_view, &sh);
  if (rc) return rc;
  PSN_REQUIRE(pts && grad && (!app || (views && rgb)), PSN_ERR_ARG, "psn_s1_train_forward: null argument");
  if (M == 0) return PSN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  S1Tape t;
  if ((rc = s1_carve(sh, app, M, tape, tape_bytes, &t))) return rc;
  Workspace w(ws, ws_bytes);
  int wmax = sh.out[sh.nh];
  for (int l = 0; l < sh.nh; ++l) wmax = sh.in[l] > wmax ? sh.in[l] : wmax;
  float* T1 = w.take<float>((size_t)M * wmax);
  float* T2 = w.take<float>((size_t)M * wmax);
  float* T3 = w.take<float>((size_t)M * wmax);
  float* gpe = w.take<float>((size_t)M * sh.pe_dim);
  float* pre = w.take<float>((size_t)M * 4);
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_s1_train_forward: workspace too small (need %zu bytes, have %lld)", w.used, (long long)ws_bytes);
  const int nh = sh.nh;
  const bool fused = train_gemm_fused();  // element-wise passes inside the GEMM epilogues (tc_gemm.cu, EPI 4-7)
  // ---- forward stack ----------------------------------------------------------------------------------------------------
  count_launch();
  k_s1_pe<<<nblk(M), 256, 0, st>>>(pts, M, octaves, rescale, t.pe, sh.pe_dim);
  for (int l = 0; l < nh; ++l) {
    const float* x = (l == 0) ? t.pe : (l == sh.skip ? t.x_skip : t.h[l - 1]);
    if (fused) {  // h_l = softplus(z_l) and s_l = sigmoid(100 z_l) written by the GEMM epilogue
      const GemmFuse fz = {t.s[l], nullptr, nullptr, sh.out[l], 0.f};
      if ((rc = gemm(0, x, sh.in[l], geo->W[l], sh.in[l], t.h[l], sh.out[l], geo->b[l], M, sh.out[l], sh.in[l], 4, st, &fz))) return rc;
    } else {
      if ((rc = gemm(0, x, sh.in[l], geo->W[l], sh.in[l], t.h[l], sh.out[l], geo->b[l], M, sh.out[l], sh.in[l], 1, st))) return rc;
      count_launch();
      k_s1_softplus<<<nblk(M * sh.out[l]), 256, 0, st>>>(t.h[l], t.s[l], M * sh.out[l]);
    }
    if (l + 1 == sh.skip) {
      count_launch();
      k_s1_skip_cat<<<nblk(M * sh.in[sh.skip]), 256, 0, st>>>(t.h[l], sh.out[l], t.pe, sh.pe_dim, M, t.x_skip);
    }
  }
  // last layer: logit + feature (written straight behind the gradient slot of the appearance input when there is one)
  if (app) {
    float* dst = t.u + 3 + sh.pv_dim + 2;  // column of the logit = last gradient slot; the feature starts one further
    if ((rc = gemm(0, t.h[nh - 1], sh.out[nh - 1], geo->W[nh], sh.in[nh], dst, sh.app_in, geo->b[nh], M, sh.out[nh], sh.in[nh], 1, st)))
      return rc;
    if (logit) {
      count_launch();
      k_s1_copy_col<<<nblk(M), 256, 0, st>>>(t.u, sh.app_in, 3 + sh.pv_dim + 2, M, logit);
    }
  } else if (logit) {
    if ((rc = gemm(0, t.h[nh - 1], sh.out[nh - 1], geo->W[nh], sh.in[nh], T1, sh.out[nh], geo->b[nh], M, sh.out[nh], sh.in[nh], 1, st)))
      return rc;
    count_launch();
    k_s1_copy_col<<<nblk(M), 256, 0, st>>>(T1, sh.out[nh], 0, M, logit);
  }
  // ---- analytic normal: reverse sweep ------------------------------------------------------------------------------------------
  count_launch();
  k_s1_bcast_row<<<nblk(M * sh.out[nh - 1]), 256, 0, st>>>(geo->W[nh], sh.out[nh - 1], M, t.ap[nh - 1]);  // row 0 of the last layer
  count_launch();
  k_s1_mul<<<nblk(M * sh.out[nh - 1]), 256, 0, st>>>(t.ap[nh - 1], t.s[nh - 1], M * sh.out[nh - 1], T1);  // dz of the last hidden layer
  float *dz = T1, *dz_next = T3;
  for (int l = nh - 1; l >= 0; --l) {
    float* dst = (l == 0 || l == sh.skip) ? T2 : t.ap[l - 1];
    if (fused && l > 0 && l != sh.skip) {
      // a'_{l-1} = dz_l W_l and, in the same epilogue, dz_{l-1} = a'_{l-1} * s_{l-1} (the next iteration's operand)
      const GemmFuse fz = {dz_next, t.s[l - 1], nullptr, sh.in[l], 0.f};
      if ((rc = gemm(1, dz, sh.out[l], geo->W[l], sh.in[l], dst, sh.in[l], nullptr, M, sh.in[l], sh.out[l], 7, st, &fz))) return rc;
      float* sw = dz; dz = dz_next; dz_next = sw;
      continue;
    }
    if ((rc = gemm(1, dz, sh.out[l], geo->W[l], sh.in[l], dst, sh.in[l], nullptr, M, sh.in[l], sh.out[l], 0, st))) return rc;
    if (l == sh.skip) {
      count_launch();
      k_s1_skip_split<<<nblk(M * sh.in[l]), 256, 0, st>>>(T2, sh.out[l - 1], sh.pe_dim, M, t.ap[l - 1], gpe);
    } else if (l == 0) {
      count_launch();
      k_s1_axpy<<<nblk(M * sh.pe_dim), 256, 0, st>>>(T2, M * sh.pe_dim, sh.skip > 0 ? 1 : 0, gpe);
    }
    if (l > 0) {
      count_launch();
      k_s1_mul<<<nblk(M * sh.out[l - 1]), 256, 0, st>>>(t.ap[l - 1], t.s[l - 1], M * sh.out[l - 1], dz);  // dz_{l-1} (its reader has run)
    }
  }
  count_launch();
  k_s1_pe_jt<<<nblk(M), 256, 0, st>>>(gpe, t.pe, sh.pe_dim, M, octaves, rescale, grad, app ? t.u + 3 + sh.pv_dim : nullptr, sh.app_in);
  // ---- appearance MLP ---------------------------------------------------------------------------------------------------------
  if (app) {
    count_launch();
    k_s1_app_input<<<nblk(M), 256, 0, st>>>(pts, views, M, octaves_view, t.u, sh.app_in);
    const int na = app->n_layers;
    for (int l = 0; l + 1 < na; ++l) {
      const float* x = l == 0 ? t.u : t.y[l - 1];
      if ((rc = gemm(0, x, app->in_dims[l], app->W[l], app->in_dims[l], t.y[l], app->out_dims[l], app->b[l], M, app->out_dims[l],
                     app->in_dims[l], 2, st)))
        return rc;
    }
    if ((rc = gemm(0, t.y[na - 2], app->in_dims[na - 1], app->W[na - 1], app->in_dims[na - 1], pre, 3, app->b[na - 1], M, 3, app->in_dims[na - 1],
                   1, st)))
      return rc;
    count_launch();
    k_s1_rgb<<<nblk(M * 3), 256, 0, st>>>(pre, M * 3, t.t3, rgb);
  }
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

// g_rgb [M,3] (app only), g_logit [M], g_grad [M,3]: any may be NULL (= zero).  Accumulates into geo->dW/db (and app->dW/db).
extern "C" int psn_s1_train_backward(const psn_train_net* geo, const psn_train_net* app, int octaves, int octaves_view, float rescale,
                                     int64_t M, const float* g_rgb, const float* g_logit, const float* g_grad, void* tape, int64_t tape_bytes,
                                     void* ws, int64_t ws_bytes, void* stream) {
  S1Shape sh;
  int rc = s1_shape(geo, app, octaves, octaves_view, &sh);
  if (rc) return rc;
  if (M == 0) return PSN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  S1Tape t;
  if ((rc = s1_carve(sh, app, M, tape, tape_bytes, &t))) return rc;
  Workspace w(ws, ws_bytes);
  int wmax = sh.out[sh.nh];
  for (int l = 0; l < sh.nh; ++l) wmax = sh.in[l] > wmax ? sh.in[l] : wmax;
  if (app) {
    wmax = sh.app_in > wmax ? sh.app_in : wmax;
    for (int l = 0; l < app->n_layers; ++l) wmax = app->out_dims[l] > wmax ? app->out_dims[l] : wmax;
  }
  float* T1 = w.take<float>((size_t)M * wmax);
  float* T2 = w.take<float>((size_t)M * wmax);
  float* A0 = w.take<float>((size_t)M * wmax);
  float* A1 = w.take<float>((size_t)M * wmax);
  float* UB = w.take<float>((size_t)M * wmax);  // ubar of the appearance input
  float* OB = w.take<float>((size_t)M * wmax);  // outbar of the last geo layer
  float* gpe_bar = w.take<float>((size_t)M * sh.pe_dim);
  float* gbar = w.take<float>((size_t)M * 4);
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_s1_train_backward: workspace too small (need %zu bytes, have %lld)", w.used, (long long)ws_bytes);
  const int nh = sh.nh;
  const bool have_u = app && g_rgb;
  const bool fused = train_gemm_fused();  // element-wise passes inside the GEMM epilogues (tc_gemm.cu, EPI 4-7)
  // ---- A. appearance MLP backward -----------------------------------------------------------------------------------------------
  if (have_u) {
    const int na = app->n_layers;
    count_launch();
    k_s1_rgb_bwd<<<nblk(M * 3), 256, 0, st>>>(g_rgb, t.t3, M * 3, T1);  // d pre_last [M,3]
    const float* dz = T1;
    int lddz = 3;
    for (int l = na - 1; l >= 0; --l) {
      const float* x = l == 0 ? t.u : t.y[l - 1];
      const int K = app->in_dims[l], N = app->out_dims[l];
      if ((rc = gemm(2, dz, lddz, x, K, app->dW[l], K, nullptr, N, K, M, 0, st))) return rc;
      count_launch();
      colsum(dz, lddz, M, N, app->db[l], st);
      float* dx = l == 0 ? UB : T2;
      if ((rc = gemm(1, dz, lddz, app->W[l], K, dx, K, nullptr, M, K, N, 0, st))) return rc;
      if (l > 0) {  // relu'
        float* nz = (dz == T1) ? A0 : T1;
        count_launch();
        k_act_bwd<<<nblk(M * K), 256, 0, st>>>(T2, K, t.y[l - 1], K, nz, K, M, K, 2);
        dz = nz;
        lddz = K;
      }
    }
  }
  // ---- B. cotangent of the analytic normal ----------------------------------------------------------------------------------------
  const bool have_g = g_grad || have_u;
  const int gcol = 3 + sh.pv_dim;
  if (have_g) {
    count_launch();
    k_s1_gbar<<<nblk(M * 3), 256, 0, st>>>(g_grad, have_u ? UB : nullptr, sh.app_in, gcol, M, gbar);
    // ---- C. second-order pass: reverse-mode sweep over the a-chain, l ascending --------------------------------------------------
    count_launch();
    k_s1_pe_j<<<nblk(M), 256, 0, st>>>(gbar, t.pe, sh.pe_dim, M, octaves, rescale, gpe_bar);
    const float* abar = gpe_bar;  // cotangent of a_0 [M, pe_dim]
    for (int l = 0; l < nh; ++l) {
      const int K = sh.in[l], N = sh.out[l];
      count_launch();
      k_s1_mul<<<nblk(M * N), 256, 0, st>>>(t.ap[l], t.s[l], M * N, T1);                        // dz_l
      if ((rc = gemm(2, T1, N, abar, K, geo->dW[l], K, nullptr, N, K, M, 0, st))) return rc;    // Wbar_l += dz_l^T abar_l
      float* an = (abar == A0) ? A1 : A0;
      if (fused) {  // dzbar = abar_l W_l^T with k_s1_second as its epilogue: abar' = dzbar s_l, ap_l := zbar2_l
        const GemmFuse fz = {nullptr, t.s[l], t.ap[l], N, 0.f};
        if ((rc = gemm(0, abar, K, geo->W[l], K, l + 1 == sh.skip ? T1 : an, N, nullptr, M, N, K, 5, st, &fz))) return rc;
      } else {
        if ((rc = gemm(0, abar, K, geo->W[l], K, T2, N, nullptr, M, N, K, 0, st))) return rc;     // dzbar = abar_l W_l^T
        count_launch();
        k_s1_second<<<nblk(M * N), 256, 0, st>>>(T2, t.ap[l], t.s[l], M * N, l + 1 == sh.skip ? T1 : an);
      }
      if (l + 1 == sh.skip) {
        count_launch();
        k_s1_skip_merge<<<nblk(M * sh.in[sh.skip]), 256, 0, st>>>(T1, N, gpe_bar, sh.pe_dim, M, an);
      }
      abar = an;
    }
    count_launch();  // a'_8 = row 0 of the last layer: its gradient is the column sum of abar'
    k_s1_colsum_row0<<<dim3((sh.out[nh - 1] + 255) / 256, (unsigned)((M + 1023) / 1024)), 256, 0, st>>>(abar, sh.out[nh - 1], M, geo->dW[nh]);
  } else {
    for (int l = 0; l < nh; ++l) PSN_CUDA_CHECK(cudaMemsetAsync(t.ap[l], 0, (size_t)M * sh.out[l] * sizeof(float), st));  // zbar2 = 0
  }
  // ---- D. ordinary backward of the forward stack -------------------------------------------------------------------------------------
  const bool have_out = g_logit || have_u;
  const float* hbar = nullptr;
  int ldh = 0;
  float hscale = 1.f;
  const float* zfused = nullptr;  // fused path: zbar of the layer about to be visited (produced by the previous GEMM's epilogue)
  if (have_out) {
    const int No = sh.out[nh], Ko = sh.in[nh];
    count_launch();
    k_s1_outbar<<<nblk(M * No), 256, 0, st>>>(g_logit, have_u ? UB : nullptr, sh.app_in, gcol + 3, sh.feat, M, OB);
    if ((rc = gemm(2, OB, No, t.h[nh - 1], Ko, geo->dW[nh], Ko, nullptr, No, Ko, M, 0, st))) return rc;
    count_launch();
    colsum(OB, No, M, No, geo->db[nh], st);
    if (fused) {  // hbar = OB W_last and zbar_{nh-1} = hbar s + zbar2 in one launch
      const GemmFuse fz = {nullptr, t.s[nh - 1], t.ap[nh - 1], Ko, 1.f};
      if ((rc = gemm(1, OB, No, geo->W[nh], Ko, A0, Ko, nullptr, M, Ko, No, 6, st, &fz))) return rc;
      zfused = A0;
    } else {
      if ((rc = gemm(1, OB, No, geo->W[nh], Ko, A0, Ko, nullptr, M, Ko, No, 0, st))) return rc;
      hbar = A0;
      ldh = Ko;
    }
  } else if (fused) {
    zfused = t.ap[nh - 1];  // no cotangent on the outputs: zbar = zbar2
  }
  if (!have_out && !have_g) return PSN_OK;
  if (fused) {
    for (int l = nh - 1; l >= 0; --l) {
      const int K = sh.in[l], N = sh.out[l];
      const float* x = (l == 0) ? t.pe : (l == sh.skip ? t.x_skip : t.h[l - 1]);
      if ((rc = gemm(2, zfused, N, x, K, geo->dW[l], K, nullptr, N, K, M, 0, st))) return rc;
      count_launch();
      colsum(zfused, N, M, N, geo->db[l], st);
      if (l > 0) {
        // zbar_{l-1} = (zbar_l W_l)[:, :out_{l-1}] * hscale * s_{l-1} + zbar2_{l-1}   (x_skip = cat[h, pe] / sqrt2: only h's share)
        const int Np = sh.out[l - 1];
        float* zn = (zfused == A0) ? A1 : A0;
        const GemmFuse fz = {nullptr, t.s[l - 1], t.ap[l - 1], Np, (l == sh.skip) ? PSN_S1_INV_SQRT2 : 1.f};
        if ((rc = gemm(1, zfused, N, geo->W[l], K, zn, Np, nullptr, M, Np, N, 6, st, &fz))) return rc;
        zfused = zn;
      }
    }
    PSN_CUDA_CHECK(cudaGetLastError());
    return PSN_OK;
  }
  for (int l = nh - 1; l >= 0; --l) {
    const int K = sh.in[l], N = sh.out[l];
    count_launch();
    k_s1_zbar<<<nblk(M * N), 256, 0, st>>>(hbar, ldh, hscale, t.s[l], t.ap[l], N, M, T1);  // zbar_l
    const float* x = (l == 0) ? t.pe : (l == sh.skip ? t.x_skip : t.h[l - 1]);
    if ((rc = gemm(2, T1, N, x, K, geo->dW[l], K, nullptr, N, K, M, 0, st))) return rc;
    count_launch();
    colsum(T1, N, M, N, geo->db[l], st);
    if (l > 0) {
      float* xb = (hbar == A0) ? A1 : A0;
      if ((rc = gemm(1, T1, N, geo->W[l], K, xb, K, nullptr, M, K, N, 0, st))) return rc;
      hbar = xb;
      ldh = K;                                          // the first out[l-1] columns are h_{l-1}'s share
      hscale = (l == sh.skip) ? PSN_S1_INV_SQRT2 : 1.f;  // x_skip = cat[h, pe] / sqrt2
    }
  }
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

extern "C" int psn_composite_bwd(const float* rgb_s, const float* alpha, int64_t N, int S, int white_background, const float* g_rgb,
                                 const float* g_acc, float* d_rgb_s, float* d_alpha, void* stream) {
  PSN_REQUIRE(rgb_s && alpha && d_rgb_s && d_alpha && S >= 1, PSN_ERR_ARG, "psn_composite_bwd: bad argument");
  PSN_REQUIRE(S <= kCompositeMaxS, PSN_ERR_SHAPE, "psn_composite_bwd: %d samples per ray > %d", S, kCompositeMaxS);
  if (N == 0) return PSN_OK;
  count_launch();
  k_composite_bwd<<<nblk(N, 128), 128, 0, (cudaStream_t)stream>>>(rgb_s, alpha, N, S, white_background, g_rgb, g_acc, d_rgb_s, d_alpha);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
