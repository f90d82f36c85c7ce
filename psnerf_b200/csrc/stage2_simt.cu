// Stage-2 photometric-stereo shading on fp32 FFMA (stage2/model/renderer.py:110-266, sgbasis.py:16-32,
// embedder.py:6-54): per-point normal/albedo/SG-weight MLPs, per-(light,point) visibility MLP, SG shading
// epilogue writing the image-shaped outputs directly (no masked-scatter temporaries).
#include "simt_mlp.cuh"
#include "launch.cuh"
#include "internal.cuh"
#include "prof.cuh"

namespace psn {

struct S2Dev {
  SimtLayer fwd[kMaxLayers];
  int n_layers, skip, final_act, in0;
};

constexpr int S2_OUT_ROWS = 32;

template <int NC>
__device__ __forceinline__ void s2_layer(const S2Dev& n, int l, const float* E, float* X, float* WS, float* OUT) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[8][NC];
  dense<NC>(n.fwd[l], l == 0 ? E : X, WS, acc);
  const bool last = (l == n.n_layers - 1);
  const int N = n.fwd[l].N;
  float* dst = last ? OUT : X;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int col = simt_col<NC>(tx, j);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = acc[i][j];
      v[i] = last ? (n.final_act == 1 ? sigmoidf_(a) : a) : fmaxf(a, 0.f);
    }
    if (col < N) store_col8(dst, col, ty, v);
  }
  if (!last) {  // rows N.. of the next input: cat[y, x0] after the skip layer (renderer.py:30-31), then zero padding
    const int kp = n.fwd[l + 1].K_pad;
    const bool cat = (l == n.skip);
    for (int idx = threadIdx.x; idx < (kp - N) * TM; idx += NT) {
      const int k = idx / TM, r = idx - k * TM;
      X[(size_t)(N + k) * LDX + r] = (cat && k < n.in0) ? E[k * LDX + r] : 0.f;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void s2_run(const S2Dev& n, const float* E, float* X, float* WS, float* OUT) {
  for (int l = 0; l < n.n_layers; ++l) {
    switch (n.fwd[l].N_pad) {
      case 256: s2_layer<8>(n, l, E, X, WS, OUT); break;
      case 128: s2_layer<4>(n, l, E, X, WS, OUT); break;
      case 64: s2_layer<2>(n, l, E, X, WS, OUT); break;
      default: s2_layer<1>(n, l, E, X, WS, OUT); break;
    }
  }
}

// embed rows [row0, row0 + 3 + 6*nf) of E for the tile: [x, sin(2^0 x), cos(2^0 x), ...] (embedder.py:27-36)
__device__ __forceinline__ void embed_rows(float* E, int row0, const float* P /*[3][TM]*/, int nf) {
  const int r = threadIdx.x & (TM - 1), q = threadIdx.x / TM;
  const float x[3] = {P[r], P[TM + r], P[2 * TM + r]};
  if (q == 0) {
    E[(row0 + 0) * LDX + r] = x[0]; E[(row0 + 1) * LDX + r] = x[1]; E[(row0 + 2) * LDX + r] = x[2];
  }
  for (int idx = q; idx < nf * 3; idx += NT / TM) {
    const int i = idx / 3, c = idx - 3 * i;
    float s, co;
    sincosf(x[c] * (float)(1 << i), &s, &co);
    E[(row0 + 3 + 6 * i + c) * LDX + r] = s;
    E[(row0 + 6 + 6 * i + c) * LDX + r] = co;
  }
}

// ---- per-point nets --------------------------------------------------------------------------------------------
// nets[0] = normal_net (optional), nets[1] = albedo_net, nets[2] = rough_net (either may be absent: n_layers == 0).
struct PointNets { S2Dev net[3]; int nf[3]; };

__global__ void __launch_bounds__(NT, 1)
k_s2_point(PointNets pn, const float* __restrict__ pts, long long Ns, float* __restrict__ normal, float* __restrict__ albedo,
           float* __restrict__ weights, int nbt, int k_rows) {
  extern __shared__ __align__(16) float smem[];
  float* E = smem;                      // [64][LDX]
  float* X = E + 64 * LDX;              // [k_rows][LDX]
  float* WS = X + (size_t)k_rows * LDX;
  float* OUT = WS + WRING_FLOATS;       // [32][LDX]
  float* P = OUT + S2_OUT_ROWS * LDX;   // [3][TM]
  const long long n_tiles = (Ns + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * TM;
    if (threadIdx.x < TM) {
      const long long i = base + threadIdx.x;
#pragma unroll
      for (int c = 0; c < 3; ++c) P[c * TM + threadIdx.x] = (i < Ns) ? pts[i * 3 + c] : 0.f;
    }
    __syncthreads();
    int nf_cur = -1;
    for (int which = 0; which < 3; ++which) {
      const S2Dev& n = pn.net[which];
      if (n.n_layers == 0) continue;
      if (pn.nf[which] != nf_cur) {
        nf_cur = pn.nf[which];
        embed_rows(E, 0, P, nf_cur);
        const int used = 3 + 6 * nf_cur;
        for (int idx = threadIdx.x; idx < (64 - used) * TM; idx += NT) {
          const int k = idx / TM, r = idx - k * TM;
          E[(used + k) * LDX + r] = 0.f;
        }
        __syncthreads();
      }
      s2_run(n, E, X, WS, OUT);
      if (threadIdx.x < TM && base + threadIdx.x < Ns) {
        const int r = threadIdx.x;
        const long long i = base + r;
        if (which == 0) {  // F.normalize(normal_net(...)) (renderer.py:130-131)
          const float a = OUT[0 * LDX + r], b = OUT[1 * LDX + r], c = OUT[2 * LDX + r];
          const float nn = fmaxf(sqrtf(a * a + b * b + c * c), 1e-12f);
          normal[i * 3] = a / nn; normal[i * 3 + 1] = b / nn; normal[i * 3 + 2] = c / nn;
        } else if (which == 1) {
          albedo[i * 3] = OUT[r]; albedo[i * 3 + 1] = OUT[LDX + r]; albedo[i * 3 + 2] = OUT[2 * LDX + r];
        } else {
          for (int k = 0; k < nbt; ++k) weights[i * nbt + k] = fmaxf(OUT[k * LDX + r], 0.f);  // F.relu (renderer.py:174)
        }
      }
      __syncthreads();
    }
  }
}

// ---- visibility MLP over (light, point) pairs, light-major ----------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_s2_vis(S2Dev n, int nf, const float* __restrict__ pts, long long Ns, const float* __restrict__ lights, long long pairs,
         float* __restrict__ vis) {
  extern __shared__ __align__(16) float smem[];
  float* E = smem;                     // [128][LDX]
  float* X = E + 128 * LDX;            // [384][LDX]
  float* WS = X + 384 * LDX;
  float* OUT = WS + WRING_FLOATS;
  float* P = OUT + S2_OUT_ROWS * LDX;  // [3][TM] points
  float* Lg = P + 3 * TM;              // [3][TM] light dirs
  const int ed = 3 + 6 * nf;
  const long long n_tiles = (pairs + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * TM;
    if (threadIdx.x < TM) {
      const long long i = base + threadIdx.x;
      const long long l = (i < pairs) ? i / Ns : 0, p = (i < pairs) ? i - l * Ns : 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        P[c * TM + threadIdx.x] = (i < pairs) ? pts[p * 3 + c] : 0.f;
        Lg[c * TM + threadIdx.x] = (i < pairs) ? lights[l * 3 + c] : 0.f;
      }
    }
    __syncthreads();
    embed_rows(E, 0, P, nf);
    embed_rows(E, ed, Lg, nf);
    for (int idx = threadIdx.x; idx < (128 - 2 * ed) * TM; idx += NT) {
      const int k = idx / TM, r = idx - k * TM;
      E[(2 * ed + k) * LDX + r] = 0.f;
    }
    __syncthreads();
    s2_run(n, E, X, WS, OUT);
    if (threadIdx.x < TM && base + threadIdx.x < pairs) vis[base + threadIdx.x] = OUT[threadIdx.x];
    __syncthreads();
  }
}

// ---- SG shading + image-shaped writes --------------------------------------------------------------------------
struct ShadeArgs {
  const float *normal, *albedo, *weights, *view, *vis, *lights, *lobe, *intensity;
  const int* slot_of_pixel;
  float *rgb, *spec, *vis_out, *normal_out, *albedo_out, *sgw_out;
  long long N, Ns;
  int L, nbasis, specular_rgb, nbt, intensity_kind, write_normal, microfacet;
  float intensity_scalar, f0;
};

// x / (y + 1e-6) with inf / nan -> 0 (microfacet.py:20-24)
__device__ __forceinline__ float div_no_nan(float x, float y) {
  const float a = __fdiv_rn(x, __fadd_rn(y, 1e-6f));
  return (isinf(a) || isnan(a)) ? 0.f : a;
}
__device__ __forceinline__ void normalize_eps(const float* v, float eps, float* o) {  // F.normalize(v, eps=eps)
  const float n = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), eps);
  o[0] = v[0] / n; o[1] = v[1] / n; o[2] = v[2] / n;
}
// GGX microfacet BRDF without the Lambert term (stage2/model/microfacet.py:35-114): Schlick Fresnel, GGX distribution and
// Smith-GGX geometry term with the reference's clamps / divide_no_nan conventions; alpha = rough^2.
__device__ __forceinline__ float microfacet_glossy(const float* l_in, const float* v_in, const float* n_in, float rough, float f0) {
  const float kPi = 3.14159265358979323846f;
  float l[3], v[3], n[3], h[3];
  normalize_eps(l_in, 1e-6f, l); normalize_eps(v_in, 1e-6f, v); normalize_eps(n_in, 1e-6f, n);
  const float hs[3] = {l[0] + v[0], l[1] + v[1], l[2] + v[2]};
  normalize_eps(hs, 1e-6f, h);
  const float ldh = l[0] * h[0] + l[1] * h[1] + l[2] * h[2];
  const float om = 1.f - ldh;
  const float f = f0 + (1.f - f0) * (om * om * om * om * om);
  const float alpha = rough * rough, a2 = alpha * alpha;
  // D
  const float cm = h[0] * n[0] + h[1] * n[1] + h[2] * n[2];
  const float cm2 = cm * cm;
  const float tm2 = div_no_nan(1.f - cm2, cm2);
  const float dd = kPi * (cm2 * cm2) * ((a2 + tm2) * (a2 + tm2));
  const float d = div_no_nan(a2 * (cm > 0.f ? 1.f : 0.f), dd);
  // G
  const float cv = n[0] * v[0] + n[1] * v[1] + n[2] * v[2];
  const float ct = h[0] * v[0] + h[1] * v[1] + h[2] * v[2];
  const float chi = div_no_nan(ct, cv) > 0.f ? 1.f : 0.f;
  const float cv2 = fminf(fmaxf(cv * cv, 0.f), 1.f);
  const float tv2 = fmaxf(div_no_nan(1.f - cv2, cv2), 0.f);
  const float g = div_no_nan(chi * 2.f, 1.f + sqrtf(1.f + a2 * tv2));
  const float ldn = l[0] * n[0] + l[1] * n[1] + l[2] * n[2];
  return div_no_nan(f * g * d, 4.f * fabsf(ldn) * fabsf(cv));
}

__global__ void k_fill_int(int* p, long long n, int v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_slot_of_pixel(const int* __restrict__ pix, long long Ns, int* __restrict__ slot_of_pixel) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Ns) slot_of_pixel[pix[i]] = (int)i;
}

// grid.y = light (plus one extra row, l == L, that writes the per-pixel outputs), threads over pixels.
__global__ void k_s2_shade(ShadeArgs a) {
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= a.N) return;
  const int l = blockIdx.y;
  const int slot = a.slot_of_pixel[px];
  if (l == a.L) {  // per-pixel outputs, pre-filled like renderer.py:133,146-152
    float n[3] = {1.f, 1.f, 1.f}, al[3] = {1.f, 1.f, 1.f};
    if (slot >= 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { n[c] = a.normal[(long long)slot * 3 + c]; al[c] = a.albedo[(long long)slot * 3 + c]; }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (a.write_normal) a.normal_out[px * 3 + c] = n[c];
      a.albedo_out[px * 3 + c] = al[c];
    }
    if (a.microfacet) {  // roughness image, pre-filled with 1 (renderer.py:140-141,204-207)
      const float r = (slot >= 0) ? a.weights[slot] : 1.f;
      a.spec[px * 3] = r; a.spec[px * 3 + 1] = r; a.spec[px * 3 + 2] = r;
    } else {
      for (int k = 0; k < a.nbt; ++k) a.sgw_out[px * a.nbt + k] = (slot >= 0) ? a.weights[(long long)slot * a.nbt + k] : 0.f;
    }
    return;
  }
  const long long o = ((long long)l * a.N + px) * 3;
  if (slot < 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a.rgb[o + c] = 1.f;
      if (!a.microfacet) a.spec[o + c] = 1.f;
      if (a.vis_out) a.vis_out[o + c] = 1.f;
    }
    return;
  }
  const float* nn = a.normal + (long long)slot * 3;
  const float* vv = a.view + (long long)slot * 3;
  const float* ll = a.lights + (long long)l * 3;
  if (a.microfacet) {  // brdf = glossy + albedo / pi (microfacet.py:62-72), then the common shading line renderer.py:187-199
    const float gl = microfacet_glossy(ll, vv, nn, a.weights[slot], a.f0);
    const float cosv = ll[0] * nn[0] + ll[1] * nn[1] + ll[2] * nn[2];
    float visr = 1.f, visc = 1.f;
    if (a.vis) {
      visr = a.vis[(long long)l * a.Ns + slot];
      visc = fminf(fmaxf(visr, 0.f), 1.f);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float brdf = gl + a.albedo[(long long)slot * 3 + c] / 3.14159265358979323846f;
      float I = a.intensity_scalar;
      if (a.intensity_kind == 1) I = a.intensity[l];
      else if (a.intensity_kind == 2) I = a.intensity[l * 3 + c];
      float r = brdf * I * cosv;
      if (a.vis) r = r * visc;
      a.rgb[o + c] = fminf(fmaxf(r, 0.f), 1.f);
      if (a.vis_out) a.vis_out[o + c] = visr;
    }
    return;
  }
  // h = F.normalize(l + v); (h*n).sum(-1) - 1 with torch's separate roundings (sgbasis.py:24-25): the lobe sharpness
  // lambda <= e^10 amplifies every ulp of this dot product 2e4 times, so no FMA contraction here.
  const float hx = __fadd_rn(ll[0], vv[0]), hy = __fadd_rn(ll[1], vv[1]), hz = __fadd_rn(ll[2], vv[2]);
  const float hn = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(hx, hx), __fmul_rn(hy, hy)), __fmul_rn(hz, hz))), 1e-12f);
  const float hdn = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fdiv_rn(hx, hn), nn[0]), __fmul_rn(__fdiv_rn(hy, hn), nn[1])),
                                        __fmul_rn(__fdiv_rn(hz, hn), nn[2])), 1.f);
  float spec[3] = {0.f, 0.f, 0.f};
  const float* w = a.weights + (long long)slot * a.nbt;
  for (int k = 0; k < a.nbasis; ++k) {
    const float D = expf(fmaxf(a.lobe[k], 0.f) * hdn);
    if (a.specular_rgb) {
      spec[0] += w[k] * D; spec[1] += w[a.nbasis + k] * D; spec[2] += w[2 * a.nbasis + k] * D;
    } else {
      spec[0] += w[k] * D;
    }
  }
  if (!a.specular_rgb) spec[1] = spec[2] = spec[0];
  const float cosv = ll[0] * nn[0] + ll[1] * nn[1] + ll[2] * nn[2];  // NOT clamped (renderer.py:187)
  float visr = 1.f, visc = 1.f;
  if (a.vis) {
    visr = a.vis[(long long)l * a.Ns + slot];
    visc = fminf(fmaxf(visr, 0.f), 1.f);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float sp = fmaxf(spec[c], 0.f);
    const float brdf = a.albedo[(long long)slot * 3 + c] + sp;
    float I = a.intensity_scalar;
    if (a.intensity_kind == 1) I = a.intensity[l];
    else if (a.intensity_kind == 2) I = a.intensity[l * 3 + c];
    float r = brdf * I * cosv;
    if (a.vis) r = r * visc;
    a.rgb[o + c] = fminf(fmaxf(r, 0.f), 1.f);
    a.spec[o + c] = sp;
    if (a.vis_out) a.vis_out[o + c] = visr;
  }
}

// ---- host -------------------------------------------------------------------------------------------------------
static int make_s2_dev(const psn_mlp* net, int expect_in, S2Dev* d, const char* what) {
  memset(d, 0, sizeof(*d));
  if (!net) return PSN_OK;
  PSN_REQUIRE(net->kind == PSN_NET_S2, PSN_ERR_ARG, "%s: expected a PSN_NET_S2 handle", what);
  d->n_layers = net->n_layers;
  d->skip = net->desc.skip;
  d->final_act = net->desc.final_act;
  d->in0 = net->in_dims[0];
  PSN_REQUIRE(d->in0 == expect_in, PSN_ERR_SHAPE, "%s: input width %d != embedding width %d", what, d->in0, expect_in);
  PSN_REQUIRE(net->fwd[net->n_layers - 1].N <= S2_OUT_ROWS, PSN_ERR_SHAPE, "%s: output width %d > %d", what,
              net->fwd[net->n_layers - 1].N, S2_OUT_ROWS);
  PSN_REQUIRE(d->skip < net->n_layers - 1, PSN_ERR_SHAPE, "%s: skip after the last layer is unsupported", what);
  for (int l = 0; l < net->n_layers; ++l) d->fwd[l] = net->fwd[l];
  return PSN_OK;
}

size_t tc_vis_workspace_bytes(long long Ns, long long L);
size_t s2_workspace_bytes(long long Ns, long long L) {
  if (L < 1) L = 1;
  const size_t a = 256;
  size_t b = 0;
  b += ((size_t)Ns * 3 * 4 + a) * 2;           // normal, albedo per slot
  b += (size_t)Ns * 32 * 4 + a;                // weights per slot
  b += (size_t)Ns * L * 4 + a;                 // raw visibility per pair
  b += (size_t)Ns * 16 * 4 + a;                // slot_of_pixel upper bound is N; callers pass max(N, Ns) as n_rays
  return b + tc_vis_workspace_bytes(Ns, L) + 4096;
}

int s2_point_nets(const psn_mlp* normal_net, int nf_n, const psn_mlp* albedo_net, const psn_mlp* rough_net, int nf,
                  const float* pts, long long Ns, float* normal, float* albedo, float* weights, int nbt, cudaStream_t st) {
  PointNets pn;
  memset(&pn, 0, sizeof(pn));
  int rc;
  if ((rc = make_s2_dev(normal_net, 3 + 6 * nf_n, &pn.net[0], "normal_net"))) return rc;
  if ((rc = make_s2_dev(albedo_net, 3 + 6 * nf, &pn.net[1], "albedo_net"))) return rc;
  if ((rc = make_s2_dev(rough_net, 3 + 6 * nf, &pn.net[2], "rough_net"))) return rc;
  pn.nf[0] = nf_n; pn.nf[1] = nf; pn.nf[2] = nf;
  int k_rows = 16;
  for (int w = 0; w < 3; ++w) {
    PSN_REQUIRE(pn.net[w].n_layers == 0 || pn.net[w].in0 <= 64, PSN_ERR_SHAPE, "stage-2 point net: embedding width %d > 64",
                pn.net[w].in0);
    for (int l = 1; l < pn.net[w].n_layers; ++l) k_rows = pn.net[w].fwd[l].K_pad > k_rows ? pn.net[w].fwd[l].K_pad : k_rows;
  }
  if (rough_net) PSN_REQUIRE(rough_net->fwd[rough_net->n_layers - 1].N == nbt, PSN_ERR_SHAPE, "rough_net output %d != %d",
                             rough_net->fwd[rough_net->n_layers - 1].N, nbt);
  if (Ns == 0) return PSN_OK;
  const size_t smem = (size_t)(64 * LDX + (size_t)k_rows * LDX + WRING_FLOATS + S2_OUT_ROWS * LDX + 3 * TM) * 4;
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_s2_point, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (Ns + TM - 1) / TM;
  const int grid = (int)(tiles < num_ctas() ? tiles : num_ctas());
  psn::count_launch();
  k_s2_point<<<grid, NT, smem, st>>>(pn, pts, Ns, normal, albedo, weights, nbt, k_rows);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

int s2_visibility_simt(const psn_mlp* vis_net, int nf, const float* pts, long long Ns, const float* lights, int L, float* vis,
                       cudaStream_t st) {
  S2Dev d;
  int rc;
  if ((rc = make_s2_dev(vis_net, 2 * (3 + 6 * nf), &d, "visibility_net"))) return rc;
  PSN_REQUIRE(vis_net, PSN_ERR_ARG, "visibility_net is null");
  PSN_REQUIRE(d.in0 <= 128 && vis_net->fwd[vis_net->n_layers - 1].N == 1, PSN_ERR_SHAPE, "visibility_net shape unsupported");
  for (int l = 1; l < d.n_layers; ++l)
    PSN_REQUIRE(d.fwd[l].K_pad <= 384, PSN_ERR_SHAPE, "visibility_net layer %d input %d > 384", l, d.fwd[l].K);
  const long long pairs = Ns * L;
  if (pairs == 0) return PSN_OK;
  const size_t smem = (size_t)(128 * LDX + 384 * LDX + WRING_FLOATS + S2_OUT_ROWS * LDX + 6 * TM) * 4;
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_s2_vis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (pairs + TM - 1) / TM;
  const int grid = (int)(tiles < num_ctas() ? tiles : num_ctas());
  psn::count_launch();
  k_s2_vis<<<grid, NT, smem, st>>>(d, nf, pts, Ns, lights, pairs, vis);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

int tc_s2_visibility(const psn_mlp* vis_net, int nf, const float* pts, long long Ns, const float* lights, int L, float* vis,
                     void* ws, size_t ws_bytes, cudaStream_t st);  // tc path
int s2_shade_images(const float* n_s, const float* a_s, const float* w_s, const float* view, const float* v_raw, const float* lights,
                    const float* lobe, const float* intensity, const psn_shade_params* prm, const int32_t* pix, long long Ns, long long N,
                    int L, float* rgb, float* spec, float* vis, float* normal, float* albedo, float* sgw, int write_normal, int* sop,
                    cudaStream_t st);

}  // namespace psn

using namespace psn;

extern "C" int psn_s2_point_nets(const psn_mlp* albedo_net, const psn_mlp* rough_net, int n_freqs, const float* pts, int64_t Ns,
                                 float* albedo, float* weights, int nbt, int precision, void* stream) {
  PSN_REQUIRE(albedo_net && rough_net && (Ns == 0 || (pts && albedo && weights)), PSN_ERR_ARG, "psn_s2_point_nets: null argument");
  (void)precision;  // the per-point nets are <1% of the stage-2 work and always run on the fp32 path
  return s2_point_nets(nullptr, n_freqs, albedo_net, rough_net, n_freqs, pts, Ns, nullptr, albedo, weights, nbt,
                       (cudaStream_t)stream);
}

extern "C" int psn_s2_visibility(const psn_mlp* vis_net, int n_freqs, const float* pts, int64_t Ns, const float* lights, int L,
                                 float* vis, void* ws, int64_t ws_bytes, int precision, void* stream) {
  PSN_REQUIRE(vis_net && (Ns == 0 || L == 0 || (pts && lights && vis)), PSN_ERR_ARG, "psn_s2_visibility: null argument");
  if (prec_is_tc(precision)) return tc_s2_visibility(vis_net, n_freqs, pts, Ns, lights, L, vis, ws, (size_t)ws_bytes,
                                                        (cudaStream_t)stream);
  return s2_visibility_simt(vis_net, n_freqs, pts, Ns, lights, L, vis, (cudaStream_t)stream);
}

// dst[i, :] = src[:] for i < rows (material editing: one albedo / one SG weight vector for every surface point)
__global__ void k_broadcast_row(const float* __restrict__ src, int width, long long rows, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * width) dst[i] = src[i % width];
}

static int shade_stage2_impl(const psn_mlp* normal_net, const psn_mlp* albedo_net, const psn_mlp* rough_net,
                             const psn_mlp* vis_net, const float* lobe, const psn_shade_params* prm, const float* pts,
                             const float* view, const float* normal_in, const int32_t* pix, int64_t Ns, int64_t N,
                             const float* lights, int L, const float* intensity, const float* albedo_new, const float* weights_new,
                             float* rgb, float* spec, float* vis, float* normal, float* albedo, float* sgw, void* ws, int64_t ws_bytes,
                             int precision, void* stream);

extern "C" int psn_shade_stage2(const psn_mlp* normal_net, const psn_mlp* albedo_net, const psn_mlp* rough_net,
                                const psn_mlp* vis_net, const float* lobe, const psn_shade_params* prm, const float* pts,
                                const float* view, const float* normal_in, const int32_t* pix, int64_t Ns, int64_t N,
                                const float* lights, int L, const float* intensity, float* rgb, float* spec, float* vis,
                                float* normal, float* albedo, float* sgw, void* ws, int64_t ws_bytes, int precision, void* stream) {
  return shade_stage2_impl(normal_net, albedo_net, rough_net, vis_net, lobe, prm, pts, view, normal_in, pix, Ns, N, lights, L, intensity,
                           nullptr, nullptr, rgb, spec, vis, normal, albedo, sgw, ws, ws_bytes, precision, stream);
}

extern "C" int psn_shade_stage2_edit(const psn_mlp* normal_net, const psn_mlp* albedo_net, const psn_mlp* rough_net,
                                     const psn_mlp* vis_net, const float* lobe, const psn_shade_params* prm, const float* pts,
                                     const float* view, const float* normal_in, const int32_t* pix, int64_t Ns, int64_t N,
                                     const float* lights, int L, const float* intensity, const float* albedo_new,
                                     const float* weights_new, float* rgb, float* spec, float* vis, float* normal, float* albedo,
                                     float* sgw, void* ws, int64_t ws_bytes, int precision, void* stream) {
  return shade_stage2_impl(normal_net, albedo_net, rough_net, vis_net, lobe, prm, pts, view, normal_in, pix, Ns, N, lights, L, intensity,
                           albedo_new, weights_new, rgb, spec, vis, normal, albedo, sgw, ws, ws_bytes, precision, stream);
}

static int shade_stage2_impl(const psn_mlp* normal_net, const psn_mlp* albedo_net, const psn_mlp* rough_net,
                             const psn_mlp* vis_net, const float* lobe, const psn_shade_params* prm, const float* pts,
                             const float* view, const float* normal_in, const int32_t* pix, int64_t Ns, int64_t N,
                             const float* lights, int L, const float* intensity, const float* albedo_new, const float* weights_new,
                             float* rgb, float* spec, float* vis, float* normal, float* albedo, float* sgw, void* ws, int64_t ws_bytes,
                             int precision, void* stream) {
  PSN_REQUIRE(albedo_net && rough_net && prm && lights && rgb && spec && albedo, PSN_ERR_ARG, "psn_shade_stage2: null argument");
  PSN_REQUIRE(prm->render_model == 1 || (lobe && sgw), PSN_ERR_ARG, "psn_shade_stage2: sgbasis needs lobe and sgw");
  PSN_REQUIRE(prm->render_model == 0 || prm->render_model == 1, PSN_ERR_ARG, "psn_shade_stage2: render_model %d", prm->render_model);
  PSN_REQUIRE(Ns == 0 || (pts && view && pix), PSN_ERR_ARG, "psn_shade_stage2: null surface inputs");
  PSN_REQUIRE(normal_net || normal_in || Ns == 0, PSN_ERR_ARG, "psn_shade_stage2: need normal_net or normal_in");
  PSN_REQUIRE(!normal_net || normal, PSN_ERR_ARG, "psn_shade_stage2: normal output required with normal_net");
  PSN_REQUIRE(!vis_net || vis, PSN_ERR_ARG, "psn_shade_stage2: vis output required with visibility_net");
  PSN_REQUIRE(L >= 1 && N >= Ns, PSN_ERR_ARG, "psn_shade_stage2: L=%d N=%lld Ns=%lld", L, (long long)N, (long long)Ns);
  PSN_REQUIRE(prm->intensity_kind == 0 || intensity, PSN_ERR_ARG, "psn_shade_stage2: per-light intensity pointer is null");
  const int nbt = prm->render_model == 1 ? 1 : (prm->specular_rgb ? 3 * prm->nbasis : prm->nbasis);
  PSN_REQUIRE(nbt <= 32, PSN_ERR_SHAPE, "psn_shade_stage2: %d SG weights > 32", nbt);
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  float* n_s = w.take<float>((size_t)Ns * 3 + 4);
  float* a_s = w.take<float>((size_t)Ns * 3 + 4);
  float* w_s = w.take<float>((size_t)Ns * nbt + 4);
  float* v_s = vis_net ? w.take<float>((size_t)Ns * L + 4) : nullptr;
  int* sop = w.take<int>((size_t)N + 4);
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "psn_shade_stage2: workspace too small (need %zu bytes, have %lld)", w.used,
              (long long)ws_bytes);
  int rc;
  {
    ProfScope prof(PSN_PROF_S2_POINT, Ns, st);
    if ((rc = s2_point_nets(normal_net, prm->n_freqs_normal, albedo_net, rough_net, prm->n_freqs_xyz, pts, Ns, n_s, a_s, w_s,
                            nbt, st)))
      return rc;
  }
  if (Ns > 0 && albedo_new) {  // renderer.py:167-168: every surface point gets the edited albedo
    psn::count_launch();
    k_broadcast_row<<<(unsigned)((Ns * 3 + 255) / 256), 256, 0, st>>>(albedo_new, 3, Ns, a_s);
  }
  if (Ns > 0 && weights_new) {  // renderer.py:175-181: ... and the edited SG weights
    psn::count_launch();
    k_broadcast_row<<<(unsigned)((Ns * nbt + 255) / 256), 256, 0, st>>>(weights_new, nbt, Ns, w_s);
  }
  if (vis_net) {
    ProfScope prof(PSN_PROF_S2_VIS, (long long)Ns * L, st);
    if (prec_is_tc(precision)) {
      const size_t off = (w.used + 255) / 256 * 256;
      rc = tc_s2_visibility(vis_net, prm->n_freqs_xyz, pts, Ns, lights, L, v_s, (char*)ws + off,
                            (size_t)ws_bytes > off ? (size_t)ws_bytes - off : 0, st);
    } else {
      rc = s2_visibility_simt(vis_net, prm->n_freqs_xyz, pts, Ns, lights, L, v_s, st);
    }
    if (rc) return rc;
  }
  return s2_shade_images(normal_net ? n_s : normal_in, a_s, w_s, view, v_s, lights, lobe, intensity, prm, pix, Ns, N, L, rgb, spec,
                         vis_net ? vis : nullptr, normal, albedo, sgw, normal_net ? 1 : 0, sop, st);
}

namespace psn {
// SG shading + image-shaped writes from per-surface-point quantities (shared by the inference and the train-step forward).
int s2_shade_images(const float* n_s, const float* a_s, const float* w_s, const float* view, const float* v_raw, const float* lights,
                    const float* lobe, const float* intensity, const psn_shade_params* prm, const int32_t* pix, long long Ns, long long N,
                    int L, float* rgb, float* spec, float* vis, float* normal, float* albedo, float* sgw, int write_normal, int* sop,
                    cudaStream_t st) {
  const int nbt = prm->render_model == 1 ? 1 : (prm->specular_rgb ? 3 * prm->nbasis : prm->nbasis);
  psn::count_launch();
  k_fill_int<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(sop, N, -1);
  if (Ns > 0) { psn::count_launch(); k_slot_of_pixel<<<(unsigned)((Ns + 255) / 256), 256, 0, st>>>(pix, Ns, sop); }
  ShadeArgs a;
  memset(&a, 0, sizeof(a));
  a.normal = n_s;
  a.albedo = a_s; a.weights = w_s; a.view = view; a.vis = v_raw; a.lights = lights; a.lobe = lobe; a.intensity = intensity;
  a.slot_of_pixel = sop;
  a.rgb = rgb; a.spec = spec; a.vis_out = v_raw ? vis : nullptr; a.normal_out = normal; a.albedo_out = albedo; a.sgw_out = sgw;
  a.N = N; a.Ns = Ns; a.L = L; a.nbasis = prm->nbasis; a.specular_rgb = prm->specular_rgb; a.nbt = nbt;
  a.intensity_kind = prm->intensity_kind; a.intensity_scalar = prm->intensity; a.write_normal = write_normal;
  a.microfacet = prm->render_model == 1 ? 1 : 0;
  a.f0 = prm->fresnel_f0;
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)(L + 1));
  psn::count_launch();
  k_s2_shade<<<grid, 256, 0, st>>>(a);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
}  // namespace psn

