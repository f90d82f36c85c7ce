// Stage-1 field evaluation on fp32 FFMA: device-side composition of the dense() primitive into
//   geo forward (PE -> 8 softplus layers, optional sigma' stash) / logit + feature heads /
//   analytic-normal reverse pass / appearance MLP.
// Reference semantics: stage1/model/network.py:85-136 (see oracle/psnerf_oracle.py for the restatement).
#pragma once
#include "simt_mlp.cuh"

namespace psn {

constexpr int PE_ROWS = 48;  // padded rows of the point-encoding buffer (39 used at 6 octaves)

struct GeoDev {
  SimtLayer fwd[kMaxLayers];  // fwd[0..n_hidden-1] hidden layers, fwd[n_hidden] = feature head
  SimtLayer rev[kMaxLayers];
  SimtLayer logit;
  const float* w_row;
  int n_hidden;  // number of softplus layers (8)
  int skip;      // layer whose input is cat[x, pe]/sqrt2, or -1
  int octaves, pe_dim;
  float rescale;
};

struct AppDev {
  SimtLayer fwd[kMaxLayers];
  int n_layers;  // 5
  int octaves_view, pe_view_dim, feat_off, in_dim;
};

// How a kernel obtains the i-th evaluation point.
enum { GEN_EXPLICIT = 0, GEN_MARCH = 1, GEN_SHADOW = 2, GEN_INDEXED_DEPTH = 3, GEN_RAY_DEPTH = 4, GEN_SHADOW_LIST = 5 };
constexpr int SHADOW_LIST_STEP_BITS = 7;  // GEN_SHADOW_LIST entry = (pair << 7) | step: at most 128 steps per shadow ray
struct PointGen {
  int kind;
  const float* pts;    // EXPLICIT: [M,3]
  const float* views;  // EXPLICIT radiance: [M,3] view dirs (ray_d)
  const float* dirs;   // ray directions [N,3]
  const float* far;    // MARCH: sphere far depth per ray [N]
  const float* depth;  // INDEXED_DEPTH: depth per list entry [M]; RAY_DEPTH: [N*S] sample depths
  const int* index;    // INDEXED_DEPTH: ray id per list entry; SHADOW_LIST: packed (pair, step) entries (k_shadow_plan)
  const float* surf;   // SHADOW: [Ns,3]
  const float* lights; // SHADOW: [L,3]
  float o[3];
  float near_, lnear, lfar;
  int S;
  long long Ns;
};

// torch.linspace(0,1,S)[s] in float32 (ATen RangeFactories: symmetric evaluation about the midpoint).
__device__ __forceinline__ float linspace01(int s, int S) {
  const float step = 1.0f / (float)(S - 1);
  return (s < S / 2) ? __fmul_rn(step, (float)s) : __fsub_rn(1.0f, __fmul_rn(step, (float)(S - s - 1)));
}
// near*(1-t) + far*t, separate roundings as the reference's tensor expression (rendering.py:447,392)
__device__ __forceinline__ float lerp_depth(float a, float b, float t) {
  return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, t)), __fmul_rn(b, t));
}
__device__ __forceinline__ float madd_rn(float o, float d, float t) { return __fadd_rn(o, __fmul_rn(d, t)); }

__device__ __forceinline__ void gen_point(const PointGen& g, long long i, float p[3], float v[3]) {
  v[0] = v[1] = v[2] = 0.f;
  if (g.kind == GEN_EXPLICIT) {
    p[0] = g.pts[i * 3 + 0]; p[1] = g.pts[i * 3 + 1]; p[2] = g.pts[i * 3 + 2];
    if (g.views) { v[0] = g.views[i * 3 + 0]; v[1] = g.views[i * 3 + 1]; v[2] = g.views[i * 3 + 2]; }
  } else if (g.kind == GEN_MARCH) {
    const long long r = i / g.S;
    const int s = (int)(i - r * g.S);
    const float d = lerp_depth(g.near_, g.far[r], linspace01(s, g.S));
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = madd_rn(g.o[c], g.dirs[r * 3 + c], d);
  } else if (g.kind == GEN_SHADOW) {
    const long long pair = i / g.S;  // light-major: pair = l*Ns + n
    const int s = (int)(i - pair * g.S);
    const long long l = pair / g.Ns, n = pair - l * g.Ns;
    const float d = lerp_depth(g.lnear, g.lfar, linspace01(s, g.S));
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = madd_rn(g.surf[n * 3 + c], g.lights[l * 3 + c], d);
  } else if (g.kind == GEN_SHADOW_LIST) {
    // the in-box samples of the shadow rays only (stage1_aux.cu:k_shadow_plan); same arithmetic as GEN_SHADOW
    const unsigned e = (unsigned)g.index[i];
    const long long pair = e >> SHADOW_LIST_STEP_BITS;
    const int s = (int)(e & ((1u << SHADOW_LIST_STEP_BITS) - 1u));
    const long long l = pair / g.Ns, n = pair - l * g.Ns;
    const float d = lerp_depth(g.lnear, g.lfar, linspace01(s, g.S));
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = madd_rn(g.surf[n * 3 + c], g.lights[l * 3 + c], d);
  } else if (g.kind == GEN_INDEXED_DEPTH) {
    const long long r = g.index[i];
    const float d = g.depth[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = madd_rn(g.o[c], g.dirs[r * 3 + c], d);
  } else {  // GEN_RAY_DEPTH
    const long long r = i / g.S;
    const float d = g.depth[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float dc = g.dirs[r * 3 + c];
      p[c] = madd_rn(g.o[c], dc, d);
      v[c] = -dc;
    }
  }
}

// PE[k][row]: [p, sin(2^0 p), cos(2^0 p), ...] of p/rescale, zero padded to PE_ROWS rows (network.py:141-150).
__device__ __forceinline__ void encode_points(const float* P /*[3][TM]*/, float* PE, int octaves, float rescale) {
  const int r = threadIdx.x & (TM - 1), q = threadIdx.x / TM;  // 4 threads per row
  const float x[3] = {P[r] / rescale, P[TM + r] / rescale, P[2 * TM + r] / rescale};
  if (q == 0) {
    PE[0 * LDX + r] = x[0]; PE[1 * LDX + r] = x[1]; PE[2 * LDX + r] = x[2];
  }
  for (int idx = q; idx < octaves * 3; idx += NT / TM) {
    const int i = idx / 3, c = idx - 3 * i;
    float s, co;
    sincosf((float)(1 << i) * x[c], &s, &co);
    PE[(3 + 6 * i + c) * LDX + r] = s;
    PE[(6 + 6 * i + c) * LDX + r] = co;
  }
  for (int k = 3 + 6 * octaves + q; k < PE_ROWS; k += NT / TM) PE[k * LDX + r] = 0.f;
}

#define PSN_SQRT2 1.41421356237309504880f

// 8 softplus layers.  X ends up holding softplus(z_{n_hidden-1}).  With STASH, sigma'(z_l) = sigmoid(100 z_l)
// of every hidden unit goes to the per-CTA global stash in the thread's own register order.
template <bool STASH>
__device__ __forceinline__ void geo_forward(const GeoDev& g, const float* PE, float* X, float* WS, float4* stash) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int l = 0; l < g.n_hidden; ++l) {
    float acc[8][8];
    dense<8>(g.fwd[l], l == 0 ? PE : X, WS, acc);
    const bool pre_skip = (l + 1 == g.skip);
    const int N = g.fwd[l].N;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = simt_col<8>(tx, j);
      float v[8], s[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = acc[i][j];
        float a = softplus100(z);
        if (pre_skip) a = a / PSN_SQRT2;
        v[i] = a;
        if (STASH) s[i] = sigmoidf_(100.f * z);
      }
      if (col < N) store_col8(X, col, ty, v);
      if (STASH) {
        stash[(size_t)(l * 16 + j * 2 + 0) * NT + threadIdx.x] = make_float4(s[0], s[1], s[2], s[3]);
        stash[(size_t)(l * 16 + j * 2 + 1) * NT + threadIdx.x] = make_float4(s[4], s[5], s[6], s[7]);
      }
    }
    if (pre_skip) {  // rows N.. of the next layer's input are pe/sqrt2 (network.py:90-91), then zero padding
      const int kp = g.fwd[l + 1].K_pad;
      for (int idx = threadIdx.x; idx < (kp - N) * TM; idx += NT) {
        const int k = idx / TM, r = idx - k * TM;
        X[(size_t)(N + k) * LDX + r] = (k < g.pe_dim) ? PE[k * LDX + r] / PSN_SQRT2 : 0.f;
      }
    }
    __syncthreads();
  }
}

// d logit / d p from the stash.  X is clobbered; GPE is a [PE_ROWS][LDX] scratch; result in G3[3][TM].
__device__ __forceinline__ void geo_reverse(const GeoDev& g, const float* P, float* X, float* GPE, float* WS,
                                            const float4* stash, float* G3) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int top = g.n_hidden - 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {  // dz_top = W_last[0,:] * sigma'(z_top)
    const int col = simt_col<8>(tx, j);
    const float w = g.w_row[col];
    const float4 s0 = stash[(size_t)(top * 16 + j * 2 + 0) * NT + threadIdx.x];
    const float4 s1 = stash[(size_t)(top * 16 + j * 2 + 1) * NT + threadIdx.x];
    const float v[8] = {w * s0.x, w * s0.y, w * s0.z, w * s0.w, w * s1.x, w * s1.y, w * s1.z, w * s1.w};
    store_col8(X, col, ty, v);
  }
  for (int idx = threadIdx.x; idx < PE_ROWS * LDX; idx += NT) GPE[idx] = 0.f;
  __syncthreads();
  for (int l = top; l >= 1; --l) {
    float acc[8][8];
    dense<8>(g.rev[l], X, WS, acc);
    const bool is_skip = (l == g.skip);
    const int nprev = g.fwd[l - 1].N;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = simt_col<8>(tx, j);
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = is_skip ? acc[i][j] / PSN_SQRT2 : acc[i][j];
      if (col < nprev) {
        const float4 s0 = stash[(size_t)((l - 1) * 16 + j * 2 + 0) * NT + threadIdx.x];
        const float4 s1 = stash[(size_t)((l - 1) * 16 + j * 2 + 1) * NT + threadIdx.x];
        v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
        v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
        store_col8(X, col, ty, v);
      } else if (is_skip && col - nprev < g.pe_dim) {
        store_col8(GPE, col - nprev, ty, v);
      }
    }
    if (is_skip) {  // zero the K padding of the next reverse GEMM (its K is nprev)
      const int kp = g.rev[l - 1].K_pad;
      for (int idx = threadIdx.x; idx < (kp - nprev) * TM; idx += NT) {
        const int k = idx / TM, r = idx - k * TM;
        X[(size_t)(nprev + k) * LDX + r] = 0.f;
      }
    }
    __syncthreads();
  }
  {
    float acc[8][2];
    dense<2>(g.rev[0], X, WS, acc);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = tx * 2 + j;
      if (col < g.pe_dim) {
#pragma unroll
        for (int i = 0; i < 8; ++i) GPE[col * LDX + ty * 8 + i] += acc[i][j];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < TM) {  // J_pe^T: d/dp [p, sin(f p), cos(f p)] = [1, f cos(f p), -f sin(f p)]
    const int r = threadIdx.x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = P[c * TM + r] / g.rescale;
      float acc = GPE[c * LDX + r];
      for (int i = 0; i < g.octaves; ++i) {
        const float f = (float)(1 << i);
        float s, co;
        sincosf(f * x, &s, &co);
        acc += f * co * GPE[(3 + 6 * i + c) * LDX + r] - f * s * GPE[(6 + 6 * i + c) * LDX + r];
      }
      G3[c * TM + r] = acc / g.rescale;
    }
  }
  __syncthreads();
}

}  // namespace psn
