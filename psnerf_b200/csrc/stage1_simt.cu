// Stage-1 fp32 FFMA kernels: occupancy / infer_occ / gradient / radiance over 64-row tiles, persistent CTAs.
#include "stage1_simt.cuh"
#include "launch.cuh"
#include "internal.cuh"

namespace psn {

// ---------------------------------------------------------------------------------------------------------
// occupancy (+ optional feature head)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_geo_occ(GeoDev g, PointGen gen, long long M_host, const int* M_dev, int out_kind, float* out, int with_feat) {
  extern __shared__ __align__(16) float smem[];
  float* PE = smem;
  float* X = PE + PE_ROWS * LDX;
  float* WS = X + 256 * LDX;
  float* P = WS + WRING_FLOATS;
  const long long M = M_dev ? (long long)*M_dev : M_host;
  const long long n_tiles = (M + TM - 1) / TM;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * TM;
    if (threadIdx.x < TM) {
      float p[3] = {0.f, 0.f, 0.f}, v[3];
      if (base + threadIdx.x < M) gen_point(gen, base + threadIdx.x, p, v);
      P[threadIdx.x] = p[0]; P[TM + threadIdx.x] = p[1]; P[2 * TM + threadIdx.x] = p[2];
    }
    __syncthreads();
    encode_points(P, PE, g.octaves, g.rescale);
    __syncthreads();
    geo_forward<false>(g, PE, X, WS, nullptr);
    {
      float acc[8][1];
      dense<1>(g.logit, X, WS, acc);
      if (tx == 0) {
        const int stride = with_feat ? (1 + g.fwd[g.n_hidden].N) : 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long idx = base + ty * 8 + i;
          if (idx < M) {
            const float z = acc[i][0];
            float o = z;
            if (!with_feat) {
              if (out_kind == PSN_OUT_ALPHA) o = sigmoidf_(z * -10.0f);
              else if (out_kind == PSN_OUT_NEG_LOGIT) o = -1.f * z;
            }
            out[idx * stride] = o;
          }
        }
      }
    }
    if (with_feat) {
      float acc[8][8];
      dense<8>(g.fwd[g.n_hidden], X, WS, acc);
      const int F = g.fwd[g.n_hidden].N;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = simt_col<8>(tx, j);
        if (col < F) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long idx = base + ty * 8 + i;
            if (idx < M) out[idx * (1 + F) + 1 + col] = acc[i][j];
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// secant refinement (rendering.py:525-555): n_iter evaluations per tile of 64 masked rays, the bracket stays in shared memory
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_geo_secant(GeoDev g, PointGen gen, SecantState sec, int n_iter, float tau) {
  extern __shared__ __align__(16) float smem[];
  float* PE = smem;
  float* X = PE + PE_ROWS * LDX;
  float* WS = X + 256 * LDX;
  float* P = WS + WRING_FLOATS;   // [3][TM]
  float* ST = P + 3 * TM;         // d_low, d_high, f_low, f_high, d_pred, dir x / y / z, z (logit): [9][TM]
  const long long M = (long long)*sec.count;
  const long long n_tiles = (M + TM - 1) / TM;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * TM;
    const int r = threadIdx.x;
    const bool live = r < TM && base + r < M;
    if (r < TM) {
      float d_low = 0.f, d_high = 0.f, f_low = -1.f, f_high = 1.f, d_pred = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
      if (live) {
        const long long slot = base + r, ray = sec.ray[slot];
        d_low = sec.d_low[slot]; d_high = sec.d_high[slot]; f_low = sec.f_low[slot]; f_high = sec.f_high[slot];
        d_pred = sec.d_pred[slot];
        dx = gen.dirs[ray * 3]; dy = gen.dirs[ray * 3 + 1]; dz = gen.dirs[ray * 3 + 2];
      }
      ST[r] = d_low; ST[TM + r] = d_high; ST[2 * TM + r] = f_low; ST[3 * TM + r] = f_high; ST[4 * TM + r] = d_pred;
      ST[5 * TM + r] = dx; ST[6 * TM + r] = dy; ST[7 * TM + r] = dz;
    }
    __syncthreads();
    for (int it = 0; it < n_iter; ++it) {
      if (r < TM) {  // p_mid = ray0 + d_pred * ray_direction (the arithmetic of GEN_INDEXED_DEPTH)
        const float d = ST[4 * TM + r];
        P[r] = madd_rn(gen.o[0], ST[5 * TM + r], d);
        P[TM + r] = madd_rn(gen.o[1], ST[6 * TM + r], d);
        P[2 * TM + r] = madd_rn(gen.o[2], ST[7 * TM + r], d);
      }
      __syncthreads();
      encode_points(P, PE, g.octaves, g.rescale);
      __syncthreads();
      geo_forward<false>(g, PE, X, WS, nullptr);
      {
        float acc[8][1];
        dense<1>(g.logit, X, WS, acc);
        if (tx == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) ST[8 * TM + ty * 8 + i] = acc[i][0];
        }
      }
      __syncthreads();
      if (r < TM) {  // f_mid = occupancy - tau with the PSN_OUT_ALPHA expression of k_geo_occ, then the update of k_secant_update
        const float f_mid = sigmoidf_(ST[8 * TM + r] * -10.0f) - tau;
        float d_low = ST[r], d_high = ST[TM + r], f_low = ST[2 * TM + r], f_high = ST[3 * TM + r];
        const float d_pred = ST[4 * TM + r];
        if (f_mid < 0.f) { d_low = d_pred; f_low = f_mid; } else { d_high = d_pred; f_high = f_mid; }
        ST[r] = d_low; ST[TM + r] = d_high; ST[2 * TM + r] = f_low; ST[3 * TM + r] = f_high;
        ST[4 * TM + r] = __fadd_rn(__fdiv_rn(__fmul_rn(-f_low, __fsub_rn(d_high, d_low)), __fsub_rn(f_high, f_low)), d_low);
      }
      __syncthreads();
    }
    if (live) {
      const long long slot = base + r;
      sec.d_low[slot] = ST[r]; sec.d_high[slot] = ST[TM + r]; sec.f_low[slot] = ST[2 * TM + r]; sec.f_high[slot] = ST[3 * TM + r];
      sec.d_pred[slot] = ST[4 * TM + r];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// analytic gradient
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_geo_grad(GeoDev g, PointGen gen, long long M_host, const int* M_dev, float* grad, float4* stash_all) {
  extern __shared__ __align__(16) float smem[];
  float* PE = smem;
  float* X = PE + PE_ROWS * LDX;
  float* GPE = X + 256 * LDX;
  float* WS = GPE + PE_ROWS * LDX;
  float* P = WS + WRING_FLOATS;
  float* G3 = P + 3 * TM;
  float4* stash = stash_all + (size_t)blockIdx.x * (kMaxLayers * 16 * NT);
  const long long M = M_dev ? (long long)*M_dev : M_host;
  const long long n_tiles = (M + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * TM;
    if (threadIdx.x < TM) {
      float p[3] = {0.f, 0.f, 0.f}, v[3];
      if (base + threadIdx.x < M) gen_point(gen, base + threadIdx.x, p, v);
      P[threadIdx.x] = p[0]; P[TM + threadIdx.x] = p[1]; P[2 * TM + threadIdx.x] = p[2];
    }
    __syncthreads();
    encode_points(P, PE, g.octaves, g.rescale);
    __syncthreads();
    geo_forward<true>(g, PE, X, WS, stash);
    geo_reverse(g, P, X, GPE, WS, stash, G3);
    if (threadIdx.x < TM && base + threadIdx.x < M) {
      const long long idx = base + threadIdx.x;
      grad[idx * 3 + 0] = G3[threadIdx.x];
      grad[idx * 3 + 1] = G3[TM + threadIdx.x];
      grad[idx * 3 + 2] = G3[2 * TM + threadIdx.x];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// radiance sample: geo fwd (stash) -> logit/feat heads -> reverse -> app MLP -> (rgb, alpha)
// ---------------------------------------------------------------------------------------------------------
constexpr int APP_ROWS = 304;  // 289 inputs padded to a multiple of KC

__global__ void __launch_bounds__(NT, 1)
k_radiance(GeoDev g, AppDev a, PointGen gen, long long M_host, const int* M_dev, float* rgb, float* alpha,
           float4* stash_all) {
  extern __shared__ __align__(16) float smem[];
  float* PE = smem;
  float* X = PE + PE_ROWS * LDX;
  float* AIN = X + 256 * LDX;           // [APP_ROWS][LDX] app-MLP input, later its activations
  float* WS = AIN + APP_ROWS * LDX;
  float* P = WS + WRING_FLOATS;         // [3][TM]
  float* V = P + 3 * TM;                // [3][TM] view dirs
  float* G3 = V + 3 * TM;               // [3][TM]
  float* AL = G3 + 3 * TM;              // [TM] alpha
  float* GPE = AL + TM;                 // [PE_ROWS][LDX] reverse-pass scratch (d logit / d pe)
  float4* stash = stash_all + (size_t)blockIdx.x * (kMaxLayers * 16 * NT);
  const long long M = M_dev ? (long long)*M_dev : M_host;
  const long long n_tiles = (M + TM - 1) / TM;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * TM;
    if (threadIdx.x < TM) {
      float p[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 1.f};
      if (base + threadIdx.x < M) gen_point(gen, base + threadIdx.x, p, v);
      P[threadIdx.x] = p[0]; P[TM + threadIdx.x] = p[1]; P[2 * TM + threadIdx.x] = p[2];
      V[threadIdx.x] = v[0]; V[TM + threadIdx.x] = v[1]; V[2 * TM + threadIdx.x] = v[2];
    }
    __syncthreads();
    encode_points(P, PE, g.octaves, g.rescale);
    __syncthreads();
    geo_forward<true>(g, PE, X, WS, stash);
    {  // logit head -> alpha = sigmoid(-10 logit) (network.py:134)
      float acc[8][1];
      dense<1>(g.logit, X, WS, acc);
      if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) AL[ty * 8 + i] = sigmoidf_(acc[i][0] * -10.0f);
      }
    }
    {  // feature head (no activation) straight into the app-MLP input rows
      float acc[8][8];
      dense<8>(g.fwd[g.n_hidden], X, WS, acc);
      const int F = g.fwd[g.n_hidden].N;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = simt_col<8>(tx, j);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[i][j];
        if (col < F) store_col8(AIN, a.feat_off + col, ty, v);
      }
    }
    __syncthreads();
    geo_reverse(g, P, X, GPE, WS, stash, G3);
    // assemble cat[p, PE(view/|view|), gradient, feat] (network.py:98,127-132)
    if (threadIdx.x < TM) {
      const int r = threadIdx.x;
      const float vx = V[r], vy = V[TM + r], vz = V[2 * TM + r];
      const float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
      const float v[3] = {vx / nrm, vy / nrm, vz / nrm};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        AIN[c * LDX + r] = P[c * TM + r];
        AIN[(3 + c) * LDX + r] = v[c];
        for (int i = 0; i < a.octaves_view; ++i) {
          float s, co;
          sincosf((float)(1 << i) * v[c], &s, &co);
          AIN[(3 + 3 + 6 * i + c) * LDX + r] = s;
          AIN[(3 + 6 + 6 * i + c) * LDX + r] = co;
        }
        AIN[(3 + a.pe_view_dim + c) * LDX + r] = G3[c * TM + r];
      }
      for (int k = a.in_dim; k < a.fwd[0].K_pad; ++k) AIN[k * LDX + r] = 0.f;
    }
    __syncthreads();
    for (int l = 0; l < a.n_layers - 1; ++l) {  // ReLU layers, in place
      float acc[8][8];
      dense<8>(a.fwd[l], AIN, WS, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = simt_col<8>(tx, j);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(acc[i][j], 0.f);
        if (col < a.fwd[l].N) store_col8(AIN, col, ty, v);
      }
      __syncthreads();
    }
    {
      float acc[8][1];
      dense<1>(a.fwd[a.n_layers - 1], AIN, WS, acc);
      if (tx < 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long idx = base + ty * 8 + i;
          if (idx < M) rgb[idx * 3 + tx] = tanhf(acc[i][0]) * 0.5f + 0.5f;
        }
      }
      if (threadIdx.x < TM && base + threadIdx.x < M) alpha[base + threadIdx.x] = AL[threadIdx.x];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
int make_geo_dev(const psn_mlp* net, GeoDev* g) {
  PSN_REQUIRE(net && net->kind == PSN_NET_GEO, PSN_ERR_ARG, "expected a PSN_NET_GEO handle");
  memset(g, 0, sizeof(*g));
  g->n_hidden = net->n_layers - 1;
  for (int l = 0; l < net->n_layers; ++l) g->fwd[l] = net->fwd[l];
  for (int l = 0; l < net->n_layers - 1; ++l) g->rev[l] = net->rev[l];
  g->logit = net->logit_head;
  g->w_row = net->w_logit_row;
  g->skip = net->desc.skip;
  g->octaves = net->desc.octaves;
  g->pe_dim = 3 + 6 * net->desc.octaves;
  g->rescale = net->desc.rescale;
  PSN_REQUIRE(g->pe_dim <= PE_ROWS && net->in_dims[0] == g->pe_dim, PSN_ERR_SHAPE,
              "geo net: first layer input %d != 3+6*octaves (octaves=%d, max 7)", net->in_dims[0], g->octaves);
  for (int l = 0; l < g->n_hidden; ++l)
    PSN_REQUIRE(net->fwd[l].N_pad == 256 && net->fwd[l].K <= 256, PSN_ERR_SHAPE,
                "geo net: hidden layer %d is [%d,%d]; this build supports hidden widths in (128,256]", l,
                net->fwd[l].N, net->fwd[l].K);
  if (g->skip >= 0) {
    PSN_REQUIRE(g->skip >= 1 && g->skip < g->n_hidden, PSN_ERR_SHAPE, "geo net: skip=%d out of range", g->skip);
    PSN_REQUIRE(net->out_dims[g->skip - 1] + g->pe_dim == net->in_dims[g->skip], PSN_ERR_SHAPE,
                "geo net: skip layer input %d != %d + pe %d", net->in_dims[g->skip], net->out_dims[g->skip - 1],
                g->pe_dim);
  }
  return PSN_OK;
}

int make_app_dev(const psn_mlp* net, const GeoDev& g, AppDev* a) {
  PSN_REQUIRE(net && net->kind == PSN_NET_APP, PSN_ERR_ARG, "expected a PSN_NET_APP handle");
  memset(a, 0, sizeof(*a));
  a->n_layers = net->n_layers;
  for (int l = 0; l < net->n_layers; ++l) a->fwd[l] = net->fwd[l];
  a->octaves_view = net->desc.octaves;
  a->pe_view_dim = 3 + 6 * net->desc.octaves;
  a->feat_off = 3 + a->pe_view_dim + 3;
  a->in_dim = a->feat_off + g.fwd[g.n_hidden].N;
  PSN_REQUIRE(net->in_dims[0] == a->in_dim && net->fwd[0].K_pad <= APP_ROWS, PSN_ERR_SHAPE,
              "app net: first layer input %d != 3+pe_view(%d)+3+feat(%d)", net->in_dims[0], a->pe_view_dim,
              g.fwd[g.n_hidden].N);
  for (int l = 0; l < net->n_layers - 1; ++l)
    PSN_REQUIRE(net->fwd[l].N_pad == 256, PSN_ERR_SHAPE, "app net: hidden layer %d width %d unsupported", l,
                net->fwd[l].N);
  PSN_REQUIRE(net->fwd[net->n_layers - 1].N == 3, PSN_ERR_SHAPE, "app net: output dim must be 3");
  return PSN_OK;
}

size_t simt_stash_bytes() { return (size_t)num_ctas() * kMaxLayers * 16 * NT * sizeof(float4); }

static size_t smem_occ() { return (size_t)(PE_ROWS * LDX + 256 * LDX + WRING_FLOATS + 3 * TM) * 4; }
static size_t smem_grad() { return (size_t)(2 * PE_ROWS * LDX + 256 * LDX + WRING_FLOATS + 6 * TM) * 4; }
static size_t smem_rad() {
  return (size_t)(PE_ROWS * LDX + 256 * LDX + APP_ROWS * LDX + WRING_FLOATS + 10 * TM + PE_ROWS * LDX) * 4;
}

int simt_occupancy(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, int out_kind, float* out,
                   int with_feat, cudaStream_t st) {
  GeoDev g;
  int rc = make_geo_dev(geo, &g);
  if (rc) return rc;
  if (M == 0 && !M_dev) return PSN_OK;
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_geo_occ, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_occ()));
  const long long tiles = M_dev ? (long long)num_ctas() : (M + TM - 1) / TM;
  const int grid = (int)(tiles < num_ctas() ? tiles : num_ctas());
  psn::count_launch();
  k_geo_occ<<<grid, NT, smem_occ(), st>>>(g, gen, M, M_dev, out_kind, out, with_feat);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

// n_iter secant iterations over the masked rays of `sec` in ONE launch (count read on the device)
int simt_secant(const psn_mlp* geo, const PointGen& gen, const SecantState& sec, int n_iter, float tau, cudaStream_t st) {
  GeoDev g;
  int rc = make_geo_dev(geo, &g);
  if (rc) return rc;
  if (n_iter <= 0) return PSN_OK;
  const size_t smem = smem_occ() + 9 * TM * 4;
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_geo_secant, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  psn::count_launch();
  k_geo_secant<<<num_ctas(), NT, smem, st>>>(g, gen, sec, n_iter, tau);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

int simt_gradient(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, float* grad, void* stash,
                  cudaStream_t st) {
  GeoDev g;
  int rc = make_geo_dev(geo, &g);
  if (rc) return rc;
  if (M == 0 && !M_dev) return PSN_OK;
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_geo_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_grad()));
  const long long tiles = M_dev ? (long long)num_ctas() : (M + TM - 1) / TM;
  const int grid = (int)(tiles < num_ctas() ? tiles : num_ctas());
  psn::count_launch();
  k_geo_grad<<<grid, NT, smem_grad(), st>>>(g, gen, M, M_dev, grad, (float4*)stash);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

int simt_radiance(const psn_mlp* geo, const psn_mlp* app, const PointGen& gen, long long M, float* rgb, float* alpha,
                  void* stash, cudaStream_t st) {
  GeoDev g;
  AppDev a;
  int rc = make_geo_dev(geo, &g);
  if (rc) return rc;
  rc = make_app_dev(app, g, &a);
  if (rc) return rc;
  if (M == 0) return PSN_OK;
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_radiance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rad()));
  const long long tiles = (M + TM - 1) / TM;
  const int grid = (int)(tiles < num_ctas() ? tiles : num_ctas());
  psn::count_launch();
  k_radiance<<<grid, NT, smem_rad(), st>>>(g, a, gen, M, nullptr, rgb, alpha, (float4*)stash);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

}  // namespace psn
