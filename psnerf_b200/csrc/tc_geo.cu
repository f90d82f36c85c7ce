// tcgen05 occupancy kernels of the stage-1 geo MLP: PE -> 8 softplus layers on tensor cores -> fp32 logit head.
//   MODE_OUT    : one value per sample (alpha / logits) - ray-march proposals, secant points, explicit queries
//   MODE_SHADOW : a 128-row tile is one shadow ray; box mask + transmittance are reduced in the epilogue and only
//                 vis[pair] is written (rendering.py:378-408), no per-sample HBM traffic at all.
//   MODE_FEAT   : NeuralNetwork.infer_occ (network.py:85-95): a ninth step, the feature head, follows the stack; out[M, 1 + 256] =
//                 [logit, feature] per sample.
// Algorithmic work per sample: 2 * 459,008 FLOP (the 256 unused feature rows of the last layer are not computed).
// CHEAP = true is the first level of the two-level surface march (PSN_PREC_TC_TWOLEVEL, api_stage1.cu): every layer in ONE pass
// A_hi W_hi (Step::single, hi-only operand stores).  Its values are only trusted away from the occupancy threshold; the
// CHEAP = false instantiations are what every other caller runs.
#include "tc_mlp.cuh"
#include "stage1_simt.cuh"
#include "launch.cuh"
#include "internal.cuh"

namespace psn {
using namespace tc;

struct TcGeoArgs {
  Program prog;
  const float* bias[8];
  const float* w_row;     // [256] logit row
  const float* b_logit;   // [>=1]
  const float* bias_feat; // MODE_FEAT: bias of the feature head (rows 1.. of the last Linear)
  int n_feat;             // MODE_FEAT: feature width (256)
  int n_out[8];           // valid output columns per layer (217 for the pre-skip layer)
  int skip;               // layer whose input is cat[x, pe]/sqrt2
  int octaves, pe_dim;
  float rescale;
};

#ifndef PSN_CHEAP_PREFETCH
#define PSN_CHEAP_PREFETCH 1  // TMEM-load prefetch in the epilogue of the single-pass program (epi_for_chunks_pf_ld)
#endif
#ifndef PSN_FULL_PREFETCH
#define PSN_FULL_PREFETCH 0   // the same TMEM-load prefetch in the three-pass program, paid for with un-prefetched (L1-resident) bias loads
#endif
constexpr int MODE_OUT = 0, MODE_SHADOW = 1, MODE_DEBUG = 2, MODE_FEAT = 3;  // MODE_DEBUG = MODE_OUT + layer dump / clock64 trace hooks

template <int MODE, bool CHEAP = false>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(NUM_THREADS, 1)
k_tc_occ(TcGeoArgs g, PointGen gen, long long M_host, const int* M_dev, int out_kind, float* out, float box, int dump_layer,
         float* dump, long long* trace) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Smem s = carve(smem_raw);
  const uint32_t tmem_base = setup(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long M = M_dev ? (long long)*M_dev : M_host;
  const long long n_tiles = (M + TILE_M - 1) / TILE_M;
  const long long iters = pair_iters(n_tiles);  // equal for both CTAs of the pair; surplus tiles are fully masked (idx >= M)

  if (warp < EPI_WARP0) {
    regs_shrink_control();
    if (warp == 0 && lane == 0) producer_loop<CHEAP>(s, g.prog, iters);  // single-pass steps exist only in the CHEAP program
    if (warp == 1) mma_loop<CHEAP>(s, g.prog, iters, tmem_base, MODE == MODE_DEBUG ? trace : nullptr);
    __syncwarp();
  } else {
    regs_grow_epilogue();
    EpiCtx e = epi_ctx(tmem_base);
    const int row = e.row, sub = e.sub;
    for (long long it = 0; it < iters; ++it) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const long long idx = tile * TILE_M + row;
      float p[3] = {0.f, 0.f, 0.f}, vdummy[3];
      if (idx < M) gen_point(gen, idx, p, vdummy);
      {
        const float x[3] = {p[0] / g.rescale, p[1] / g.rescale, p[2] / g.rescale};
        epi_write_pe(s, row, sub, x, g.octaves);
      }
      named_bar_sync(1, EPI_THREADS);  // encoding table complete (also: every warp is done with the previous tile's staging area)
      {  // layer-0 operand: the point encoding, zero padded to one 64-wide K block (16 columns per sub)
        float v[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const int k = sub * CW + i;
          v[i] = k < g.pe_dim ? s.pe[k * TILE_M + row] : 0.f;
        }
        epi_store_a16(e, e.a_col0(), sub * CW, v, CHEAP);
        epi_signal_a(s, 0);
      }
      float part = 0.f;  // partial logit over this thread's 64 columns
#pragma unroll 1
      for (int l = 0; l < 8; ++l) {
        // timeline hook: lane 0 of the epilogue warps of ONE TMEM lane quadrant (dump_layer doubles as the quadrant selector there)
        const bool tr = (MODE == MODE_DEBUG) && trace && it == TRACE_ITER && blockIdx.x == 0 && row == 32 * (dump_layer < 0 ? 0 : dump_layer & 3);
        const float* bias = g.bias[l];
        const bool pre_skip = (l + 1 == g.skip);
        const int n_out = g.n_out[l];
        auto pre_bias = [&](int col, Bias16& b) {
          if (!(PSN_FULL_PREFETCH && !CHEAP)) load_bias16(bias, col, b);
        };
        auto chunk = [&](int pass, int col, float (&v)[CW], const Bias16& b) {
          if (tr && pass == 0) trace[64 + sub * 40 + l * 5] = clock64();
          if (PSN_FULL_PREFETCH && !CHEAP) {  // bias straight from L1 at its use: the registers of the two prefetched copies hold the TMEM chunk of the next pass instead
            Bias16 bl;
            load_bias16(bias, col, bl);
            add16(v, bl.b);
          } else {
            add16(v, b.b);
          }
          if (CHEAP && l < 7 && !(pre_skip && col + CW > n_out)) {
            // level-1 march program, ordinary chunk: activation in packed fp16, result = the hi-only operand columns
            uint32_t r8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) r8[u] = softplus_scaled_cheap_h2(v[2 * u], v[2 * u + 1]);
            tmem_st8(e.tmem_base + e.lane_addr + e.d_col0() + (uint32_t)col, r8);
            epi_signal_a(s, pass);
            return;
          }
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = CHEAP ? softplus_scaled_cheap(v[i]) : softplus_scaled(v[i]);  // s = P h (scaled domain, tc_pack.cu)
          if (MODE == MODE_DEBUG) {
            if (dump && dump_layer == l && idx < M) {  // bring-up hook: the MMA-produced columns only
#pragma unroll
              for (int i = 0; i < CW; ++i) dump[idx * 256 + col + i] = v[i] * (pre_skip ? PSN_SOFTPLUS_C * 0.70710678118654752440f : PSN_SOFTPLUS_C);  // h (/ sqrt2)
            }
          }
          if (MODE == MODE_FEAT && l == 7) {  // h_7 feeds both the fp32 logit row and the feature-head step
            const float4* w4 = reinterpret_cast<const float4*>(g.w_row + col);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float4 w = __ldg(w4 + t);
              part = fmaf(v[4 * t], w.x, part); part = fmaf(v[4 * t + 1], w.y, part);
              part = fmaf(v[4 * t + 2], w.z, part); part = fmaf(v[4 * t + 3], w.w, part);
            }
            epi_store_a16(e, e.d_col0(), col, v, false);
            epi_signal_a(s, pass);
          } else if (l < 7) {
            if (pre_skip && col + CW > n_out) {  // columns n_out.. of the skip layer's input are pe/sqrt2 (network.py:90-91)
#pragma unroll
              for (int i = 0; i < CW; ++i) {
                const int k = col + i - n_out;
                if (k >= 0) v[i] = k < g.pe_dim ? s.pe[k * TILE_M + row] : 0.f;  // the 1 / sqrt2 (and P) sit in the packed skip-layer columns
              }
            }
            epi_store_a16(e, e.d_col0(), col, v, CHEAP);
            epi_signal_a(s, pass);
            if (tr) trace[64 + sub * 40 + l * 5 + 1 + pass] = clock64();
          } else {
            const float4* w4 = reinterpret_cast<const float4*>(g.w_row + col);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float4 w = __ldg(w4 + t);
              part = fmaf(v[4 * t], w.x, part); part = fmaf(v[4 * t + 1], w.y, part);
              part = fmaf(v[4 * t + 2], w.z, part); part = fmaf(v[4 * t + 3], w.w, part);
            }
          }
        };
        if ((CHEAP && PSN_CHEAP_PREFETCH) || (!CHEAP && PSN_FULL_PREFETCH)) epi_for_chunks_pf_ld<Bias16>(s, e, pre_bias, chunk);
        else epi_for_chunks_pf<Bias16>(s, e, pre_bias, chunk);
        e.step_ctr++;
      }
      if (MODE == MODE_FEAT) {  // step 8: feature head, no activation (network.py:95 returns the raw last Linear)
        const int stride = 1 + g.n_feat;
        epi_for_chunks_pf<Bias16>(s, e, [&](int col, Bias16& b) { load_bias16(g.bias_feat, col, b); },
                                  [&](int pass, int col, float (&v)[CW], const Bias16& b) {
          fma16(v, PSN_SOFTPLUS_C, b.b);  // accumulator = W_feat s_7 = P W_feat h_7
          if (idx < M) {
#pragma unroll
            for (int i = 0; i < CW; ++i)
              if (col + i < g.n_feat) out[idx * stride + 1 + col + i] = v[i];
          }
        });
        e.step_ctr++;
      }
      tc_fence_before();  // order this tile's last TMEM reads before the next tile's operand stores / MMAs
      s.stage[sub * TILE_M + row].x = part;
      named_bar_sync(1, EPI_THREADS);
      if (sub == 0) {
        const float z = ((s.stage[row].x + s.stage[TILE_M + row].x) + (s.stage[2 * TILE_M + row].x + s.stage[3 * TILE_M + row].x)) +
                        __ldg(g.b_logit);
        if (MODE == MODE_FEAT) {
          if (idx < M) out[idx * (1 + g.n_feat)] = z;
        } else if (MODE != MODE_SHADOW) {
          if (idx < M) {
            float o = z;
            if (out_kind == PSN_OUT_ALPHA) o = 1.f / (1.f + __expf(10.f * z));
            else if (out_kind == PSN_OUT_NEG_LOGIT) o = -z;
            out[idx] = o;
          }
        } else {
          // shadow ray: rows are the 128 march steps of pair `tile`; alpha zeroed outside the box (rendering.py:402-408)
          float a = 1.f / (1.f + __expf(10.f * z));
          const bool inside = (p[0] <= box) && (p[0] >= -box) && (p[1] <= box) && (p[1] >= -box) && (p[2] <= box) && (p[2] >= -box);
          if (!inside || idx >= M) a = 0.f;
          const float t = (1.f - a) + 1e-6f;
          float incl = t;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const float u = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl *= u;
          }
          float excl = __shfl_up_sync(0xffffffffu, incl, 1);
          if (lane == 0) excl = 1.f;
          const int q = row >> 5;
          if (lane == 31) s.c->g3[q] = incl;
          named_bar_sync(2, 128);
          float pre = 1.f;
          for (int qq = 0; qq < q; ++qq) pre *= s.c->g3[qq];
          float w = a * (pre * excl);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(0xffffffffu, w, off);
          if (lane == 0) s.c->g3[8 + q] = w;
          named_bar_sync(2, 128);
          if (row == 0 && tile < n_tiles) out[tile] = 1.f - (s.c->g3[8] + s.c->g3[9] + s.c->g3[10] + s.c->g3[11]);
        }
      }
      // no trailing barrier: the staging area, the encoding table and g3 are next written after the next tile's encoding
      // barrier / a later sub-0 barrier, which every reader of this tile has to reach first
    }
  }
  teardown(tmem_base);
}

// ---- host -------------------------------------------------------------------------------------------------------------
namespace tc {
int tc_grid(const void* kernel, long long tiles) {
  static int max_ctas = 0;  // one device per process
  if (!max_ctas) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(num_ctas() / CLUSTER * CLUSTER));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute at;
    memset(&at, 0, sizeof(at));
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = CLUSTER; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) == cudaSuccess && n > 0) {
      max_ctas = CLUSTER * n < num_ctas() / CLUSTER * CLUSTER ? CLUSTER * n : num_ctas() / CLUSTER * CLUSTER;
    } else {
      (void)cudaGetLastError();
      max_ctas = num_ctas() / CLUSTER * CLUSTER;
    }
  }
  long long g = tiles < max_ctas ? tiles : max_ctas;
  g = (g + CLUSTER - 1) / CLUSTER * CLUSTER;
  return (int)(g < CLUSTER ? CLUSTER : g);
}
}  // namespace tc

static int make_tc_geo(const psn_mlp* geo, TcGeoArgs* a) {
  PSN_REQUIRE(geo && geo->kind == PSN_NET_GEO, PSN_ERR_ARG, "expected a PSN_NET_GEO handle");
  PSN_REQUIRE(geo->tc_ok, PSN_ERR_SHAPE,
              "geo net shape is not supported by the tensor-core path (needs 8 hidden layers of width 129..256, pe <= 64, feat 256)");
  memset(a, 0, sizeof(*a));
  a->prog.n_steps = 8;
  for (int l = 0; l < 8; ++l) {
    a->prog.step[l].w_off = geo->tc_step[TCG_FWD0 + l].w_off;
    a->prog.step[l].nkb = geo->tc_step[TCG_FWD0 + l].nkb;
    a->prog.step[l].n_pad = geo->tc_step[TCG_FWD0 + l].n_pad;
    a->prog.blob[l] = geo->tc_blob;
    a->bias[l] = geo->tc_bias_scaled[l];
    a->n_out[l] = geo->fwd[l].N;
  }
  a->w_row = geo->tc_w_logit_row_scaled;  // the logit dot runs over s_7 = P h_7
  a->b_logit = geo->logit_head.bias;
  a->skip = geo->desc.skip;
  a->octaves = geo->desc.octaves;
  a->pe_dim = 3 + 6 * geo->desc.octaves;
  a->rescale = geo->desc.rescale;
  if (a->skip >= 0)
    PSN_REQUIRE(a->skip >= 1 && a->skip <= 7 && geo->fwd[a->skip - 1].N + a->pe_dim == 256, PSN_ERR_SHAPE,
                "tensor path: skip layer input must be exactly 256 wide");
  return PSN_OK;
}

template <int MODE, bool CHEAP = false>
static int launch_tc_occ(const TcGeoArgs& a, const PointGen& gen, long long M, const int* M_dev, int out_kind, float* out, float box,
                         int dump_layer, float* dump, cudaStream_t st, long long* trace = nullptr) {
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_occ<MODE, CHEAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  {
    const int rcr = check_launch_regs((const void*)k_tc_occ<MODE, CHEAP>, "k_tc_occ");
    if (rcr) return rcr;
  }
  const long long tiles = M_dev ? (long long)num_ctas() : (M + TILE_M - 1) / TILE_M;
  const int grid = tc_grid((const void*)k_tc_occ<MODE, CHEAP>, tiles);
  count_launch();
  k_tc_occ<MODE, CHEAP><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(a, gen, M, M_dev, out_kind, out, box, dump_layer, dump, trace);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

int tc_occupancy(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, int out_kind, float* out, cudaStream_t st) {
  TcGeoArgs a;
  int rc = make_tc_geo(geo, &a);
  if (rc) return rc;
  if (M == 0 && !M_dev) return PSN_OK;
  return launch_tc_occ<MODE_OUT>(a, gen, M, M_dev, out_kind, out, 0.f, -1, nullptr, st);
}

// NeuralNetwork.infer_occ: [M, 1 + feat] = logit and feature vector (network.py:85-95), the stack plus the feature-head step
int tc_infer_occ(const psn_mlp* geo, const PointGen& gen, long long M, float* out, cudaStream_t st) {
  TcGeoArgs a;
  int rc = make_tc_geo(geo, &a);
  if (rc) return rc;
  if (M == 0) return PSN_OK;
  a.prog.n_steps = 9;
  a.prog.step[8].w_off = geo->tc_step[TCG_FEAT].w_off;
  a.prog.step[8].nkb = geo->tc_step[TCG_FEAT].nkb;
  a.prog.step[8].n_pad = geo->tc_step[TCG_FEAT].n_pad;
  a.prog.blob[8] = geo->tc_blob;
  a.bias_feat = geo->fwd[8].bias;
  a.n_feat = geo->fwd[8].N;
  return launch_tc_occ<MODE_FEAT>(a, gen, M, nullptr, PSN_OUT_LOGIT, out, 0.f, -1, nullptr, st);
}

// Single-pass evaluation of the same stack (first level of the two-level march): NOT within the parity tolerance on its own.
int tc_occupancy_cheap(const psn_mlp* geo, const PointGen& gen, long long M, int out_kind, float* out, cudaStream_t st) {
  TcGeoArgs a;
  int rc = make_tc_geo(geo, &a);
  if (rc) return rc;
  if (M == 0) return PSN_OK;
  for (int l = 0; l < 8; ++l) a.prog.step[l].single = 1;
  return launch_tc_occ<MODE_OUT, true>(a, gen, M, nullptr, out_kind, out, 0.f, -1, nullptr, st);
}

// pairs shadow rays of gen.S == 128 steps each; vis[pair]
int tc_shadow(const psn_mlp* geo, const PointGen& gen, long long pairs, float box, float* vis, cudaStream_t st) {
  TcGeoArgs a;
  int rc = make_tc_geo(geo, &a);
  if (rc) return rc;
  PSN_REQUIRE(gen.S == TILE_M, PSN_ERR_SHAPE, "fused tensor-core shadow pass needs n_steps == 128 (got %d)", gen.S);
  if (pairs == 0) return PSN_OK;
  return launch_tc_occ<MODE_SHADOW>(a, gen, pairs * TILE_M, nullptr, PSN_OUT_ALPHA, vis, box, -1, nullptr, st);
}

}  // namespace psn

using namespace psn;

// sm_100a builds always carry the tcgen05 kernels (engine.tc_available()).
extern "C" int psn_has_tensor_path(void) { return 1; }

// Bring-up / test hook: run the tensor-core geo stack on explicit points and dump the activations that layer `layer`
// hands to the next one (fp32, before the fp16 hi/lo split): out[M, 256].
extern "C" int psn_tc_debug_layer(const psn_mlp* geo, const float* pts, int64_t M, int layer, float* out, float* logits,
                                  void* stream) {
  PSN_REQUIRE(geo && pts && out && logits && layer >= 0 && layer < 8, PSN_ERR_ARG, "psn_tc_debug_layer: bad argument");
  TcGeoArgs a;
  int rc = make_tc_geo(geo, &a);
  if (rc) return rc;
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_EXPLICIT;
  gen.pts = pts;
  return launch_tc_occ<MODE_DEBUG>(a, gen, M, nullptr, PSN_OUT_LOGIT, logits, 0.f, layer, out, (cudaStream_t)stream);
}

// Bring-up tool: clock64() timeline of one tile of the occupancy kernel (see TRACE_ITER in tc_mlp.cuh): trace[256] int64 device.
extern "C" int psn_tc_debug_trace(const psn_mlp* geo, const float* pts, int64_t M, float* out, long long* trace, void* stream) {
  return psn_tc_debug_trace_q(geo, pts, M, out, trace, 0, stream);
}
extern "C" int psn_tc_debug_trace_q(const psn_mlp* geo, const float* pts, int64_t M, float* out, long long* trace, int quadrant,
                                    void* stream) {
  PSN_REQUIRE(geo && pts && out && trace && quadrant >= 0 && quadrant < 4, PSN_ERR_ARG, "psn_tc_debug_trace: bad argument");
  TcGeoArgs a;
  int rc = make_tc_geo(geo, &a);
  if (rc) return rc;
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_EXPLICIT;
  gen.pts = pts;
  return launch_tc_occ<MODE_DEBUG>(a, gen, M, nullptr, PSN_OUT_ALPHA, out, 0.f, quadrant, nullptr, (cudaStream_t)stream, trace);
}
