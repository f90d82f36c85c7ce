// tcgen05 radiance-sample kernel (stage1/model/network.py:122-136): per 128-sample tile, 23 layer-pipelined steps
//   s0..s7   geo layers 0..7 (softplus; sigma'(z) = sigmoid(100 z) of every unit stashed to an L2-resident scratch)
//   s8       feature head (rows 1..256 of the last geo layer, no activation)
//   s9       appearance layer 0, feature part (columns 33.. of lina0): partial pre-activation parked in the scratch
//   s10..s16 reverse pass through layers 7..1:  dx = dz W_l ;  dz_{l-1} = dx * sigma'(z_{l-1})   (analytic normal)
//   s17      reverse layer 0 -> d logit / d pe -> J_pe^T -> gradient; assembles [p, PE(view), gradient] for s18
//   s18      appearance layer 0, remaining 33 inputs (+ parked partial + bias, ReLU)
//   s19..s21 appearance layers 1..3 (ReLU);   s22 appearance layer 4 (N = 3) -> tanh * 0.5 + 0.5
// plus the fp32 logit head (alpha) folded into s7's epilogue.  The same kernel with only s0..s7, s10..s17 is the
// gradient (surface normal) kernel.  Algorithmic work: 2,509,824 FLOP per sample.
//
// Mixed program (PSN_PREC_TC_MIXED, radiance only): s0..s7 stay three-pass - alpha and the sigma' stash need the split product
// (a 2- or 1-pass geo forward misses the 1e-4 gate on alpha: tests/precision_study.py) - but everything the appearance MLP
// consumes tolerates plain fp16 operands: s8..s22 run ONE pass A_hi W_hi (Step::single), their producers write only the hi
// half of the A operand.  rgb moves by 6e-6 rel-L2 per sample / 1e-6 per rendered pixel (same tool); the gradient OUTPUT
// (surface normals) is never taken from this program - tc_gradient always runs the three-pass one.
// (An "H16" variant - sigma' rebuilt in the reverse layers from the fp16 activations instead of a unorm16 stash written by the forward
// layers - ran on the B200 in round 2: bit-compatible gates, 173.4 ms against 173.3 ms.  Not faster, removed; profiles/README.md.)
#include <stdlib.h>

#include "tc_mlp.cuh"
#include "stage1_simt.cuh"
#include "launch.cuh"
#include "internal.cuh"

namespace psn {
using namespace tc;

// Per-CTA scratch (global, sized to stay L2-resident: 148 x 640 KB = 92.5 MB of the 126 MB L2):
//   stash  : sigma'(z) of the 8 x 256 hidden units of the tile's rows as unorm16, uint4 [layer][col/8][row]   (512 KB)
//   parked : fp32 partial pre-activation of appearance layer 0, float4 [col/4][row]                          (128 KB)
// unorm16 keeps the analytic normal within 4e-6 rel-L2 / 2.5e-5 per-row of the fp32-stash result (fp16 would cost 8e-5 /
// 6e-4); the first version stashed fp32 (166 MB over all CTAs) and ncu showed every byte of it going to DRAM and back.
constexpr int SCR_STASH_U4 = 8 * 32 * TILE_M;
constexpr int SCR_PART_F4 = 64 * TILE_M;
constexpr int SCR_U4_PER_CTA = SCR_STASH_U4 + SCR_PART_F4;

// sigma' in [0, 1] -> unorm16 via the 2^23 magic-number add (FFMA + PRMT: no F2I / I2F on the XU pipe next to the MUFUs)
__device__ __forceinline__ uint32_t q16_pair(float a, float b) {
  const uint32_t ua = __float_as_uint(fmaf(a, 65535.f, 8388608.f)), ub = __float_as_uint(fmaf(b, 65535.f, 8388608.f));
  return __byte_perm(ua, ub, 0x5410);
}
__device__ __forceinline__ void dq16_pair(uint32_t w, float& a, float& b) {  // exact integers 0..65535 as floats
  a = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - 8388608.f;
  b = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.f;
}
#define PSN_INV_U16 1.52590218966964e-05f  /* 1 / 65535 */

struct TcRadArgs {
  Program prog;
  const float* gbias[8];
  const float* bias_feat;
  const float* w_row;     // row 0 of the last geo layer (reverse seed dz_7 = row * sigma')
  const float* w_row_s;   // the same row / P for the fp32 logit dot over the scaled activations s_7 = P h_7
  const float* b_logit;
  int n_out[8];
  int skip, octaves, pe_dim;
  float rescale;
  const float* abias[5];
  int octaves_view, pe_view_dim;
  uint4* scratch;
  int with_app;  // 1: radiance (23 steps), 0: gradient only (16 steps)
  int mixed;     // 1: steps s8.. are single-pass (radiance only)
};

#define PSN_INV_SQRT2 0.70710678118654752440f

// Eight softplus activations (scaled domain, tc_pack.cu) with their derivatives: v <- max(z, lg2(1 + e)), sg <- sigma' = e / (1 + e), e = 2^min(z, 30).
// The forward layers of this kernel are bound by the XU pipe (three MUFUs per activation: ex2, lg2, rcp - clock64: 10 200 cycles per
// layer against 8000 for layers without the derivative), so the reciprocals are batched four at a time: ONE MUFU.RCP of the product
// t1 t2 t3 t4 and nine FMULs give the four 1 / t_i (t <= 2^30 + 1 keeps the product below 2^121; sigma' only has to be good to the
// 7.6e-6 of its unorm16 stash, the three extra roundings cost 2e-7).  Clamping at 30 instead of 40 changes nothing: for z > 30 the
// lg2 term is below z and sigma' rounds to 65535 / 65535 either way.
__device__ __forceinline__ void softplus8_d(float* v, float (&sg)[8]) {
  float e[8], t[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    e[u] = ex2_approx(fminf(v[u], 30.f));
    t[u] = 1.f + e[u];
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float p01 = t[4 * q] * t[4 * q + 1], p23 = t[4 * q + 2] * t[4 * q + 3];
    const float r = rcp_approx(p01 * p23);
    const float r01 = r * p23, r23 = r * p01;  // 1 / (t0 t1), 1 / (t2 t3)
    sg[4 * q] = e[4 * q] * (r01 * t[4 * q + 1]);
    sg[4 * q + 1] = e[4 * q + 1] * (r01 * t[4 * q]);
    sg[4 * q + 2] = e[4 * q + 2] * (r23 * t[4 * q + 3]);
    sg[4 * q + 3] = e[4 * q + 3] * (r23 * t[4 * q + 2]);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], lg2_approx(t[u]));
}

// TRACE (bring-up tool only): clock64 timeline of tile iteration TRACE_ITER of CTA 0 - MMA-lane slots as in mma_loop, plus
// trace[192 + step] / trace[224 + step] = time at which row 0 / sub 0 finished / started the epilogue of that step.
// Scratch accesses with an L2 evict_last policy (HINT, the default; PSNERF_B200_STASH_HINT=0 turns it off): the per-CTA stash / parked area is
// rewritten every tile and dead in between, but ncu shows every byte of it written back to DRAM (162 GB per launch); keeping these
// lines at the bottom of the eviction order should let the next tile overwrite them in L2 instead.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <bool HINT>
__device__ __forceinline__ void scr_st(uint4* p, uint4 v, unsigned long long pol) {
  if (HINT)
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
  else
    *p = v;
}
template <bool HINT>
__device__ __forceinline__ uint4 scr_ld(const uint4* p, unsigned long long pol) {
  if (HINT) {
    uint4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.b32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
  }
  return __ldcg(p);
}

// HINT only: once a warp has consumed its 512 contiguous bytes of stash / parked scratch (32 rows x 16 B), the four 128-byte lines
// are dead until the next tile rewrites them.  discard.global.L2 drops them from L2 WITHOUT a write-back - the point of the exercise:
// ncu shows every byte of the scratch going to DRAM (135-162 GB per launch) although nothing ever reads it from there.
template <bool HINT>
__device__ __forceinline__ void scr_discard(const void* p, int row) {
  if (HINT) {
    __syncwarp();  // every lane's load of this line has completed (each lane has used its value)
    if ((row & 7) == 0) asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
  }
}

template <bool TRACE, bool HINT = false>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(NUM_THREADS, 1)
k_tc_rad(TcRadArgs g, PointGen gen, long long M_host, const int* M_dev, float* __restrict__ rgb, float* __restrict__ alpha,
         float* __restrict__ grad_out, long long* trace) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Smem s = carve(smem_raw);
  const uint32_t tmem_base = setup(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long M = M_dev ? (long long)*M_dev : M_host;
  const long long n_tiles = (M + TILE_M - 1) / TILE_M;
  const long long iters = pair_iters(n_tiles);  // equal for both CTAs of the pair; surplus tiles are fully masked (idx >= M)

  if (warp < EPI_WARP0) {
    regs_shrink_control();
    if (warp == 0 && lane == 0) producer_loop(s, g.prog, iters);
    if (warp == 1) mma_loop(s, g.prog, iters, tmem_base, TRACE ? trace : nullptr);
    __syncwarp();
  } else {
    regs_grow_epilogue();
    EpiCtx e = epi_ctx(tmem_base);
    const int row = e.row, sub = e.sub;
    int tstep = 0;  // trace only
#define PSN_RAD_MARK(slot)                                                                                        \
  do {                                                                                                              \
    if (TRACE && trace && it == TRACE_ITER && blockIdx.x == 0 && row == 0 && sub == 0) trace[(slot) + tstep] = clock64(); \
  } while (0)
    const bool ho = g.mixed != 0;  // operands of the single-pass steps: hi half only
    uint4* stash = g.scratch + (size_t)blockIdx.x * SCR_U4_PER_CTA;
    float4* parked = reinterpret_cast<float4*>(stash + SCR_STASH_U4);
    const unsigned long long pol = HINT ? l2_policy_evict_last() : 0ull;
    for (long long it = 0; it < iters; ++it) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const long long idx = tile * TILE_M + row;
      float p[3] = {0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 1.f};
      if (idx < M) gen_point(gen, idx, p, vd);
      {
        const float x[3] = {p[0] / g.rescale, p[1] / g.rescale, p[2] / g.rescale};
        epi_write_pe(s, row, sub, x, g.octaves);
      }
      named_bar_sync(1, EPI_THREADS);  // encoding table complete
      {
        float v[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const int k = sub * CW + i;
          v[i] = k < g.pe_dim ? s.pe[k * TILE_M + row] : 0.f;
        }
        epi_store_a16(e, e.a_col0(), sub * CW, v);
        epi_signal_a(s, 0);
      }
      // ---- s0..s7: geo forward --------------------------------------------------------------------------------------
      float part = 0.f;
      tstep = 0;
#pragma unroll 1
      for (int l = 0; l < 8; ++l) {
        const float* bias = g.gbias[l];
        const bool pre_skip = (l + 1 == g.skip);
        const int n_out = g.n_out[l];
        const bool ho_l = ho && l == 7;  // h_7 feeds the (single-pass) feature head s8
        epi_for_chunks_pf<Bias16, false>(s, e, [&](int col, Bias16& b) { load_bias16(bias, col, b); },
                                  [&](int pass, int col, float (&v)[CW], const Bias16& b) {
          add16(v, b.b);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            float sg[8];
            softplus8_d(&v[8 * t], sg);
            scr_st<HINT>(&stash[(size_t)(l * 32 + (col >> 3) + t) * TILE_M + row],
                         make_uint4(q16_pair(sg[0], sg[1]), q16_pair(sg[2], sg[3]), q16_pair(sg[4], sg[5]), q16_pair(sg[6], sg[7])), pol);
            if (l == 7 && !g.with_app) {  // gradient only: seed dz_7 = W_last[0,:] * sigma'(z_7) directly
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(g.w_row + col) + 2 * t + h);
                v[8 * t + 4 * h] = w.x * sg[4 * h]; v[8 * t + 4 * h + 1] = w.y * sg[4 * h + 1];
                v[8 * t + 4 * h + 2] = w.z * sg[4 * h + 2]; v[8 * t + 4 * h + 3] = w.w * sg[4 * h + 3];
              }
            }
          }
          if (l == 7 && g.with_app) {
            const float4* w4 = reinterpret_cast<const float4*>(g.w_row_s + col);  // row / P: the dot runs over s_7 = P h_7
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float4 w = __ldg(w4 + t);
              part = fmaf(v[4 * t], w.x, part); part = fmaf(v[4 * t + 1], w.y, part);
              part = fmaf(v[4 * t + 2], w.z, part); part = fmaf(v[4 * t + 3], w.w, part);
            }
          }
          if (pre_skip && col + CW > n_out) {  // columns n_out.. of the skip layer's input are pe/sqrt2 (network.py:90-91)
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              const int k = col + i - n_out;
              if (k >= 0) v[i] = k < g.pe_dim ? s.pe[k * TILE_M + row] : 0.f;  // 1 / sqrt2 and P sit in the packed skip-layer columns
            }
          }
          epi_store_a16(e, e.d_col0(), col, v, ho_l);
          epi_signal_a(s, pass);
        });
        PSN_RAD_MARK(192);
        ++tstep;
        e.step_ctr++;
      }
      float g_acc[3] = {0.f, 0.f, 0.f};  // this thread's share of J_pe^T (d logit / d pe)
      if (g.with_app) {
        // ---- s8: feature head -> A (no activation) ---------------------------------------------------------------------
        epi_for_chunks_pf<Bias16, false>(s, e, [&](int col, Bias16& b) { load_bias16(g.bias_feat, col, b); },
                                  [&](int pass, int col, float (&v)[CW], const Bias16& b) {
          fma16(v, PSN_SOFTPLUS_C, b.b);  // accumulator = W_feat s_7 = P W_feat h_7
          epi_store_a16(e, e.d_col0(), col, v, ho);
          epi_signal_a(s, pass);
        });
        PSN_RAD_MARK(192);
        ++tstep;
        e.step_ctr++;
        // ---- s9: appearance layer 0, feature part -> parked; then the reverse seed dz_7 -> A -------------------------------
        struct Seed { uint4 q[2]; float4 w[4]; };
        epi_for_chunks_pf<Seed, false>(s, e, [&](int col, Seed& o) {
#pragma unroll
          for (int t = 0; t < 2; ++t) o.q[t] = scr_ld<HINT>(&stash[(size_t)(7 * 32 + (col >> 3) + t) * TILE_M + row], pol);
#pragma unroll
          for (int t = 0; t < 4; ++t) o.w[t] = __ldg(reinterpret_cast<const float4*>(g.w_row + col) + t);
        }, [&](int pass, int col, float (&v)[CW], const Seed& o) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            scr_st<HINT>(reinterpret_cast<uint4*>(&parked[(size_t)((col >> 2) + t) * TILE_M + row]),
                         make_uint4(__float_as_uint(v[4 * t]), __float_as_uint(v[4 * t + 1]), __float_as_uint(v[4 * t + 2]),
                                    __float_as_uint(v[4 * t + 3])), pol);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint4 q = o.q[t];
            float sg[8];
            dq16_pair(q.x, sg[0], sg[1]); dq16_pair(q.y, sg[2], sg[3]); dq16_pair(q.z, sg[4], sg[5]); dq16_pair(q.w, sg[6], sg[7]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 w = o.w[2 * t + h];
              v[8 * t + 4 * h] = (w.x * PSN_INV_U16) * sg[4 * h]; v[8 * t + 4 * h + 1] = (w.y * PSN_INV_U16) * sg[4 * h + 1];
              v[8 * t + 4 * h + 2] = (w.z * PSN_INV_U16) * sg[4 * h + 2]; v[8 * t + 4 * h + 3] = (w.w * PSN_INV_U16) * sg[4 * h + 3];
            }
          }
          epi_store_a16(e, e.d_col0(), col, v, ho);
          epi_signal_a(s, pass);
#pragma unroll
          for (int t = 0; t < 2; ++t) scr_discard<HINT>(&stash[(size_t)(7 * 32 + (col >> 3) + t) * TILE_M + row], row);
        });
        PSN_RAD_MARK(192);
        ++tstep;
        e.step_ctr++;
      }
      // ---- s10..s16: reverse through layers 7..1 -----------------------------------------------------------------------------
#pragma unroll 1
      for (int l = 7; l >= 1; --l) {
        const bool is_skip = (l == g.skip);
        const int nprev = g.n_out[l - 1];
        const float scale = is_skip ? PSN_INV_SQRT2 * PSN_INV_U16 : PSN_INV_U16;
        struct Sig { uint4 q[2]; };
        epi_for_chunks_pf<Sig, false>(s, e, [&](int col, Sig& o) {
#pragma unroll
          for (int t = 0; t < 2; ++t) o.q[t] = scr_ld<HINT>(&stash[(size_t)((l - 1) * 32 + (col >> 3) + t) * TILE_M + row], pol);
        }, [&](int pass, int col, float (&v)[CW], const Sig& o) {
          if (is_skip && col + CW > nprev) {  // encoding part of the skip input: contributes J_pe^T directly
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              const int k = col + i - nprev;
              if (k >= 0 && k < g.pe_dim) {
                int cc;
                const float t = pe_jac_tab(s, row, k, &cc) * (v[i] * PSN_INV_SQRT2);
                g_acc[0] += cc == 0 ? t : 0.f; g_acc[1] += cc == 1 ? t : 0.f; g_acc[2] += cc == 2 ? t : 0.f;
              }
            }
          }
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint4 q = o.q[t];
            float sg[8];
            dq16_pair(q.x, sg[0], sg[1]); dq16_pair(q.y, sg[2], sg[3]); dq16_pair(q.z, sg[4], sg[5]); dq16_pair(q.w, sg[6], sg[7]);
#pragma unroll
            for (int u = 0; u < 8; ++u) v[8 * t + u] = (v[8 * t + u] * scale) * sg[u];
          }
          if (col + CW > nprev) {
#pragma unroll
            for (int i = 0; i < CW; ++i)
              if (col + i >= nprev) v[i] = 0.f;
          }
          epi_store_a16(e, e.d_col0(), col, v, ho);
          epi_signal_a(s, pass);
#pragma unroll
          for (int t = 0; t < 2; ++t) scr_discard<HINT>(&stash[(size_t)((l - 1) * 32 + (col >> 3) + t) * TILE_M + row], row);
        });
        PSN_RAD_MARK(192);
        ++tstep;
        e.step_ctr++;
      }
      // ---- s17: reverse layer 0 -> gradient ------------------------------------------------------------------------------
      epi_wait_d(s, e);
      if (sub * CW < g.pe_dim) {  // d logit / d pe: columns 0..pe_dim-1 of this step, 16 per sub
        float v[CW];
        epi_load16(e, sub * CW, v);
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          const int k = sub * CW + i;
          if (k < g.pe_dim) {
            int cc;
            const float t = pe_jac_tab(s, row, k, &cc) * v[i];
            g_acc[0] += cc == 0 ? t : 0.f; g_acc[1] += cc == 1 ? t : 0.f; g_acc[2] += cc == 2 ? t : 0.f;
          }
        }
      }
      PSN_RAD_MARK(192);
      ++tstep;
      e.step_ctr++;
      tc_fence_before();
      s.stage[sub * TILE_M + row] = make_float4(g_acc[0], g_acc[1], g_acc[2], part);
      named_bar_sync(1, EPI_THREADS);
      float gr[3], logit;
      {  // fixed summation order => bit-reproducible; every sub needs the gradient for its share of the next operand
        const float4 a0 = s.stage[row], a1 = s.stage[TILE_M + row], a2 = s.stage[2 * TILE_M + row], a3 = s.stage[3 * TILE_M + row];
        gr[0] = ((a0.x + a1.x) + (a2.x + a3.x)) / g.rescale;
        gr[1] = ((a0.y + a1.y) + (a2.y + a3.y)) / g.rescale;
        gr[2] = ((a0.z + a1.z) + (a2.z + a3.z)) / g.rescale;
        logit = ((a0.w + a1.w) + (a2.w + a3.w)) + __ldg(g.b_logit);
      }
      if (sub == 0 && grad_out && idx < M) { grad_out[idx * 3] = gr[0]; grad_out[idx * 3 + 1] = gr[1]; grad_out[idx * 3 + 2] = gr[2]; }
      if (g.with_app) {
        {  // [p, PE(view/|view|), gradient, 0...] -> K block 0, 16 columns per sub (network.py:98,127-132)
          // The view encoding goes through the smem table (the point encoding in it is dead after the Jacobian reads above, which
          // every warp finished before the staging barrier): three sincosf per thread instead of sixteen sinf / cosf selections -
          // the clock64 timeline showed 16 000 idle MMA cycles per tile at this hand-over.
          const float nv = sqrtf(vd[0] * vd[0] + vd[1] * vd[1] + vd[2] * vd[2]);
          const float vn[3] = {vd[0] / nv, vd[1] / nv, vd[2] / nv};
          epi_write_pe(s, row, sub, vn, g.octaves_view);
          named_bar_sync(1, EPI_THREADS);
          const int o_g = 3 + g.pe_view_dim;
          float v[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const int k = sub * CW + i;
            float val = 0.f;
            if (k < 3) val = (k == 0 ? p[0] : (k == 1 ? p[1] : p[2]));
            else if (k < o_g) val = s.pe[(k - 3) * TILE_M + row];
            else if (k < o_g + 3) val = (k == o_g ? gr[0] : (k == o_g + 1 ? gr[1] : gr[2]));
            v[i] = val;
          }
          epi_store_a16(e, e.a_col0(), sub * CW, v, ho);
          epi_signal_a(s, 0);
        }
        // ---- s18: appearance layer 0 (rest) + parked + bias, ReLU ------------------------------------------------------------
        struct Park { float4 pk[4], b[4]; };
        epi_for_chunks_pf<Park, false>(s, e, [&](int col, Park& o) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            {
              const uint4 u = scr_ld<HINT>(reinterpret_cast<const uint4*>(&parked[(size_t)((col >> 2) + t) * TILE_M + row]), pol);
              o.pk[t] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
            }
            o.b[t] = __ldg(reinterpret_cast<const float4*>(g.abias[0] + col) + t);
          }
        }, [&](int pass, int col, float (&v)[CW], const Park& o) {
          add16(v, o.pk);
          add16(v, o.b);
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
          epi_store_a16(e, e.d_col0(), col, v, ho);
          epi_signal_a(s, pass);
#pragma unroll
          for (int t = 0; t < 4; ++t) scr_discard<HINT>(&parked[(size_t)((col >> 2) + t) * TILE_M + row], row);
        });
        PSN_RAD_MARK(192);
        ++tstep;
        e.step_ctr++;
        // ---- s19..s21: appearance layers 1..3 -------------------------------------------------------------------------------
#pragma unroll 1
        for (int l = 1; l <= 3; ++l) {
          const float* bias = g.abias[l];
          epi_for_chunks_pf<Bias16, false>(s, e, [&](int col, Bias16& b) { load_bias16(bias, col, b); },
                                    [&](int pass, int col, float (&v)[CW], const Bias16& b) {
            add16(v, b.b);
#pragma unroll
            for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
            epi_store_a16(e, e.d_col0(), col, v, ho);
            epi_signal_a(s, pass);
          });
          PSN_RAD_MARK(192);
        ++tstep;
        e.step_ctr++;
        }
        // ---- s22: appearance layer 4 -> rgb ------------------------------------------------------------------------------------
        epi_wait_d(s, e);
        if (sub == 0) {
          float v[CW];
          epi_load16(e, 0, v);
          if (idx < M) {
            rgb[idx * 3 + 0] = tanhf(v[0] + __ldg(g.abias[4] + 0)) * 0.5f + 0.5f;
            rgb[idx * 3 + 1] = tanhf(v[1] + __ldg(g.abias[4] + 1)) * 0.5f + 0.5f;
            rgb[idx * 3 + 2] = tanhf(v[2] + __ldg(g.abias[4] + 2)) * 0.5f + 0.5f;
            alpha[idx] = 1.f / (1.f + __expf(10.f * logit));
          }
        }
        PSN_RAD_MARK(192);
        ++tstep;
        e.step_ctr++;
        tc_fence_before();
      }
    }
  }
  teardown(tmem_base);
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
static void put_step(Program& p, int i, const psn_mlp* net, int idx) {
  p.step[i].w_off = net->tc_step[idx].w_off;
  p.step[i].nkb = net->tc_step[idx].nkb;
  p.step[i].n_pad = net->tc_step[idx].n_pad;
  p.blob[i] = net->tc_blob;
}

static int make_tc_rad(const psn_mlp* geo, const psn_mlp* app, void* scratch, TcRadArgs* a, int mixed = 0) {
  PSN_REQUIRE(geo && geo->kind == PSN_NET_GEO && geo->tc_ok, PSN_ERR_SHAPE, "geo net is not packed for the tensor-core path");
  memset(a, 0, sizeof(*a));
  int n = 0;
  for (int l = 0; l < 8; ++l) {
    put_step(a->prog, n++, geo, TCG_FWD0 + l);
    a->gbias[l] = geo->tc_bias_scaled[l];
    a->n_out[l] = geo->fwd[l].N;
  }
  a->with_app = app ? 1 : 0;
  if (app) {
    PSN_REQUIRE(app->kind == PSN_NET_APP && app->tc_ok, PSN_ERR_SHAPE, "app net is not packed for the tensor-core path");
    put_step(a->prog, n++, geo, TCG_FEAT);
    put_step(a->prog, n++, app, TCA_L0F);
  }
  for (int l = 7; l >= 1; --l) put_step(a->prog, n++, geo, TCG_REV_TOP + (7 - l));
  put_step(a->prog, n++, geo, TCG_REV0);
  if (app) {
    put_step(a->prog, n++, app, TCA_L0R);
    for (int l = 1; l <= 3; ++l) put_step(a->prog, n++, app, TCA_L1 + (l - 1));
    put_step(a->prog, n++, app, TCA_L4);
    for (int l = 0; l < 5; ++l) a->abias[l] = app->fwd[l].bias;
    a->octaves_view = app->desc.octaves;
    a->pe_view_dim = 3 + 6 * app->desc.octaves;
    PSN_REQUIRE(3 + a->pe_view_dim + 3 + 256 == app->in_dims[0], PSN_ERR_SHAPE, "app net input layout mismatch");
    PSN_REQUIRE(a->pe_view_dim <= PE_K, PSN_ERR_SHAPE, "tensor path: view encoding wider than %d", PE_K);
  }
  a->prog.n_steps = n;
  a->mixed = (mixed && app) ? 1 : 0;
  if (a->mixed)
    for (int i = 8; i < n; ++i) a->prog.step[i].single = 1;  // s8.. : feature head, reverse sweep, appearance MLP
  a->bias_feat = geo->fwd[8].bias;
  a->w_row = geo->w_logit_row;
  a->w_row_s = geo->tc_w_logit_row_scaled;
  a->b_logit = geo->logit_head.bias;
  a->skip = geo->desc.skip;
  a->octaves = geo->desc.octaves;
  a->pe_dim = 3 + 6 * geo->desc.octaves;
  a->rescale = geo->desc.rescale;
  a->scratch = (uint4*)scratch;
  PSN_REQUIRE(a->skip >= 2 && a->skip <= 7 && geo->fwd[a->skip - 1].N + a->pe_dim == 256, PSN_ERR_SHAPE,
              "tensor path: the skip layer input must be exactly 256 wide");
  return PSN_OK;
}

size_t tc_stash_bytes() {
  const size_t tc = (size_t)num_ctas() * SCR_U4_PER_CTA * sizeof(uint4);
  const size_t si = simt_stash_bytes();
  return tc > si ? tc : si;
}

static int launch_tc_rad(const TcRadArgs& a, const PointGen& gen, long long M, const int* M_dev, float* rgb, float* alpha,
                         float* grad, cudaStream_t st, long long* trace = nullptr) {
  // L2 handling of the per-CTA scratch (evict_last accesses + discard.global.L2 once a line is dead): on by default since round 2
  // (DRAM traffic of a 512 x 512 x 128 launch 181.5 GB -> 11.3 GB, kernel 164.9 -> 160.2 ms under ncu; profiles/r2_ncu_rad_stash_discard.json);
  // PSNERF_B200_STASH_HINT=0 selects the plain accesses for A/B measurements.  Results are bit-identical either way.
  static const bool env_hint = [] { const char* v = getenv("PSNERF_B200_STASH_HINT"); return !(v && v[0] == '0'); }();
  const bool hint = env_hint && !trace;
  const void* kfn = trace ? (const void*)k_tc_rad<true> : (hint ? (const void*)k_tc_rad<false, true> : (const void*)k_tc_rad<false>);
  PSN_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  {
    const int rcr = check_launch_regs(kfn, "k_tc_rad");
    if (rcr) return rcr;
  }
  const long long tiles = M_dev ? (long long)num_ctas() : (M + TILE_M - 1) / TILE_M;
  const int grid = tc_grid(kfn, tiles);
  count_launch();
  if (trace) k_tc_rad<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(a, gen, M, M_dev, rgb, alpha, grad, trace);
  else if (hint) k_tc_rad<false, true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(a, gen, M, M_dev, rgb, alpha, grad, nullptr);
  else k_tc_rad<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(a, gen, M, M_dev, rgb, alpha, grad, nullptr);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

int tc_radiance(const psn_mlp* geo, const psn_mlp* app, const PointGen& gen, long long M, float* rgb, float* alpha, void* stash,
                int mixed, cudaStream_t st) {
  PSN_REQUIRE(app, PSN_ERR_ARG, "tc_radiance: app net is null");
  TcRadArgs a;
  int rc = make_tc_rad(geo, app, stash, &a, mixed);
  if (rc) return rc;
  if (M == 0) return PSN_OK;
  return launch_tc_rad(a, gen, M, nullptr, rgb, alpha, nullptr, st);
}

int tc_gradient(const psn_mlp* geo, const PointGen& gen, long long M, const int* M_dev, float* grad, void* stash, cudaStream_t st) {
  TcRadArgs a;
  int rc = make_tc_rad(geo, nullptr, stash, &a);
  if (rc) return rc;
  if (M == 0 && !M_dev) return PSN_OK;
  return launch_tc_rad(a, gen, M, M_dev, nullptr, nullptr, grad, st);
}

}  // namespace psn

using namespace psn;

// Bring-up tool: clock64() timeline of one tile of the radiance kernel (explicit points / view directions): trace is int64[256].
extern "C" int psn_tc_debug_trace_rad(const psn_mlp* geo, const psn_mlp* app, const float* pts, const float* views, int64_t M,
                                      float* rgb, float* alpha, void* stash, long long* trace, int mixed, void* stream) {
  PSN_REQUIRE(geo && app && pts && views && rgb && alpha && stash && trace, PSN_ERR_ARG, "psn_tc_debug_trace_rad: bad argument");
  TcRadArgs a;
  int rc = make_tc_rad(geo, app, stash, &a, mixed);
  if (rc) return rc;
  PointGen gen;
  memset(&gen, 0, sizeof(gen));
  gen.kind = GEN_EXPLICIT;
  gen.pts = pts;
  gen.views = views;
  return launch_tc_rad(a, gen, M, nullptr, rgb, alpha, nullptr, (cudaStream_t)stream, trace);
}
