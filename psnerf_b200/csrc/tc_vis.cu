// tcgen05 path of the stage-2 visibility MLP (stage2/model/renderer.py:193: 126 -> 256 x8 -> 1 over L x Ns pairs).
//
// Algebraic restructuring (disclosed in DESIGN.md; FLOPs are still reported canonically): the network input is
// cat[embed(p), embed(l)], so layer 0 and the skip layer (whose input is cat[y, embed(p), embed(l)]) split into a
// per-point and a per-light partial product:
//     z0(l, n) = P0[n] + L0[l],          z5(l, n) = W5[:, :256] y + P5[n] + L5[l]
// P*/L* are tiny exact fp32 tables (fp32 FFMA kernel below, Ns x 63 x 512 MAC); the per-pair work is seven
// 256x256 layers on tensor cores plus an fp32 dot-product head.  Tiles are (128 consecutive points) x (one light);
// a CTA walks a contiguous range of tile ids with the light index fastest, so a point block's P rows are re-read
// from L1/L2, not HBM.
#include "tc_mlp.cuh"
#include "simt_mlp.cuh"
#include "launch.cuh"
#include "internal.cuh"

namespace psn {
using namespace tc;

// ---- fp32 tables --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
k_s2_tables(SimtLayer la, SimtLayer lb, int nf, const float* __restrict__ x, long long n, float* __restrict__ outa,
            float* __restrict__ outb) {
  extern __shared__ __align__(16) float smem[];
  float* E = smem;             // [64][LDX]
  float* WS = E + 64 * LDX;
  float* P = WS + WRING_FLOATS;  // [3][TM]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int used = 3 + 6 * nf;
  for (long long tile = blockIdx.x; tile < (n + TM - 1) / TM; tile += gridDim.x) {
    const long long base = tile * TM;
    if (threadIdx.x < TM) {
      const long long i = base + threadIdx.x;
#pragma unroll
      for (int c = 0; c < 3; ++c) P[c * TM + threadIdx.x] = (i < n) ? x[i * 3 + c] : 0.f;
    }
    __syncthreads();
    {
      const int r = threadIdx.x & (TM - 1), q = threadIdx.x / TM;
      const float v[3] = {P[r], P[TM + r], P[2 * TM + r]};
      if (q == 0) { E[0 * LDX + r] = v[0]; E[1 * LDX + r] = v[1]; E[2 * LDX + r] = v[2]; }
      for (int idx = q; idx < nf * 3; idx += NT / TM) {
        const int i = idx / 3, c = idx - 3 * i;
        float s, co;
        sincosf(v[c] * (float)(1 << i), &s, &co);
        E[(3 + 6 * i + c) * LDX + r] = s;
        E[(6 + 6 * i + c) * LDX + r] = co;
      }
      for (int k = used + q; k < 64; k += NT / TM) E[k * LDX + r] = 0.f;
    }
    __syncthreads();
    for (int which = 0; which < 2; ++which) {
      float acc[8][8];
      dense<8>(which ? lb : la, E, WS, acc);
      float* o = which ? outb : outa;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long row = base + ty * 8 + i;
        if (row < n) {
          *reinterpret_cast<float4*>(o + row * 256 + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          *reinterpret_cast<float4*>(o + row * 256 + 128 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
      }
    }
    __syncthreads();
  }
}

// ---- tensor-core pair kernel ------------------------------------------------------------------------------------------
struct TcVisArgs {
  Program prog;          // 7 steps: layers 1..4, 5 (y part), 6, 7
  const float* bias[7];  // bias of layers 1..4, (unused: folded into P5), 6, 7
  const float* w_last;   // [256]
  const float* b_last;   // [1]
  const float *P0, *P5;  // [Ns][256]
  const float *L0, *L5;  // [L][256]
  long long Ns;
  int L;
};

__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(NUM_THREADS, 1)
k_tc_vis(TcVisArgs g, float* __restrict__ vis) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Smem s = carve(smem_raw);
  const uint32_t tmem_base = setup(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pblocks = (g.Ns + TILE_M - 1) / TILE_M;
  const long long n_tiles = pblocks * g.L;
  const long long chunk = (n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t0 = chunk * blockIdx.x;
  const long long t1 = t0 + chunk;   // every CTA runs `chunk` iterations (the pair shares each weight stage); tiles >= n_tiles are masked
  const long long iters = chunk;
  if (warp < EPI_WARP0) {
    regs_shrink_control();
    if (warp == 0 && lane == 0) producer_loop<false>(s, g.prog, iters);  // no single-pass steps in this program
    if (warp == 1) mma_loop<false>(s, g.prog, iters, tmem_base);
    __syncwarp();
  } else {
    regs_grow_epilogue();
    EpiCtx e = epi_ctx(tmem_base);
    const int row = e.row, sub = e.sub;
    // tile -> (point row, light, validity); dummy tiles (t >= n_tiles) recompute the last real one and write nothing
    auto locate = [&](long long t_, long long& nn_, int& l_, long long& n_, bool& valid_) {
      const long long tt = t_ < n_tiles ? t_ : n_tiles - 1;
      const long long pb = tt / g.L;
      l_ = (int)(tt - pb * g.L);
      n_ = pb * TILE_M + row;
      valid_ = n_ < g.Ns && t_ < n_tiles;
      nn_ = valid_ ? n_ : g.Ns - 1;
    };
    // layer 0 of a tile: relu(P0[n] + L0[l]) for the 16 columns [col, col+16) -> A operand of the first tensor step
    auto layer0_chunk = [&](long long nn_, int l_, uint32_t col0, int pass, int col) {
      float v[CW];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g.P0 + nn_ * 256 + col) + i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(g.L0 + (long long)l_ * 256 + col) + i);
        v[4 * i + 0] = fmaxf(a.x + b.x, 0.f); v[4 * i + 1] = fmaxf(a.y + b.y, 0.f);
        v[4 * i + 2] = fmaxf(a.z + b.z, 0.f); v[4 * i + 3] = fmaxf(a.w + b.w, 0.f);
      }
      epi_store_a16(e, col0, col, v);
      epi_signal_a(s, pass);
    };
    long long nn, n, nn2 = 0, n2 = 0;
    int l, l2 = 0;
    bool valid, valid2 = false;
    if (iters > 0) {
      locate(t0, nn, l, n, valid);
#pragma unroll 1
      for (int pass = 0; pass < 4; ++pass) layer0_chunk(nn, l, e.a_col0(), pass, 64 * pass + CW * sub);
    }
    for (long long t = t0; t < t1; ++t) {
      // The next tile's layer 0 is produced INSIDE the last epilogue of this one: pass p of the last step frees accumulator columns
      // [64p, 64p+64) and the next tile's K block p goes straight into them, so the MMA warp starts the next tile while this one is
      // still being reduced (the epilogue of this kernel is cheap - ReLU - and the hand-over used to leave the tensor pipe idle for
      // four exposed table loads plus a layer-time per tile).
      const bool has_next = t + 1 < t1;
      if (has_next) locate(t + 1, nn2, l2, n2, valid2);
      float part = 0.f;
#pragma unroll 1
      for (int st = 0; st < 7; ++st) {
        const float* bias = g.bias[st];
        struct Add2 { float4 a[4], b[4]; };
        epi_for_chunks_pf<Add2>(s, e, [&](int col, Add2& o) {
          if (st == 4) {  // skip layer: + P5[n] (bias folded) + L5[l]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              o.a[i] = __ldg(reinterpret_cast<const float4*>(g.P5 + nn * 256 + col) + i);
              o.b[i] = __ldg(reinterpret_cast<const float4*>(g.L5 + (long long)l * 256 + col) + i);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              o.a[i] = __ldg(reinterpret_cast<const float4*>(bias + col) + i);
              o.b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }, [&](int pass, int col, float (&v)[CW], const Add2& o) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[4 * i + 0] += o.a[i].x + o.b[i].x; v[4 * i + 1] += o.a[i].y + o.b[i].y;
            v[4 * i + 2] += o.a[i].z + o.b[i].z; v[4 * i + 3] += o.a[i].w + o.b[i].w;
          }
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
          if (st < 6) {
            epi_store_a16(e, e.d_col0(), col, v);
            epi_signal_a(s, pass);
          } else {
            const float4* w4 = reinterpret_cast<const float4*>(g.w_last + col);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float4 w = __ldg(w4 + t);
              part = fmaf(v[4 * t], w.x, part); part = fmaf(v[4 * t + 1], w.y, part);
              part = fmaf(v[4 * t + 2], w.z, part); part = fmaf(v[4 * t + 3], w.w, part);
            }
            if (has_next) layer0_chunk(nn2, l2, e.d_col0(), pass, col);
          }
        });
        e.step_ctr++;
      }
      tc_fence_before();
      named_bar_sync(1, EPI_THREADS);  // every sub-0 warp has read the previous tile's staging area
      s.stage[sub * TILE_M + row].x = part;
      named_bar_sync(1, EPI_THREADS);
      if (sub == 0 && valid)
        vis[(long long)l * g.Ns + n] = ((s.stage[row].x + s.stage[TILE_M + row].x) + (s.stage[2 * TILE_M + row].x + s.stage[3 * TILE_M + row].x)) +
                                       __ldg(g.b_last);
      nn = nn2; l = l2; n = n2; valid = valid2;
    }
  }
  teardown(tmem_base);
}

int s2_visibility_simt(const psn_mlp* vis_net, int nf, const float* pts, long long Ns, const float* lights, int L, float* vis,
                       cudaStream_t st);

size_t tc_vis_workspace_bytes(long long Ns, long long L) { return (size_t)(Ns + L) * 2 * 256 * sizeof(float) + 4096; }

int tc_s2_visibility(const psn_mlp* net, int nf, const float* pts, long long Ns, const float* lights, int L, float* vis, void* ws,
                     size_t ws_bytes, cudaStream_t st) {
  PSN_REQUIRE(net && net->kind == PSN_NET_S2, PSN_ERR_ARG, "visibility_net: expected a PSN_NET_S2 handle");
  if (!net->tc_ok || net->in_dims[0] != 2 * (3 + 6 * nf) || 3 + 6 * nf > 64)
    return s2_visibility_simt(net, nf, pts, Ns, lights, L, vis, st);  // shapes outside the tensor plan: fp32 kernels
  if (Ns == 0 || L == 0) return PSN_OK;
  Workspace w(ws, (long long)ws_bytes);
  float* P0 = w.take<float>((size_t)Ns * 256);
  float* P5 = w.take<float>((size_t)Ns * 256);
  float* L0 = w.take<float>((size_t)L * 256);
  float* L5 = w.take<float>((size_t)L * 256);
  PSN_REQUIRE(w.ok, PSN_ERR_WORKSPACE, "visibility (tensor path): workspace too small (need %zu bytes, have %zu)", w.used, ws_bytes);
  const size_t smem_t = (size_t)(64 * LDX + WRING_FLOATS + 3 * TM) * 4;
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_s2_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
  {
    const long long tiles = (Ns + TM - 1) / TM;
    count_launch();
    k_s2_tables<<<(int)(tiles < num_ctas() ? tiles : num_ctas()), NT, smem_t, st>>>(net->vis_aux[0], net->vis_aux[2], nf, pts, Ns, P0, P5);
    const long long tl = (L + TM - 1) / TM;
    count_launch();
    k_s2_tables<<<(int)(tl < num_ctas() ? tl : num_ctas()), NT, smem_t, st>>>(net->vis_aux[1], net->vis_aux[3], nf, lights, L, L0, L5);
    PSN_CUDA_CHECK(cudaGetLastError());
  }
  TcVisArgs a;
  memset(&a, 0, sizeof(a));
  a.prog.n_steps = 7;
  const int layer_of_step[7] = {1, 2, 3, 4, 5, 6, 7};
  for (int i = 0; i < 7; ++i) {
    a.prog.step[i].w_off = net->tc_step[TCV_L1 + i].w_off;
    a.prog.step[i].nkb = net->tc_step[TCV_L1 + i].nkb;
    a.prog.step[i].n_pad = net->tc_step[TCV_L1 + i].n_pad;
    a.prog.blob[i] = net->tc_blob;
    a.bias[i] = net->fwd[layer_of_step[i]].bias;
  }
  a.w_last = net->w_last_row;
  a.b_last = net->fwd[8].bias;
  a.P0 = P0; a.P5 = P5; a.L0 = L0; a.L5 = L5;
  a.Ns = Ns;
  a.L = L;
  PSN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_vis, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  {
    const int rcr = check_launch_regs((const void*)k_tc_vis, "k_tc_vis");
    if (rcr) return rcr;
  }
  const long long n_tiles = ((Ns + TILE_M - 1) / TILE_M) * L;
  const int grid = tc_grid((const void*)k_tc_vis, n_tiles);
  count_launch();
  k_tc_vis<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(a, vis);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

}  // namespace psn
