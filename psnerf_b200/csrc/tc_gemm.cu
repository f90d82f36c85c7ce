// tcgen05 GEMM of the train steps (stage1_train.cu, stage2_train.cu): fp32 in, fp32 out, TF32 x 3 split products.
//
// The train steps are chains of dense products over fp32 matrices that live in HBM (saved activations, gradients, weights):
//   FORM 0 (NT): C[m,n]  = sum_k A[m,k] B[n,k]   forward Linear    (A = X [M,K],  B = W [N,K])     network.py:92, renderer.py:28
//   FORM 1 (NN): C[m,n]  = sum_k A[m,k] B[k,n]   input gradient    (A = dZ [M,K], B = W [K,N])
//   FORM 2 (TN): C[m,n] += sum_k A[k,m] B[k,n]   weight gradient   (A = dZ [K,M], B = X [K,N]), K = samples, split over CTAs
// Gradients span many binades (1e-9 ... 1), so the fp16 hi/lo split of the inference kernels (tc_mlp.cuh) is not usable here; TF32
// keeps the fp32 exponent.  Every operand value x is split as x = hi + lo with hi = tf32(x), lo = tf32(x - hi) and a product is
// three kind::tf32 MMAs  hi hi + lo hi + hi lo  accumulated in fp32 (TMEM): per-product error ~2^-21, i.e. an fp32-class GEMM at a
// sixth of the fp16 tensor rate - still ~9 x the FFMA GEMM it replaces (train_gemm.cuh:k_gemm, 26 TFLOP/s measured).
//
// One CTA = one 128-row tile of C x up to 256 columns, persistent over tiles.  Warps 0-3: epilogue (TMEM lane quadrant = warp),
// warp 4: MMA issuer, warps 5-12: loaders.  The loaders read fp32 from global (coalesced along the contiguous dimension of the
// operand), split, and write the hi and lo tiles [rows][32 k] straight into the K-major SWIZZLE_128B layout the MMA reads (a
// transposing write for the operands whose contiguous dimension is not k) - no pre-pass over HBM, no extra copy of any matrix.
// K block = 32 elements (one 128-byte swizzle row); stage = A_hi, A_lo (16 KB each) + B_hi, B_lo (32 KB each) = 96 KB, two stages.
// The fp32 accumulator is double-buffered in TMEM (2 x 256 columns) so that the epilogue of tile i runs under the MMAs of tile i+1.
#include <stdlib.h>

#include "tc_mlp.cuh"
#include "launch.cuh"

namespace psn {
using namespace tc;

namespace gemm_tc {

constexpr int BM = 128;             // rows of C per tile (UMMA M)
constexpr int BK = 32;              // K elements per block: 128 bytes of tf32
constexpr int MAX_BN = 256;         // columns of C per tile (UMMA N <= 256)
constexpr int A_TILE = BM * 128;    // bytes: [128 rows][128 B]
constexpr int B_TILE = MAX_BN * 128;
constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;  // A_hi, A_lo, B_hi, B_lo
constexpr int STAGES = 2;
constexpr int EPI_WARPS = 4, LOAD_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 1 + LOAD_WARPS) * 32;  // 416
constexpr int LOAD_THREADS = LOAD_WARPS * 32;

struct Bars {
  unsigned long long full[STAGES];    // loaders -> MMA (one arrival per loader warp)
  unsigned long long empty[STAGES];   // MMA -> loaders (tcgen05.commit)
  unsigned long long acc_full[2];     // MMA -> epilogue (tcgen05.commit)
  unsigned long long acc_empty[2];    // epilogue -> MMA (EPI_WARPS arrivals)
  unsigned int tmem_base;
};
constexpr int SMEM = STAGES * STAGE + (int)sizeof(Bars) + 1024;

struct Args {
  const float* A; const float* B; float* C; const float* bias;
  long long lda, ldb, ldc;
  long long M, K;        // FORM 2: M = rows of C (columns of A), K = samples
  int N;
  int epi;               // 0 none, 1 +bias, 2 +bias relu, 3 +bias sigmoid (FORM 0 / 1)
  int bn;                // columns per tile (multiple of 16, <= 256)
  int n_chunks;          // ceil(N / bn)
  long long m_tiles;
  int k_splits;          // FORM 2: CTAs sharing one C tile (atomicAdd epilogue)
  long long k_per_split; // FORM 2: multiple of BK
};

// kind::tf32 instruction descriptor: D = F32 (bit 4), A / B format TF32 = 2 (bits 7-9 / 10-12), a_major / b_major (bits 15 / 16:
// 0 = K-major, 1 = MN-major operand tile), N >> 3 (bits 17-22), M >> 4 (bits 24-28)
__device__ __forceinline__ uint32_t idesc_tf32(uint32_t n, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | (8u << 24);
}
// MN-major operand tiles [32 k][rows] for the operands whose contiguous dimension in memory is the row index (B of FORM 1, A and B of
// FORM 2): a float4 of 4 consecutive rows stays one 16-byte chunk, the loaders never transpose.  For 32-bit operands the MN-major
// canonical layout is SWIZZLE_128B_BASE32B (cute/atom/mma_traits_sm100.hpp): blocks of 32 rows (128 bytes) x 4 k = 512-byte atoms with
// the 32-byte chunk index XOR k % 4; the eight k groups of a block are adjacent (SBO = 512), row blocks follow at LBO = 4096; K step
// ks (8 k) of an MMA starts at + ks * 1024; a_major / b_major = 1 in the instruction descriptor.  (Bring-up, r2: the plain SWIZZLE_128B
// and the unswizzled MN-major layouts make the MMA return zeros for tf32; this one passes tests/test_gpu_tc_gemm.py, 47 cases.)
template <int MNL> struct MnLayout;
template <> struct MnLayout<2> {
  static constexpr uint32_t TYPE = 1;  // UMMA::LayoutType::SWIZZLE_128B_BASE32B
  __device__ static uint32_t lbo(bool) { return 4096; }
  __device__ static uint32_t sbo(bool) { return 512; }
  __device__ static uint32_t kstep(bool) { return 1024; }
  __device__ static uint32_t off(int r, int k, bool) {
    return (uint32_t)(r >> 5) * 4096u + (uint32_t)(k >> 2) * 512u + (uint32_t)(k & 3) * 128u +
           (((((uint32_t)r & 31u) >> 3) ^ ((uint32_t)k & 3u)) << 5) + ((uint32_t)r & 7u) * 4u;
  }
};
// descriptor: start address >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout type << 61
template <int MNL>
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, bool is_b) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(MnLayout<MNL>::lbo(is_b) >> 4) << 16) | ((uint64_t)(MnLayout<MNL>::sbo(is_b) >> 4) << 32) |
         (1ull << 46) | ((uint64_t)MnLayout<MNL>::TYPE << 61);
}

__device__ __forceinline__ void umma_ss_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Round to TF32 (10 explicit mantissa bits, ties away from zero like cvt.rna.tf32.f32) with two integer ops: the conversion
// instruction issues on the quarter-rate XU pipe, and a K block needs 24 576 of them per CTA - as many XU cycles as the twelve MMAs
// of the block take on the tensor pipe.  (Inf / NaN inputs are not rounded correctly; the train steps never produce them.)
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u); }
// byte offset of element (row r, k) of a [rows][32] tf32 tile in the K-major SWIZZLE_128B layout
__device__ __forceinline__ uint32_t sw_off(int r, int k) { return (uint32_t)r * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 3u) * 4u; }

// Tile decomposition shared by the three roles.
struct TileIt {
  long long m_tile; int n_chunk; int split;
};
template <int FORM>
__device__ __forceinline__ long long n_tiles_total(const Args& a) { return a.m_tiles * a.n_chunks * (FORM == 2 ? a.k_splits : 1); }
template <int FORM>
__device__ __forceinline__ TileIt tile_at(const Args& a, long long t) {
  TileIt it;
  if (FORM == 2) { it.split = (int)(t % a.k_splits); t /= a.k_splits; } else it.split = 0;
  it.n_chunk = (int)(t % a.n_chunks);
  it.m_tile = t / a.n_chunks;
  return it;
}
template <int FORM>
__device__ __forceinline__ void k_range(const Args& a, const TileIt& it, long long* k0, long long* k1) {
  if (FORM == 2) {
    *k0 = (long long)it.split * a.k_per_split;
    *k1 = *k0 + a.k_per_split < a.K ? *k0 + a.k_per_split : a.K;
  } else { *k0 = 0; *k1 = a.K; }
}

// ---- loaders: 256 threads fill one stage (A_hi, A_lo, B_hi, B_lo) for K block [kb, kb + 32) ----------------------------------------
// Every thread first ISSUES all of its global loads for both operands (4 float4 of A, up to 8 of B: the latency of one HBM / L2
// round trip per K block instead of twelve), then splits and stores.
struct Frag { float4 v; };
// rows x 32 tile of an operand whose k index is CONTIGUOUS in memory (value(r, k) = P[(row0 + r) * ld + kb + k]):
// thread t owns the 16-byte chunk (t & 7) of rows (t >> 3) + 32 j
template <int NIT>
__device__ __forceinline__ void fetch_k_contig(Frag (&f)[NIT], const float* __restrict__ P, long long ld, long long row0, long long n_rows_valid,
                                               int rows, long long kb, long long k_end, int t, bool vec_ok) {
  const int c = t & 7;
  const long long gk = kb + 4 * c;
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int r = (t >> 3) + 32 * j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long gr = row0 + r;
    if (r < rows && gr < n_rows_valid && gk < k_end) {
      const float* src = P + gr * ld + gk;
      if (vec_ok && gk + 3 < k_end) v = __ldg(reinterpret_cast<const float4*>(src));
      else {
        v.x = __ldg(src);
        if (gk + 1 < k_end) v.y = __ldg(src + 1);
        if (gk + 2 < k_end) v.z = __ldg(src + 2);
        if (gk + 3 < k_end) v.w = __ldg(src + 3);
      }
    }
    f[j].v = v;
  }
}
template <int NIT>
__device__ __forceinline__ void store_k_contig(const Frag (&f)[NIT], unsigned char* hi, unsigned char* lo, int rows, int t) {
  const int c = t & 7;
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int r = (t >> 3) + 32 * j;
    if (r >= rows) break;
    const float4 v = f[j].v;
    const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    const float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
    const uint32_t off = (uint32_t)r * 128u + (((uint32_t)c ^ ((uint32_t)r & 7u)) << 4);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}
// rows x 32 tile of an operand whose ROW index is contiguous in memory (value(r, k) = P[(kb + k) * ld + row0 + r]): transposing
// write.  Item idx = t + 256 j: rows 4 (idx % (rows/4)) .. +3 at k = idx / (rows/4); consecutive threads read consecutive float4 of
// one memory row.
template <int NIT>
__device__ __forceinline__ void fetch_row_contig(Frag (&f)[NIT], const float* __restrict__ P, long long ld, long long row0, long long n_rows_valid,
                                                 int rows, long long kb, long long k_end, int t, bool vec_ok) {
  const int quads = rows >> 2;
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int idx = t + LOAD_THREADS * j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < quads * BK) {
      const int q = idx % quads, k = idx / quads;
      const long long gk = kb + k, gr = row0 + 4 * q;
      if (gk < k_end && gr < n_rows_valid) {
        const float* src = P + gk * ld + gr;
        if (vec_ok && gr + 3 < n_rows_valid) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          v.x = __ldg(src);
          if (gr + 1 < n_rows_valid) v.y = __ldg(src + 1);
          if (gr + 2 < n_rows_valid) v.z = __ldg(src + 2);
          if (gr + 3 < n_rows_valid) v.w = __ldg(src + 3);
        }
      }
    }
    f[j].v = v;
  }
}
template <int NIT>
__device__ __forceinline__ void store_row_contig(const Frag (&f)[NIT], unsigned char* hi, unsigned char* lo, int rows, int t) {
  const int quads = rows >> 2;
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int idx = t + LOAD_THREADS * j;
    if (idx >= quads * BK) break;
    const int q = idx % quads, k = idx / quads;
    const float v[4] = {f[j].v.x, f[j].v.y, f[j].v.z, f[j].v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float h = to_tf32(v[i]);
      const uint32_t off = sw_off(4 * q + i, k);
      *reinterpret_cast<float*>(hi + off) = h;
      *reinterpret_cast<float*>(lo + off) = to_tf32(v[i] - h);
    }
  }
}

// The same fragments written as an MN-major tile: the four consecutive rows of a float4 stay one 16-byte chunk.
template <int MNL, int NIT>
__device__ __forceinline__ void store_row_contig_mn(const Frag (&f)[NIT], unsigned char* hi, unsigned char* lo, int rows, int t, bool is_b) {
  const int quads = rows >> 2;
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int idx = t + LOAD_THREADS * j;
    if (idx >= quads * BK) break;
    const int q = idx % quads, k = idx / quads;
    const float4 v = f[j].v;
    const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    const float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
    const uint32_t off = MnLayout<MNL>::off(4 * q, k, is_b);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

// MNL: 0 = every operand tile K-major (row-contiguous operands are transposed by the loaders), 2 = MN-major tiles (MnLayout<2>)
template <int FORM, int MNL>
__global__ void __launch_bounds__(THREADS, 1) k_tc_gemm(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(base + STAGES * STAGE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->full[s], LOAD_WARPS); mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->acc_full[b], 1); mbar_init(&bars->acc_empty[b], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) tmem_alloc_512(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);
  const long long total = n_tiles_total<FORM>(a);
  const int bn = a.bn;

  if (warp < EPI_WARPS) {
    // ---- epilogue: TMEM lanes 32 warp .. +31 = tile rows; 16 columns at a time ----------------------------------------------------
    uint32_t acc_phase = 0;  // bit b = parity to wait for on acc_full[b]
    long long it_ctr = 0;
    const bool c_vec = ((reinterpret_cast<uintptr_t>(a.C) | (uintptr_t)(a.ldc * 4)) & 15u) == 0;  // 16-byte stores (bn is a multiple of 16)
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it_ctr) {
      const TileIt it = tile_at<FORM>(a, t);
      const uint32_t buf = (uint32_t)(it_ctr & 1);
      mbar_wait(&bars->acc_full[buf], (acc_phase >> buf) & 1u);
      acc_phase ^= (1u << buf);
      tc_fence_after();
      const long long row = it.m_tile * BM + warp * 32 + lane;
      const int col0 = it.n_chunk * bn;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 256u;
      for (int c = 0; c < bn; c += 16) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)c, v);
        if (row < a.M) {
          float* dst = a.C + row * a.ldc + col0 + c;
          if (FORM == 2) {
            if (c_vec && col0 + c + 15 < a.N) {  // 16-byte reductions (sm_90+): a quarter of the atomic traffic of the split-K sum
#pragma unroll
              for (int q = 0; q < 4; ++q)
                atomicAdd(reinterpret_cast<float4*>(dst) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col0 + c + i < a.N) atomicAdd(dst + i, v[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = col0 + c + i;
              float x = v[i];
              if (a.epi >= 1 && n < a.N) x += __ldg(a.bias + n);
              if (a.epi == 2) x = fmaxf(x, 0.f);
              if (a.epi == 3) x = 1.f / (1.f + expf(-x));
              v[i] = x;
            }
            if (c_vec && col0 + c + 15 < a.N) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col0 + c + i < a.N) dst[i] = v[i];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
    }
  } else if (warp == EPI_WARPS) {
    // ---- MMA issuer (whole warp converged, one elected lane issues) -----------------------------------------------------------------
    uint32_t stage = 0, phase = 0, acc_phase = 0;
    long long it_ctr = 0;
    constexpr bool A_MN = MNL != 0 && FORM == 2, B_MN = MNL != 0 && FORM >= 1;  // operands whose contiguous dimension is not k
    constexpr int ML = MNL == 0 ? 2 : MNL;  // (any valid MnLayout for the dead branches)
    const uint32_t idesc = idesc_tf32((uint32_t)bn, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it_ctr) {
      const TileIt it = tile_at<FORM>(a, t);
      long long k0, k1;
      k_range<FORM>(a, it, &k0, &k1);
      const uint32_t buf = (uint32_t)(it_ctr & 1);
      if (it_ctr >= 2) {  // the epilogue has drained this accumulator buffer (its use two tiles ago)
        mbar_wait(&bars->acc_empty[buf], (acc_phase >> buf) & 1u);
        acc_phase ^= (1u << buf);
      }
      tc_fence_after();
      const uint32_t d_addr = tmem_base + buf * 256u;
      uint32_t first = 1;
      for (long long kb = k0; kb < k1; kb += BK) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(base + stage * STAGE);
        const uint32_t sA[2] = {sa, sa + A_TILE}, sB[2] = {sa + 2 * A_TILE, sa + 2 * A_TILE + B_TILE};  // hi, lo
        uint64_t dA[2], dB[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          dA[h] = A_MN ? umma_desc_mn<ML>(sA[h], false) : (UMMA_DESC_HI | (uint64_t)umma_desc_lo(sA[h]));
          dB[h] = B_MN ? umma_desc_mn<ML>(sB[h], true) : (UMMA_DESC_HI | (uint64_t)umma_desc_lo(sB[h]));
        }
        const uint32_t a_step = A_MN ? MnLayout<ML>::kstep(false) : 32u, b_step = B_MN ? MnLayout<ML>::kstep(true) : 32u;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = dA[0] + ((ks * a_step) >> 4), al = dA[1] + ((ks * a_step) >> 4);  // advances the start-address field
            const uint64_t bh = dB[0] + ((ks * b_step) >> 4), bl = dB[1] + ((ks * b_step) >> 4);
            umma_ss_tf32(d_addr, ah, bh, idesc, (first && ks == 0) ? 0u : 1u);
            umma_ss_tf32(d_addr, al, bh, idesc, 1u);
            umma_ss_tf32(d_addr, ah, bl, idesc, 1u);
          }
          umma_commit(&bars->empty[stage]);
        }
        __syncwarp();
        first = 0;
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(&bars->acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ---- loaders ---------------------------------------------------------------------------------------------------------------------
    const int t = threadIdx.x - (EPI_WARPS + 1) * 32;
    uint32_t stage = 0, phase = 0;
    const bool a_vec = ((reinterpret_cast<uintptr_t>(a.A) | (uintptr_t)(a.lda * 4)) & 15u) == 0;
    const bool b_vec = ((reinterpret_cast<uintptr_t>(a.B) | (uintptr_t)(a.ldb * 4)) & 15u) == 0;
    for (long long tt = blockIdx.x; tt < total; tt += gridDim.x) {
      const TileIt it = tile_at<FORM>(a, tt);
      long long k0, k1;
      k_range<FORM>(a, it, &k0, &k1);
      const long long m0 = it.m_tile * BM;
      const long long n0 = (long long)it.n_chunk * bn;
      for (long long kb = k0; kb < k1; kb += BK) {
        Frag fa[BM / 32], fb[MAX_BN / 32];  // 4 + 8 float4 in flight per thread
        if (FORM == 0) {
          fetch_k_contig(fa, a.A, a.lda, m0, a.M, BM, kb, k1, t, a_vec);
          fetch_k_contig(fb, a.B, a.ldb, n0, a.N, bn, kb, k1, t, b_vec);
        } else if (FORM == 1) {
          fetch_k_contig(fa, a.A, a.lda, m0, a.M, BM, kb, k1, t, a_vec);
          fetch_row_contig(fb, a.B, a.ldb, n0, a.N, bn, kb, k1, t, b_vec && (n0 & 3) == 0);
        } else {
          fetch_row_contig(fa, a.A, a.lda, m0, a.M, BM, kb, k1, t, a_vec && (m0 & 3) == 0);
          fetch_row_contig(fb, a.B, a.ldb, n0, a.N, bn, kb, k1, t, b_vec && (n0 & 3) == 0);
        }
        mbar_wait(&bars->empty[stage], phase ^ 1u);  // the loads above are in flight while the MMAs release the stage
        unsigned char* sa = base + stage * STAGE;
        unsigned char *Ah = sa, *Al = sa + A_TILE, *Bh = sa + 2 * A_TILE, *Bl = sa + 2 * A_TILE + B_TILE;
        if (FORM == 0) {
          store_k_contig(fa, Ah, Al, BM, t);
          store_k_contig(fb, Bh, Bl, bn, t);
        } else if (FORM == 1) {
          store_k_contig(fa, Ah, Al, BM, t);
          if (MNL != 0) store_row_contig_mn<MNL == 0 ? 2 : MNL>(fb, Bh, Bl, bn, t, true); else store_row_contig(fb, Bh, Bl, bn, t);
        } else {
          if (MNL != 0) { store_row_contig_mn<MNL == 0 ? 2 : MNL>(fa, Ah, Al, BM, t, false); store_row_contig_mn<MNL == 0 ? 2 : MNL>(fb, Bh, Bl, bn, t, true); }
          else { store_row_contig(fa, Ah, Al, BM, t); store_row_contig(fb, Bh, Bl, bn, t); }
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();              // one arrival per warp: 256 single-thread arrivals on one mbarrier serialise (~1000 cycles per K block)
        if (lane == 0) mbar_arrive(&bars->full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc_512(tmem_base);
}

}  // namespace gemm_tc

// PSNERF_B200_TRAIN_GEMM=ffma keeps the fp32 FFMA GEMM (train_gemm.cuh) for A/B measurements and as the cross-check path of the
// gradient tests (read on every call: a getenv against GEMMs of >= 30 us).
bool train_gemm_use_tc() {
  const char* e = getenv("PSNERF_B200_TRAIN_GEMM");
  return !(e && !strcmp(e, "ffma"));
}

int tc_gemm(int form, const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc, const float* bias, long long M,
            int N, long long K, int epi, cudaStream_t st) {
  using namespace gemm_tc;
  if (M == 0 || N == 0 || K == 0) return PSN_OK;
  Args a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.B = B; a.C = C; a.bias = bias;
  a.lda = lda; a.ldb = ldb; a.ldc = ldc;
  a.M = M; a.K = K; a.N = N; a.epi = epi;
  // columns per tile: the whole N when it fits one MMA, else equal chunks; multiples of 16 (UMMA N at M = 128)
  a.n_chunks = (N + MAX_BN - 1) / MAX_BN;
  a.bn = ((N + a.n_chunks - 1) / a.n_chunks + 15) / 16 * 16;
  a.m_tiles = (M + BM - 1) / BM;
  a.k_splits = 1;
  a.k_per_split = K;
  const int ctas = num_ctas();
  if (form == 2) {
    // K = samples.  Enough splits to fill the machine, each at least 8 K blocks long - and at most 1024 samples: every MMA adds its
    // partial sum into the fp32 accumulator with truncation, so a long chain on one accumulator loses accuracy linearly in its
    // length (3 x 128 instructions per 1024 samples ~ 1e-5); the partial tiles are combined by fp32 atomics instead.
    const long long tiles = a.m_tiles * a.n_chunks;
    long long splits = (ctas + tiles - 1) / tiles;
    const long long max_splits = (K + 8 * BK - 1) / (8 * BK);
    const long long min_splits = (K + 1023) / 1024;
    if (splits > max_splits) splits = max_splits;
    if (splits < min_splits) splits = min_splits;
    if (splits < 1) splits = 1;
    a.k_per_split = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    a.k_splits = (int)((K + a.k_per_split - 1) / a.k_per_split);
  }
  const long long total = a.m_tiles * a.n_chunks * a.k_splits;
  const int grid = (int)(total < ctas ? total : ctas);
  static bool attr_set = false;  // one device per process (num_ctas() makes the same assumption)
  if (!attr_set) {
    PSN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    PSN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    PSN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    PSN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    PSN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  // Row-contiguous operands (B of FORM 1, A and B of FORM 2) as MN-major tiles; PSNERF_B200_GEMM_MN=0 selects K-major tiles filled by
  // transposing 4-byte stores instead (the first verified version: cross-check / A/B; stage-1 train step 136 -> 107 ms with MN-major).
  const char* e_mn = getenv("PSNERF_B200_GEMM_MN");
  const bool mn = !(e_mn && e_mn[0] == '0');
  count_launch();
  if (form == 0) k_tc_gemm<0, 0><<<grid, THREADS, SMEM, st>>>(a);
  else if (form == 1) { if (mn) k_tc_gemm<1, 2><<<grid, THREADS, SMEM, st>>>(a); else k_tc_gemm<1, 0><<<grid, THREADS, SMEM, st>>>(a); }
  else { if (mn) k_tc_gemm<2, 2><<<grid, THREADS, SMEM, st>>>(a); else k_tc_gemm<2, 0><<<grid, THREADS, SMEM, st>>>(a); }
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

}  // namespace psn

// Test hook: one GEMM through the tcgen05 kernel (tests/test_gpu_tc_gemm.py).  form / epi as in train_gemm.cuh.
extern "C" int psn_tc_gemm_debug(int form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                                 const float* bias, int64_t M, int N, int64_t K, int epi, void* stream) {
  PSN_REQUIRE(form >= 0 && form <= 2 && A && B && C && (epi == 0 || bias), PSN_ERR_ARG, "psn_tc_gemm_debug: bad argument");
  return psn::tc_gemm(form, A, lda, B, ldb, C, ldc, bias, M, N, K, epi, (cudaStream_t)stream);
}
