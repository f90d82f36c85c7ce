// tcgen05 GEMM of the train steps (stage1_train.cu, stage2_train.cu): fp32 in, fp32 out, TF32 x 3 split products.
//
// The train steps are chains of dense products over fp32 matrices that live in HBM (saved activations, gradients, weights):
//   FORM 0 (NT): C[m,n]  = sum_k A[m,k] B[n,k]   forward Linear    (A = X [M,K],  B = W [N,K])     network.py:92, renderer.py:28
//   FORM 1 (NN): C[m,n]  = sum_k A[m,k] B[k,n]   input gradient    (A = dZ [M,K], B = W [K,N])
//   FORM 2 (TN): C[m,n] += sum_k A[k,m] B[k,n]   weight gradient   (A = dZ [K,M], B = X [K,N]), K = samples, split over CTAs
// Gradients span many binades (1e-9 ... 1), so the fp16 hi/lo split of the inference kernels (tc_mlp.cuh) is not usable here; TF32
// keeps the fp32 exponent.  Every operand value x is split as x = hi + lo and a product is three kind::tf32 MMAs  hi hi + lo hi + hi lo
// accumulated in fp32 (TMEM): per-product error ~2^-21, i.e. an fp32-class GEMM at a sixth of the fp16 tensor rate.  kind::tf32 reads the
// upper 19 bits of each 32-bit operand element, so hi is simply the raw value (truncated by the hardware) and lo = rn_tf32(x - trunc(x)).
//
// One CTA = one 128-row tile of C x up to 256 columns, persistent over tiles.  Warps 0-3: epilogue (TMEM lane quadrant = warp),
// warp 4: MMA issuer, warps 5-12: loaders.  The loaders read fp32 from global (coalesced along the contiguous dimension of the
// operand), split, and write the hi and lo tiles [rows][32 k] straight into the layout the MMA reads - K-major SWIZZLE_128B for the
// operands whose contiguous dimension is k, MN-major tiles for the others - no pre-pass over HBM, no extra copy of any matrix.
// K block = 32 elements (one 128-byte swizzle row); stage = A_hi, A_lo (16 KB each) + B_hi, B_lo (32 KB each) = 96 KB, two stages.
// The fp32 accumulator is double-buffered in TMEM (2 x 256 columns) so that the epilogue of tile i runs under the MMAs of tile i+1.
// Epilogues: row-per-thread stores for 16-byte-aligned outputs, a transposed (coalescing) one for ragged leading dimensions and the
// split-K reductions; EPI 4-7 fuse the element-wise passes of the stage-1 train step (train_gemm.cuh).
// What was measured and dropped in r2 (profiles/r2_time_gemm_variants.json, r2_ncu_train_gemm_gen*_nt.json): a second register buffer in
// the loaders (spills, slower), and a second-generation kernel that staged raw tiles with cp.async and derived the lo tiles with
// converter warps (correct on the first run, tensor pipe 18 % against 36 %: three more mbarrier hand-offs per K block).
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
constexpr int BARS_BYTES = ((int)sizeof(Bars) + 127) / 128 * 128;
constexpr int SMEM = STAGES * STAGE + BARS_BYTES + EPI_WARPS * 32 * 128 + 1024;  // + the epilogue warps' transpose buffers (epilogue_tile)

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
  // fused element-wise epilogues of the second-generation kernel (epi 4-7, FORM 0 / 1): [M, N] matrices with leading dimension lde
  float* C2; const float* E1; float* E2; long long lde; float scale;
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
// The split the loaders store (r2): kind::tf32 reads the upper 19 bits of every 32-bit operand element, so the raw fp32 value IS
// hi = x truncated to TF32 (no instruction), and lo = x - trunc(x) is exact in fp32; adding half a TF32 ulp to its bit pattern makes
// the hardware truncation of lo a round-to-nearest.  Three integer / fp ops per element instead of five (verified by the second-
// generation kernel, which takes hi straight from the cp.async'ed raw tile: tests/test_gpu_tc_gemm.py).
__device__ __forceinline__ float lo_of(float x) { return __uint_as_float(__float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)) + 0x1000u); }
__device__ __forceinline__ float4 lo_of4(const float4& v) { return make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w)); }
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

// Flattened (tile, K block) sequence of one CTA - every role of a kernel walks the same one.
template <int FORM>
struct BlockSeq {
  const Args* a;
  long long t, total, k0, k1, kb;
  TileIt it;
  __device__ explicit BlockSeq(const Args& a_) : a(&a_), t(blockIdx.x), total(n_tiles_total<FORM>(a_)), k0(0), k1(0), kb(0) {
    if (t < total) { it = tile_at<FORM>(*a, t); k_range<FORM>(*a, it, &k0, &k1); kb = k0; }
  }
  __device__ bool valid() const { return t < total; }
  __device__ bool first_of_tile() const { return kb == k0; }
  __device__ bool last_of_tile() const { return kb + BK >= k1; }
  __device__ void next() {
    kb += BK;
    if (kb >= k1) {
      t += gridDim.x;
      if (t < total) { it = tile_at<FORM>(*a, t); k_range<FORM>(*a, it, &k0, &k1); kb = k0; }
    }
  }
};

// ---- epilogue of one accumulator tile, by one warp (its 32 TMEM lanes = 32 rows of C) ----------------------------------------------------
// A thread that reads TMEM owns ONE ROW, so storing straight from the tcgen05.ld registers makes every STG of a warp touch 32
// different 128-byte lines, 16 bytes each: r2 measurements showed the forward / input-gradient forms bound by exactly that (the time of
// a 524288-row product did not depend on K; ncu: l1tex 60 % busy at 36 % tensor activity).  Each 32-column group therefore goes
// through a 4 KB per-warp transpose buffer (16-byte chunks XOR-swizzled by row & 7: conflict-free both ways) and leaves as full
// 128-byte row segments: lane l stores columns 4 (l & 7) .. +3 of rows 4 i + (l >> 3).  The saved matrices of the fused modes
// (EPI 4-7, train_gemm.cuh) are read / written in the same pattern.
constexpr int EPI_STG_BYTES = 32 * 128;
template <int FORM>
__device__ __forceinline__ void epilogue_tile(const Args& a, uint32_t taddr, unsigned char* stg, long long row0, int col0, int bn, int lane) {
  const bool c_vec = ((reinterpret_cast<uintptr_t>(a.C) | (uintptr_t)(a.ldc * 4)) & 15u) == 0;
  const bool e_vec = ((reinterpret_cast<uintptr_t>(a.C2) | reinterpret_cast<uintptr_t>(a.E1) | reinterpret_cast<uintptr_t>(a.E2) |
                       (uintptr_t)(a.lde * 4)) & 15u) == 0;
  const int j = lane & 7, rsub = lane >> 3;
  const int epi = a.epi;
  for (int c = 0; c < bn; c += 32) {
    {
      float v[16];
      tmem_ld16(taddr + (uint32_t)c, v);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (c + 16 < bn) {
        tmem_ld16(taddr + (uint32_t)c + 16u, v);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(stg + lane * 128 + (((q + 4) ^ (lane & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    __syncwarp();
    const int n = col0 + c + 4 * j;                       // first of this lane's four columns
    const int nv = (c + 4 * j < bn) ? (a.N - n < 4 ? a.N - n : 4) : 0;  // valid ones (<= 0: none)
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (FORM != 2 && epi >= 1 && epi <= 4) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e < nv) bias[e] = __ldg(a.bias + n + e);
    }
    if (nv > 0) {
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + rsub;
        const long long row = row0 + r;
        if (row >= a.M) break;  // rows ascend with i
        const float4 xv = *reinterpret_cast<const float4*>(stg + r * 128 + ((j ^ (r & 7)) << 4));
        float x[4] = {xv.x, xv.y, xv.z, xv.w};
        float* dst = a.C + row * a.ldc + n;
        if (FORM == 2) {
          if (c_vec && nv == 4) atomicAdd(reinterpret_cast<float4*>(dst), xv);  // 16-byte reduction (sm_90+)
          else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < nv) atomicAdd(dst + e, x[e]);
          }
          continue;
        }
        float w[4] = {0.f, 0.f, 0.f, 0.f};        // second output of the fused modes
        const long long eoff = row * a.lde + n;
        float e1[4] = {0.f, 0.f, 0.f, 0.f}, e2[4] = {0.f, 0.f, 0.f, 0.f};
        if (epi >= 5) {
          if (e_vec && nv == 4) {
            const float4 t1 = __ldg(reinterpret_cast<const float4*>(a.E1 + eoff));
            e1[0] = t1.x; e1[1] = t1.y; e1[2] = t1.z; e1[3] = t1.w;
            if (epi != 7) {
              const float4 t2 = *reinterpret_cast<const float4*>(a.E2 + eoff);
              e2[0] = t2.x; e2[1] = t2.y; e2[2] = t2.z; e2[3] = t2.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < nv) { e1[e] = __ldg(a.E1 + eoff + e); if (epi != 7) e2[e] = a.E2[eoff + e]; }
          }
        }
        switch (epi) {
          case 1:
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] += bias[e];
            break;
          case 2:
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e] + bias[e], 0.f);
            break;
          case 3:
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = 1.f / (1.f + expf(-(x[e] + bias[e])));
            break;
          case 4:  // softplus_100 + its derivative (k_s1_softplus)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float z = x[e] + bias[e], tt = 100.f * z;
              w[e] = 1.f / (1.f + expf(-tt));
              x[e] = tt > 20.f ? z : log1pf(expf(tt)) / 100.f;
            }
            break;
          case 5:  // second-order pass (k_s1_second): C = d s ; E2 := d E2 100 s (1 - s)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              w[e] = x[e] * e2[e] * (100.f * e1[e] * (1.f - e1[e]));
              x[e] = x[e] * e1[e];
            }
            break;
          case 6:  // zbar (k_s1_zbar): C = acc scale E1 + E2
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = x[e] * a.scale * e1[e] + e2[e];
            break;
          case 7:  // C = acc ; C2 = acc E1
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = x[e] * e1[e];
            break;
          default: break;
        }
        if (c_vec && nv == 4) *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (e < nv) dst[e] = x[e];
        }
        if (epi == 4 || epi == 5 || epi == 7) {
          float* d2 = (epi == 5 ? a.E2 : a.C2) + eoff;
          if (e_vec && nv == 4) *reinterpret_cast<float4*>(d2) = make_float4(w[0], w[1], w[2], w[3]);
          else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < nv) d2[e] = w[e];
          }
        }
      }
    }
    __syncwarp();  // every lane has read the group before the next one overwrites the buffer
  }
}

// Row-per-thread epilogue: every thread stores its own row straight from the tcgen05.ld registers (16-byte pieces of 32 different
// lines per warp instruction).  Shorter dependent chain than epilogue_tile, which wins whenever the 16-byte accesses are possible
// (aligned C): the kernel picks per call.
template <int FORM>
__device__ __forceinline__ void epilogue_rows(const Args& a, uint32_t taddr, long long row, int col0, int bn) {
  const bool c_vec = ((reinterpret_cast<uintptr_t>(a.C) | (uintptr_t)(a.ldc * 4)) & 15u) == 0;
  const bool e_vec = ((reinterpret_cast<uintptr_t>(a.C2) | reinterpret_cast<uintptr_t>(a.E1) | reinterpret_cast<uintptr_t>(a.E2) |
                       (uintptr_t)(a.lde * 4)) & 15u) == 0;
  const int epi = a.epi;
  for (int c = 0; c < bn; c += 16) {
    float v[16], w[16];
    tmem_ld16(taddr + (uint32_t)c, v);
    if (row >= a.M) continue;
    float* dst = a.C + row * a.ldc + col0 + c;
    const bool full = col0 + c + 15 < a.N;
    if (FORM == 2) {
      if (c_vec && full) {
#pragma unroll
        for (int q = 0; q < 4; ++q) atomicAdd(reinterpret_cast<float4*>(dst) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col0 + c + i < a.N) atomicAdd(dst + i, v[i]);
      }
      continue;
    }
    const long long eoff = row * a.lde + col0 + c;
    if (epi >= 1 && epi <= 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (col0 + c + i < a.N) v[i] += __ldg(a.bias + col0 + c + i);
    }
    if (epi == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (epi == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 1.f / (1.f + expf(-v[i]));
    } else if (epi == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float tt = 100.f * v[i];
        w[i] = 1.f / (1.f + expf(-tt));
        v[i] = tt > 20.f ? v[i] : log1pf(expf(tt)) / 100.f;
      }
    } else if (epi >= 5) {
      float e1[16], e2[16];
      if (e_vec && full) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(a.E1 + eoff) + q);
          e1[4 * q] = t1.x; e1[4 * q + 1] = t1.y; e1[4 * q + 2] = t1.z; e1[4 * q + 3] = t1.w;
          if (epi != 7) {
            const float4 t2 = reinterpret_cast<const float4*>(a.E2 + eoff)[q];
            e2[4 * q] = t2.x; e2[4 * q + 1] = t2.y; e2[4 * q + 2] = t2.z; e2[4 * q + 3] = t2.w;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          e1[i] = e2[i] = 0.f;
          if (col0 + c + i < a.N) { e1[i] = __ldg(a.E1 + eoff + i); if (epi != 7) e2[i] = a.E2[eoff + i]; }
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (epi == 5) { w[i] = v[i] * e2[i] * (100.f * e1[i] * (1.f - e1[i])); v[i] *= e1[i]; }
        else if (epi == 6) v[i] = v[i] * a.scale * e1[i] + e2[i];
        else w[i] = v[i] * e1[i];
      }
    }
    if (c_vec && full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (col0 + c + i < a.N) dst[i] = v[i];
    }
    if (epi == 4 || epi == 5 || epi == 7) {
      float* d2 = (epi == 5 ? a.E2 : a.C2) + eoff;
      if (e_vec && full) {
#pragma unroll
        for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(d2)[q] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col0 + c + i < a.N) d2[i] = w[i];
      }
    }
  }
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
    const float4 h = v;
    const float4 l = lo_of4(v);
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
  const bool pow2 = (quads & (quads - 1)) == 0;  // 128- and 256-row tiles: shifts instead of two integer divisions per 16-byte item
  const int qsh = 31 - __clz(quads);
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int idx = t + LOAD_THREADS * j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < quads * BK) {
      const int q = pow2 ? (idx & (quads - 1)) : idx % quads, k = pow2 ? (idx >> qsh) : idx / quads;
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
  const bool pow2 = (quads & (quads - 1)) == 0;  // 128- and 256-row tiles: shifts instead of two integer divisions per 16-byte item
  const int qsh = 31 - __clz(quads);
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int idx = t + LOAD_THREADS * j;
    if (idx >= quads * BK) break;
    const int q = pow2 ? (idx & (quads - 1)) : idx % quads, k = pow2 ? (idx >> qsh) : idx / quads;
    const float v[4] = {f[j].v.x, f[j].v.y, f[j].v.z, f[j].v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t off = sw_off(4 * q + i, k);
      *reinterpret_cast<float*>(hi + off) = v[i];
      *reinterpret_cast<float*>(lo + off) = lo_of(v[i]);
    }
  }
}

// The same fragments written as an MN-major tile: the four consecutive rows of a float4 stay one 16-byte chunk.
template <int MNL, int NIT>
__device__ __forceinline__ void store_row_contig_mn(const Frag (&f)[NIT], unsigned char* hi, unsigned char* lo, int rows, int t, bool is_b) {
  const int quads = rows >> 2;
  const bool pow2 = (quads & (quads - 1)) == 0;  // 128- and 256-row tiles: shifts instead of two integer divisions per 16-byte item
  const int qsh = 31 - __clz(quads);
#pragma unroll
  for (int j = 0; j < NIT; ++j) {
    const int idx = t + LOAD_THREADS * j;
    if (idx >= quads * BK) break;
    const int q = pow2 ? (idx & (quads - 1)) : idx % quads, k = pow2 ? (idx >> qsh) : idx / quads;
    const float4 v = f[j].v;
    const float4 h = v;
    const float4 l = lo_of4(v);
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
  unsigned char* epi_stg = base + STAGES * STAGE + BARS_BYTES;
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
    const bool c_rows = ((reinterpret_cast<uintptr_t>(a.C) | (uintptr_t)(a.ldc * 4)) & 15u) == 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++it_ctr) {
      const TileIt it = tile_at<FORM>(a, t);
      const uint32_t buf = (uint32_t)(it_ctr & 1);
      mbar_wait(&bars->acc_full[buf], (acc_phase >> buf) & 1u);
      acc_phase ^= (1u << buf);
      tc_fence_after();
      // aligned forward / input-gradient outputs: row-per-thread stores; ragged leading dimensions and the split-K reductions: the
      // transposed epilogue (r2 A/B: 0.25 vs 0.47 ms on an epilogue-bound aligned product, 0.81 vs 0.57 ms with ldc = 217)
      if (FORM != 2 && c_rows)
        epilogue_rows<FORM>(a, tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 256u, it.m_tile * BM + warp * 32 + lane, it.n_chunk * bn, bn);
      else
        epilogue_tile<FORM>(a, tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 256u, epi_stg + warp * EPI_STG_BYTES, it.m_tile * BM + warp * 32,
                            it.n_chunk * bn, bn, lane);
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
    auto fetch = [&](Frag (&fa)[BM / 32], Frag (&fb)[MAX_BN / 32], const BlockSeq<FORM>& q) {
      const long long m0 = q.it.m_tile * BM, n0 = (long long)q.it.n_chunk * bn;
      if (FORM == 0) {
        fetch_k_contig(fa, a.A, a.lda, m0, a.M, BM, q.kb, q.k1, t, a_vec);
        fetch_k_contig(fb, a.B, a.ldb, n0, a.N, bn, q.kb, q.k1, t, b_vec);
      } else if (FORM == 1) {
        fetch_k_contig(fa, a.A, a.lda, m0, a.M, BM, q.kb, q.k1, t, a_vec);
        fetch_row_contig(fb, a.B, a.ldb, n0, a.N, bn, q.kb, q.k1, t, b_vec && (n0 & 3) == 0);
      } else {
        fetch_row_contig(fa, a.A, a.lda, m0, a.M, BM, q.kb, q.k1, t, a_vec && (m0 & 3) == 0);
        fetch_row_contig(fb, a.B, a.ldb, n0, a.N, bn, q.kb, q.k1, t, b_vec && (n0 & 3) == 0);
      }
    };
    auto store = [&](const Frag (&fa)[BM / 32], const Frag (&fb)[MAX_BN / 32]) {
      mbar_wait(&bars->empty[stage], phase ^ 1u);
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
    };
    // One register buffer per thread.  (r2 A/B, profiles/r2_time_gemm_variants.json: a second buffer - loads of block i + 1 issued before
    // block i is split and stored - needs 96 live registers for the fragments alone, spills under the 128-register cap of this
    // 416-thread CTA and was 10-50 % SLOWER in every form; so was staging raw tiles with cp.async plus converter warps.)
    Frag fa0[BM / 32], fb0[MAX_BN / 32];
    for (BlockSeq<FORM> q(a); q.valid(); q.next()) {
      fetch(fa0, fb0, q);
      store(fa0, fb0);
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

// PSNERF_B200_TRAIN_FUSED=1 runs the element-wise passes of the stage-1 train step inside the GEMM epilogues (EPI 4-7); the default is
// decided by measurement.
bool train_gemm_fused() {
  const char* e = getenv("PSNERF_B200_TRAIN_FUSED");
  return train_gemm_use_tc() && !(e && e[0] == '0');
}

int tc_gemm(int form, const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc, const float* bias, long long M,
            int N, long long K, int epi, cudaStream_t st, const GemmFuse* fz) {
  using namespace gemm_tc;
  if (M == 0 || N == 0 || K == 0) return PSN_OK;
  Args a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.B = B; a.C = C; a.bias = bias;
  a.lda = lda; a.ldb = ldb; a.ldc = ldc;
  a.M = M; a.K = K; a.N = N; a.epi = epi;
  PSN_REQUIRE(epi <= 3 || (fz && form != 2), PSN_ERR_ARG, "tc_gemm: fused epilogue %d without its operands", epi);
  if (fz) { a.C2 = fz->C2; a.E1 = fz->E1; a.E2 = fz->E2; a.lde = fz->lde; a.scale = fz->scale; }
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
  count_launch();
  // Row-contiguous operands (B of FORM 1, A and B of FORM 2) as MN-major tiles; PSNERF_B200_GEMM_MN=0 selects K-major
  // tiles filled by transposing 4-byte stores instead (the first verified version; stage-1 train step 136 -> 107 ms with MN-major).
  const char* e_mn = getenv("PSNERF_B200_GEMM_MN");
  const bool mn = !(e_mn && e_mn[0] == '0');
  if (form == 0) k_tc_gemm<0, 0><<<grid, THREADS, SMEM, st>>>(a);
  else if (form == 1) { if (mn) k_tc_gemm<1, 2><<<grid, THREADS, SMEM, st>>>(a); else k_tc_gemm<1, 0><<<grid, THREADS, SMEM, st>>>(a); }
  else { if (mn) k_tc_gemm<2, 2><<<grid, THREADS, SMEM, st>>>(a); else k_tc_gemm<2, 0><<<grid, THREADS, SMEM, st>>>(a); }
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

}  // namespace psn

// Test hooks: one GEMM through the tcgen05 kernel (tests/test_gpu_tc_gemm.py).  form / epi as in train_gemm.cuh; the _fused variant
// takes the operands of the fused epilogues 4-7 (second-generation kernel): C2 / E1 / E2 are [M, N] with leading dimension lde.
extern "C" int psn_tc_gemm_debug(int form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                                 const float* bias, int64_t M, int N, int64_t K, int epi, void* stream) {
  PSN_REQUIRE(form >= 0 && form <= 2 && A && B && C && (epi == 0 || bias) && epi <= 3, PSN_ERR_ARG, "psn_tc_gemm_debug: bad argument");
  return psn::tc_gemm(form, A, lda, B, ldb, C, ldc, bias, M, N, K, epi, (cudaStream_t)stream, nullptr);
}
extern "C" int psn_tc_gemm_debug_fused(int form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                                       const float* bias, int64_t M, int N, int64_t K, int epi, float* C2, const float* E1, float* E2,
                                       int64_t lde, float scale, void* stream) {
  PSN_REQUIRE((form == 0 || form == 1) && A && B && C && epi >= 4 && epi <= 7, PSN_ERR_ARG, "psn_tc_gemm_debug_fused: bad argument");
  PSN_REQUIRE((epi != 4 || (bias && C2)) && (epi != 5 || (E1 && E2)) && (epi != 6 || (E1 && E2)) && (epi != 7 || (E1 && C2)), PSN_ERR_ARG,
              "psn_tc_gemm_debug_fused: missing operand of epilogue %d", epi);
  psn::GemmFuse fz;
  fz.C2 = C2; fz.E1 = E1; fz.E2 = E2; fz.lde = lde; fz.scale = scale;
  return psn::tc_gemm(form, A, lda, B, ldb, C, ldc, bias, M, N, K, epi, (cudaStream_t)stream, &fz);
}
