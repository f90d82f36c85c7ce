// tcgen05 / TMEM / mbarrier / bulk-copy primitives and the layer-pipelined fused-MLP machinery (sm_100a).
//
// One CTA (one per SM, persistent) owns a 128-row tile.  Per "step" (one Linear layer)
//     D[128 x N] (fp32, TMEM)  =  A[128 x K] (fp16 hi+lo, TMEM)  x  W[N x K]^T (fp16 hi+lo, shared ring)
// is evaluated as three UMMA passes  A_hi W_hi + A_lo W_hi + A_hi W_lo  (error-compensated split operands:
// the dropped A_lo W_lo term is 2^-22 relative, fp32 accumulate) so the result tracks an fp32 GEMM.
//
// TMEM (all 512 columns) = two 256-column regions R0 / R1.  Step sc accumulates D into R[sc & 1]; the epilogue converts that
// accumulator IN PLACE into the next step's A operand: the 16 fp32 columns [16c, 16c+16) of a row become 8 columns of packed
// fp16 hi (K elements 16c..16c+15, two per 32-bit column, even k in the low half) followed by 8 columns of packed fp16 lo, so
// K-step c of step sc+1 reads its A operand straight from TMEM (tcgen05.mma with A in TMEM) at R[sc & 1] + 16c (hi) and
// + 16c + 8 (lo) while it accumulates into the other region.  Activations therefore never touch shared memory: the MMA reads
// only the weight operand from smem (8 KB per instruction instead of 12 KB) and the epilogue's stores go to TMEM
// (tcgen05.st), which is what lifts the shared-memory bandwidth ceiling the first version of these kernels ran into
// (ncu: tensor pipe 54 % active with the smem tensor-read path already at 40 %).
// Shared memory: weight ring = 6 stages x 32 KB, each one pre-swizzled [N rows][64 K] K-major SWIZZLE_128B tile streamed
// from L2 with cp.async.bulk (UBLKCP) + mbarrier complete_tx; a [40][128] fp32 point-encoding table; an 8 KB staging area.
//
// CTA pairs: the kernels launch as clusters of two CTAs (one TPC).  Both CTAs walk the same step program on different tiles, so
// every weight tile is fetched from L2 ONCE per pair: CTA r issues the tiles with (tile index & 1) == r as a multicast bulk
// copy that lands in both CTAs' rings and completes both CTAs' w_full barriers; a stage is refilled when BOTH MMA warps have
// released it (tcgen05.commit multicast onto the w_empty barriers of the pair, 2 arrivals).  This halves the L2 -> SM weight
// stream, which had become the limiter (clock64: ~1500 cycles / layer waiting for weight stages at 4 KB/clk chip-wide).
//
// Warp roles: warp 0 = TMEM allocator + weight producer (one lane), warp 1 = MMA issuer (one lane), warps 2-3 idle (the
// control warpgroup releases registers with setmaxnreg.dec);
// warps 4..19 = epilogue: warp w reads TMEM lanes 32*(w%4).. and is "sub" s = (w-4)/4 of its lane quadrant; in pass p sub s
// owns columns [64p + 16s, +16), so every pass completes one 64-wide K block of the next layer.
// Layer pipelining: the epilogue of step L signals a_ready[kb] after each pass and the MMA warp starts step L+1's K block kb at
// once.  The LAST K block of every step is issued as columns [0, 64) (twelve N = 64 MMAs, committed to d_q[buf][0]) followed by
// the remaining columns (d_q[buf][1]): epilogue pass 0 needs only the first part, so it starts a quarter of a K block after the
// last a_ready instead of a whole one and the tensor pipe works on the other columns underneath pass 0.
#pragma once
#include "common.cuh"

namespace psn {
namespace tc {

constexpr int TILE_M = 128;
constexpr int KBLK = 64;                      // K elements per block (128 bytes of fp16)
constexpr int A_MAX_KB = 4;                   // K <= 256
constexpr int W_STAGE_BYTES = 256 * 128;      // 32 KB: [256 N rows][64 K] fp16
constexpr int W_STAGES = 6;
constexpr int PE_K = 40;                      // rows of the fp32 point-encoding table (39 used at 6 octaves)
constexpr int PE_BYTES = PE_K * TILE_M * 4;   // pe[k][row]
constexpr int STAGE_BYTES = 4 * TILE_M * 16;  // cross-sub reduction staging: float4 [4 subs][128 rows]
constexpr int NUM_THREADS = 640;  // warpgroup 0 = control (setmaxnreg 32), warpgroups 1-4 = epilogue (setmaxnreg 112)
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 512;
constexpr int EPI_SUBS = 4;       // epilogue warps per TMEM lane quadrant
constexpr int CW = 16;            // columns per epilogue chunk: in pass p sub s owns columns [64 p + 16 s, +16)
#ifndef PSN_CLUSTER
#define PSN_CLUSTER 2
#endif
constexpr int CLUSTER = PSN_CLUSTER;   // CTAs per cluster sharing one weight stream (multicast); 2 = one TPC (r2 A/B of 4: see profiles/README.md)
constexpr int A_READY_ARRIVALS = 16;   // a K block (64 columns) is written by the 16 epilogue warps in ONE pass; lane 0 of each arrives

// ---- shared-memory control block (after the 1024-aligned A and W regions) ---------------------------------
struct Ctrl {
  unsigned long long w_full[W_STAGES];
  unsigned long long w_empty[W_STAGES];
  unsigned long long a_ready[A_MAX_KB];
  unsigned long long d_q[2][2];   // accumulator buffer x {columns [0, 64), columns [64, N)}
  unsigned int tmem_base;
  unsigned int pad;
  float g3[64];            // scratch of the shadow-ray transmittance scan
};
constexpr int SMEM_BYTES = W_STAGES * W_STAGE_BYTES + PE_BYTES + STAGE_BYTES + (int)sizeof(Ctrl);  // dynamic smem is declared __align__(1024)
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
#ifndef PSN_MBAR_HINT_NS
#define PSN_MBAR_HINT_NS 2000
#endif
__device__ __forceinline__ bool mbar_try_wait(void* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)PSN_MBAR_HINT_NS)  // suspend-time hint (ns): sleep in hardware instead of spinning
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// all threads of both CTAs (also a CTA-wide barrier)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// waits whose arrivals come from the other CTA of the pair (multicast tcgen05.commit)
__device__ __forceinline__ bool mbar_try_wait_cluster(void* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster_relaxed(void* bar, uint32_t parity) {
  while (!mbar_try_wait_cluster(bar, parity)) __nanosleep(32);
}
// bulk copy delivered to the same smem offset (and the same mbarrier offset) of every CTA in cta_mask
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, void* bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_512(uint32_t* smem_dst) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_dst)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// kind::f16, A = B = fp16 (format 0), D = fp32, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | (8u << 24); }

// D[tmem] (+)= A[tmem] * B[smem]^T : A operand = 8 TMEM columns (16 packed fp16 of K per row) at tmem_a
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
// same, arriving on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_multicast(void* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void umma_commit(void* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// issue / wait halves of the 16-column load: global loads placed between them overlap the TMEM read
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 16 consecutive 32-bit columns: thread (lane i) writes row (lane_base + i), columns col..col+15
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// register re-allocation between warpgroups (all 4 warps of a warpgroup must execute it)
// The CTA's register pool is what it was launched with (ptxas: 96 regs x 640 threads); setmaxnreg.inc BLOCKS until the pool
// has room, so the budget must close exactly: 4 control warps x 32 + 16 epilogue warps x 112 = 96 x 20 warps.
constexpr int LAUNCH_REGS = 96, CONTROL_REGS = 32, EPILOGUE_REGS = 112;  // LAUNCH_REGS is verified at launch (check_launch_regs)
static_assert(CONTROL_REGS * 4 + EPILOGUE_REGS * 16 <= LAUNCH_REGS * 20, "setmaxnreg budget exceeds the CTA register pool (deadlock)");
__device__ __forceinline__ void regs_shrink_control() { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory"); }
__device__ __forceinline__ void regs_grow_epilogue() { asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Host-side guard: setmaxnreg.inc deadlocks if the kernel was compiled to fewer launch registers than the budget assumes.
inline int check_launch_regs(const void* kernel, const char* name) {
  cudaFuncAttributes at;
  PSN_CUDA_CHECK(cudaFuncGetAttributes(&at, kernel));
  PSN_REQUIRE(at.numRegs >= 96 && at.numRegs <= 102, PSN_ERR_CUDA,
              "%s was compiled with %d registers/thread; the setmaxnreg budget (tc_mlp.cuh) assumes 96", name, at.numRegs);
  return PSN_OK;
}

// Both CTAs of a pair must run the same number of tile iterations (they share every weight stage): the pair uses the count of
// its even CTA, the odd one may get one masked dummy tile.  Round-robin tile assignment: tile = blockIdx.x + it * gridDim.x.
__device__ __forceinline__ long long pair_iters(long long n_tiles) {
  const long long b0 = (long long)(blockIdx.x - blockIdx.x % CLUSTER);
  return n_tiles > b0 ? (n_tiles - b0 + gridDim.x - 1) / gridDim.x : 0;
}
// persistent grid: an even number of CTAs, at most 2 x the clusters the device can hold at once
int tc_grid(const void* kernel, long long tiles);

// ---- step table -------------------------------------------------------------------------------------------------
struct Step {
  uint32_t w_off;  // byte offset (from the net's tile blob) of tile (kb=0, hi); tiles follow as [kb][hi, lo]
  uint8_t nkb;     // K blocks of 64
  uint8_t single;  // 1: ONE pass A_hi W_hi (plain fp16 operands) - only the hi tiles are streamed and the producing epilogue
                   // writes only the hi half of the A operand (epi_store_a16 hi_only); 0: the three-pass split product
  uint16_t n_pad;  // N (multiple of 16, <= 256); a tile is n_pad x 128 bytes
};
// The same 8 bytes as the loops that never see single-pass steps read them (ALLOW_SINGLE = false: `single` is 0 there, so nkb and
// single together are the 16-bit K-block count).  Keeps the code of those kernels what it was before Step::single existed.
struct StepNoSingle {
  uint32_t w_off;
  uint16_t nkb;
  uint16_t n_pad;
};
static_assert(sizeof(StepNoSingle) == sizeof(Step), "Step views must overlay");
template <bool ALLOW_SINGLE> struct StepView { using type = Step; };
template <> struct StepView<false> { using type = StepNoSingle; };
__device__ __forceinline__ bool step_is_single(const Step& s) { return s.single != 0; }
__device__ __forceinline__ bool step_is_single(const StepNoSingle&) { return false; }
constexpr int MAX_STEPS = 24;
struct Program {
  Step step[MAX_STEPS];
  const unsigned char* blob[MAX_STEPS];  // tile blob base per step (steps may come from different nets)
  int n_steps;
};

// ---- shared-memory carve-up ---------------------------------------------------------------------------------------
struct Smem {
  unsigned char* w;   // weight ring, 1024-aligned
  float* pe;          // [PE_K][TILE_M] point encoding of the current tile (written by the epilogue warps)
  float4* stage;      // [4 subs][TILE_M] cross-sub reduction staging
  Ctrl* c;
};
__device__ __forceinline__ Smem carve(unsigned char* raw) {
  Smem s;
  if (smem_u32(raw) & 1023u) __trap();  // SWIZZLE_128B operands need 1024-byte alignment
  s.w = raw;
  s.pe = reinterpret_cast<float*>(s.w + W_STAGES * W_STAGE_BYTES);
  s.stage = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s.pe) + PE_BYTES);
  s.c = reinterpret_cast<Ctrl*>(reinterpret_cast<unsigned char*>(s.stage) + STAGE_BYTES);
  return s;
}

// Barrier init (thread 0) + TMEM allocation (warp 0).  Ends with a CTA barrier; returns the TMEM base address.
__device__ __forceinline__ uint32_t setup(const Smem& s) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(&s.c->w_full[i], 1); mbar_init(&s.c->w_empty[i], CLUSTER); }
    for (int i = 0; i < A_MAX_KB; ++i) mbar_init(&s.c->a_ready[i], A_READY_ARRIVALS);
    for (int i = 0; i < 2; ++i)
      for (int q = 0; q < 2; ++q) mbar_init(&s.c->d_q[i][q], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_512(&s.c->tmem_base);
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any multicast copy / commit can reach them
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(&s.c->tmem_base);
}
__device__ __forceinline__ void teardown(uint32_t tmem_base) {
  tc_fence_before();
  cluster_sync_all();  // the peer's last multicast commits have landed on this CTA's barriers before its smem goes away
  if ((threadIdx.x >> 5) == 0) tmem_dealloc_512(tmem_base);
}

// ---- producer: stream every weight tile of `iters` tile-iterations through the ring (warp 0, lane 0) ------------------
// Both CTAs of the pair run this loop over the same tile sequence; tile t is fetched (multicast) by CTA (t & 1).
// ALLOW_SINGLE = false compiles the Step::single handling out (kernels whose programs never contain single-pass steps).
template <bool ALLOW_SINGLE = true>
__device__ __forceinline__ void producer_loop(const Smem& s, const Program& prog, long long iters) {
  uint32_t stage = 0, phase = 0;
  const uint32_t rank = cluster_ctarank();
  uint32_t t_parity = 0;  // tile counter modulo CLUSTER: CTA `rank` issues the tiles with counter == rank
  for (long long it = 0; it < iters; ++it) {
    for (int st = 0; st < prog.n_steps; ++st) {
      const typename StepView<ALLOW_SINGLE>::type sp = reinterpret_cast<const typename StepView<ALLOW_SINGLE>::type&>(prog.step[st]);
      const uint32_t tile_bytes = (uint32_t)sp.n_pad * 128u;
      const unsigned char* src = prog.blob[st] + sp.w_off;
      const bool single = ALLOW_SINGLE && step_is_single(sp);
      const int n_tiles = single ? sp.nkb : 2 * sp.nkb;
      for (int t = 0; t < n_tiles; ++t, t_parity = (t_parity + 1u == CLUSTER ? 0u : t_parity + 1u)) {
        mbar_wait_cluster_relaxed(&s.c->w_empty[stage], phase ^ 1u);     // released by the MMA warps of both CTAs
        mbar_arrive_expect_tx(&s.c->w_full[stage], tile_bytes);          // this CTA's copy of the tile
        const int blob_tile = single ? 2 * t : (t ^ 1);  // ring order per K block: lo tile, then hi tile (blob: hi, lo); single: hi only
        if (t_parity == rank)
          bulk_g2s_multicast(s.w + stage * W_STAGE_BYTES, src + (size_t)blob_tile * tile_bytes, tile_bytes, &s.c->w_full[stage],
                             (uint16_t)((1u << CLUSTER) - 1u));
        if (++stage == W_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  }
}

// ---- MMA issuer (warp 1, all lanes converged; one elected lane issues) -----------------------------------------------
// The whole warp runs the loop and waits on the barriers; tcgen05.mma / tcgen05.commit sit under an elect.sync predicate.
// ptxas then emits straight-line UTCHMMA sequences.  (Issuing from an `if (lane == 0)` branch instead wraps EVERY UTCHMMA
// in an ELECT / BRA.U.ANY retry loop plus per-instruction descriptor arithmetic: the clock64 timeline showed ~86 cycles per
// issue, which made the N = 64 MMAs of the split tail issue-bound.)
// trace (optional, bring-up tool): for CTA 0, tile iteration TRACE_ITER the elected lane stores clock64() of every passed
// a_ready wait (slot st*8 + kb), the cycles the step spent waiting for activations / weight stages (slots +4 / +5) and the
// time the step's last commit was issued (slot st*8 + 7).
constexpr int TRACE_ITER = 3;
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Shared-memory descriptor of a K-major SWIZZLE_128B operand: start address >> 4 in bits 0-13, leading-dimension byte offset
// (ignored for swizzled K-major layouts; 1) in bits 16-29, stride between 8-row swizzle atoms 1024 B >> 4 in bits 32-45, descriptor
// version 1 at bit 46, layout type SWIZZLE_128B (2) in bits 61-63.  Only the low word depends on the address, so advancing the
// start by `bytes` (a multiple of 16 that stays inside the 256 KB window) is an add on the low word.
constexpr uint64_t UMMA_DESC_HI = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc_at(uint32_t lo, uint32_t bytes) { return UMMA_DESC_HI | (uint64_t)(lo + (bytes >> 4)); }

template <bool ALLOW_SINGLE = true>
__device__ __forceinline__ void mma_loop(const Smem& s, const Program& prog, long long iters, uint32_t tmem_base,
                                         long long* trace = nullptr) {
  uint32_t stage = 0, phase = 0;   // weight ring
  uint32_t a_phase = 0;            // bit kb = parity to wait for on a_ready[kb]
  uint32_t step_ctr = 0;           // selects the TMEM regions
  const uint32_t w_base = smem_u32(s.w);
  for (long long it = 0; it < iters; ++it) {
    for (int st = 0; st < prog.n_steps; ++st, ++step_ctr) {
      const typename StepView<ALLOW_SINGLE>::type sp = reinterpret_cast<const typename StepView<ALLOW_SINGLE>::type&>(prog.step[st]);
      const uint32_t buf = step_ctr & 1u;
      const uint32_t d_addr = tmem_base + buf * 256u;           // accumulator of this step
      const uint32_t a_addr = tmem_base + (buf ^ 1u) * 256u;    // its A operand = the previous step's accumulator, converted in place
      const uint32_t idesc = umma_idesc(sp.n_pad);
      long long t_wait_a = 0, t_wait_w = 0;  // bring-up trace only (dead code otherwise)
      for (int kb = 0; kb < sp.nkb; ++kb) {
        if (ALLOW_SINGLE && step_is_single(sp)) {
          // single-pass step: one weight stage (the hi tile) and four MMAs A_hi W_hi per K block; same split of the last K block
          const uint32_t st_w = stage, ph_w = phase;
          if (++stage == W_STAGES) { stage = 0; phase ^= 1u; }
          mbar_wait(&s.c->w_full[st_w], ph_w);
          mbar_wait(&s.c->a_ready[kb], (a_phase >> kb) & 1u);
          a_phase ^= (1u << kb);
          tc_fence_after();
          const uint32_t a_s = a_addr + (uint32_t)kb * 64u;
          const uint32_t lo_w = umma_desc_lo(w_base + st_w * W_STAGE_BYTES);
          if (elect_one()) {
            if (trace && it == TRACE_ITER && blockIdx.x == 0) trace[st * 8 + kb] = clock64();
            if (kb + 1 < sp.nkb) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_ts_f16(d_addr, a_s + ks * 16, umma_desc_at(lo_w, ks * 32), idesc, (kb | ks) ? 1u : 0u);
            } else {
              const uint32_t n0 = sp.n_pad < 64 ? (uint32_t)sp.n_pad : 64u;
              const uint32_t id0 = umma_idesc(n0);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma_ts_f16(d_addr, a_s + ks * 16, umma_desc_at(lo_w, ks * 32), id0, (kb | ks) ? 1u : 0u);
              umma_commit(&s.c->d_q[buf][0]);
              if (sp.n_pad > 64) {
                const uint32_t id1 = umma_idesc((uint32_t)sp.n_pad - 64u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_ts_f16(d_addr + 64u, a_s + ks * 16, umma_desc_at(lo_w, 8192 + ks * 32), id1, (kb | ks) ? 1u : 0u);
              }
              umma_commit(&s.c->d_q[buf][1]);
              if (trace && it == TRACE_ITER && blockIdx.x == 0) trace[st * 8 + 7] = clock64();
            }
            umma_commit_multicast(&s.c->w_empty[st_w], (uint16_t)((1u << CLUSTER) - 1u));
          }
          __syncwarp();
          continue;
        }
        // this K block's two weight stages (lo tile first: it is consumed - and released - first).  They are checked BEFORE the
        // activations: the weights are normally long there (the checks cost ~100 cycles of barrier round trip each), whereas the
        // a_ready -> first MMA latency sits on the critical epilogue -> MMA -> epilogue chain of every layer.
        const long long tw0 = trace ? clock64() : 0;
        const uint32_t st_lo = stage, ph_lo = phase;
        if (++stage == W_STAGES) { stage = 0; phase ^= 1u; }
        const uint32_t st_hi = stage, ph_hi = phase;
        if (++stage == W_STAGES) { stage = 0; phase ^= 1u; }
        mbar_wait(&s.c->w_full[st_lo], ph_lo);
        mbar_wait(&s.c->w_full[st_hi], ph_hi);
        const long long tw1 = trace ? clock64() : 0;
        mbar_wait(&s.c->a_ready[kb], (a_phase >> kb) & 1u);
        a_phase ^= (1u << kb);
        tc_fence_after();
        const uint32_t a_kb = a_addr + (uint32_t)kb * 64u;
        if (trace) {
          const long long tw2 = clock64();
          t_wait_w += tw1 - tw0;
          t_wait_a += tw2 - tw1;
          if (it == TRACE_ITER && blockIdx.x == 0 && st >= 1 && st <= 4 && (threadIdx.x & 31) == 0) {
            trace[224 + (st - 1) * 8 + kb] = tw2 - tw1;      // per K block: wait for activations ...
            trace[224 + (st - 1) * 8 + 4 + kb] = tw1 - tw0;  // ... and for its two weight stages (layers 1..4)
          }
        }
        const uint32_t lo_hi = umma_desc_lo(w_base + st_hi * W_STAGE_BYTES), lo_lo = umma_desc_lo(w_base + st_lo * W_STAGE_BYTES);
        if (elect_one()) {
          if (trace && it == TRACE_ITER && blockIdx.x == 0) {
            trace[st * 8 + kb] = clock64();
            trace[st * 8 + 4] = t_wait_a;  // cycles this step spent waiting for activations ...
            trace[st * 8 + 5] = t_wait_w;  // ... and for weight stages
          }
          if (kb + 1 < sp.nkb) {
            // A_hi W_lo over the four K-steps (all N columns), release the lo stage, then A_hi W_hi + A_lo W_hi: the lo stage is
            // back with the producer two thirds of a K block earlier
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ts_f16(d_addr, a_kb + ks * 16, umma_desc_at(lo_lo, ks * 32), idesc, (kb | ks) ? 1u : 0u);
            umma_commit_multicast(&s.c->w_empty[st_lo], (uint16_t)((1u << CLUSTER) - 1u));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t bh = umma_desc_at(lo_hi, ks * 32);
              umma_ts_f16(d_addr, a_kb + ks * 16, bh, idesc, 1u);
              umma_ts_f16(d_addr, a_kb + ks * 16 + 8, bh, idesc, 1u);
            }
            umma_commit_multicast(&s.c->w_empty[st_hi], (uint16_t)((1u << CLUSTER) - 1u));
          } else {
            // last K block of the step: columns [0, 64) first, committed on their own (d_q[buf][0]) so that epilogue pass 0
            // starts a quarter of a K block after the last a_ready; the remaining columns follow underneath that pass
            const uint32_t n0 = sp.n_pad < 64 ? (uint32_t)sp.n_pad : 64u;
            const uint32_t id0 = umma_idesc(n0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t bh = umma_desc_at(lo_hi, ks * 32), bl = umma_desc_at(lo_lo, ks * 32);
              umma_ts_f16(d_addr, a_kb + ks * 16, bh, id0, (kb | ks) ? 1u : 0u);
              umma_ts_f16(d_addr, a_kb + ks * 16 + 8, bh, id0, 1u);
              umma_ts_f16(d_addr, a_kb + ks * 16, bl, id0, 1u);
            }
            umma_commit(&s.c->d_q[buf][0]);
            if (sp.n_pad > 64) {
              const uint32_t id1 = umma_idesc((uint32_t)sp.n_pad - 64u);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {  // weight rows 64.. start 64 * 128 bytes into the tile
                const uint64_t bh = umma_desc_at(lo_hi, 8192 + ks * 32), bl = umma_desc_at(lo_lo, 8192 + ks * 32);
                umma_ts_f16(d_addr + 64u, a_kb + ks * 16, bh, id1, (kb | ks) ? 1u : 0u);
                umma_ts_f16(d_addr + 64u, a_kb + ks * 16 + 8, bh, id1, 1u);
                umma_ts_f16(d_addr + 64u, a_kb + ks * 16, bl, id1, 1u);
              }
            }
            umma_commit(&s.c->d_q[buf][1]);
            umma_commit_multicast(&s.c->w_empty[st_lo], (uint16_t)((1u << CLUSTER) - 1u));
            umma_commit_multicast(&s.c->w_empty[st_hi], (uint16_t)((1u << CLUSTER) - 1u));
          }
          if (trace && it == TRACE_ITER && blockIdx.x == 0 && kb + 1 == sp.nkb) trace[st * 8 + 7] = clock64();
        }
        __syncwarp();
      }
    }
  }
}

// ---- epilogue-side helpers ---------------------------------------------------------------------------------------------
struct EpiCtx {
  uint32_t tmem_base;
  uint32_t step_ctr;   // global step counter (same sequence as the MMA warp)
  int row;             // tile row owned by this thread (TMEM lane)
  int sub;             // 0..3: owns columns [64 p + 16 sub, +16) in pass p
  uint32_t lane_addr;  // (32 * quadrant) << 16
  // first TMEM column of the current step's accumulator / of the region the current step's A operand lives in
  __device__ __forceinline__ uint32_t d_col0() const { return (step_ctr & 1u) * 256u; }
  __device__ __forceinline__ uint32_t a_col0() const { return ((step_ctr & 1u) ^ 1u) * 256u; }
};
__device__ __forceinline__ EpiCtx epi_ctx(uint32_t tmem_base) {
  EpiCtx e;
  const int warp = threadIdx.x >> 5;
  const int q = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
  e.tmem_base = tmem_base;
  e.step_ctr = 0;
  e.row = q * 32 + (threadIdx.x & 31);
  e.sub = (warp - EPI_WARP0) >> 2;
  e.lane_addr = (uint32_t)(q * 32) << 16;
  return e;
}
// wait for part q of the current step's accumulator: 0 = columns [0, 64), 1 = the rest
__device__ __forceinline__ void epi_wait_q(const Smem& s, const EpiCtx& e, int q) {
  mbar_wait(&s.c->d_q[e.step_ctr & 1u][q], (e.step_ctr >> 1) & 1u);
  tc_fence_after();
}
// PSN_EPI_SKEW (cycles per sub; 0 = off): the four epilogue warps of a scheduler leave the accumulator wait of pass 0 in the same
// cycle and then walk ld -> MUFU burst -> ALU / convert -> st in lock-step, so the XU pipe is oversubscribed in one phase and idle
// in the next.  When the wait actually blocked (epilogue-bound regime), sub s starts s * PSN_EPI_SKEW cycles late: the bursts of the
// four warps then interleave instead of colliding.  Nothing re-synchronises the warps until the next step's pass 0.
#ifndef PSN_EPI_SKEW
#define PSN_EPI_SKEW 0
#endif
#ifndef PSN_EPI_UNROLL
#define PSN_EPI_UNROLL 1  // the four passes of epi_for_chunks_pf fully unrolled: no register copies of the prefetched side loads (16 moves per pass);
                          // r2 A/B on the relit view: shadow pass 334.2 -> 320.9 ms.  0 = rolled loop (the round-1 form)
#endif
__device__ __forceinline__ void epi_wait_q0_skewed(const Smem& s, const EpiCtx& e) {
  void* bar = &s.c->d_q[e.step_ctr & 1u][0];
  const uint32_t par = (e.step_ctr >> 1) & 1u;
  if (PSN_EPI_SKEW > 0) {
    uint32_t done;  // test_wait: non-blocking (try_wait suspends in hardware and would report "not blocked" after sleeping)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(par) : "memory");
    if (!done) {
      mbar_wait(bar, par);
      const unsigned d = (unsigned)e.sub * (unsigned)PSN_EPI_SKEW;
      const unsigned t0 = (unsigned)clock();
      while ((unsigned)clock() - t0 < d) {
      }
    }
  } else {
    mbar_wait(bar, par);
  }
  tc_fence_after();
}
// wait for the whole accumulator of the current step
__device__ __forceinline__ void epi_wait_d(const Smem& s, const EpiCtx& e) {
  mbar_wait(&s.c->d_q[e.step_ctr & 1u][0], (e.step_ctr >> 1) & 1u);
  mbar_wait(&s.c->d_q[e.step_ctr & 1u][1], (e.step_ctr >> 1) & 1u);
  tc_fence_after();
}
__device__ __forceinline__ void epi_load16(const EpiCtx& e, int col, float (&v)[CW]) {
  tmem_ld16(e.tmem_base + e.lane_addr + e.d_col0() + (uint32_t)col, v);
}

// Visit the 16-column chunks this thread owns in the current accumulator: f(pass, col, v[16], buf), col = 64*pass + 16*sub.  Pass 0
// waits for columns [0, 64) only; one pass of the four subs covers exactly one 64-column K block of the next step's A operand.
// The walk carries a software-pipelined side load: pre(col, buf) issues the global / L2 loads a chunk needs (bias, stashed
// sigma', parked partials, per-point tables) into a register struct; it runs for pass 0 BEFORE the accumulator wait and for pass
// p+1 between the issue and the wait of pass p's TMEM load, so the ~700-cycle L2 latency of those loads is off the critical
// path of every pass (with four epilogue warps per scheduler it used to be exposed four times per layer).
// UNROLL: the four passes fully unrolled (no register copies of the prefetched side loads: 16 moves per pass; r2: shadow pass 334 -> 321
// ms) - the default of the occupancy / visibility kernels.  The radiance kernel keeps the rolled loop: unrolled it is no faster and the
// write-backs of its dead scratch lines rose from 9.8 to 37 GB per launch (the discards land later relative to the evictions).
template <class Buf, bool UNROLL = (PSN_EPI_UNROLL != 0), class Pre, class F>
__device__ __forceinline__ void epi_for_chunks_pf(const Smem& s, const EpiCtx& e, Pre&& pre, F&& f) {
  const uint32_t base = e.tmem_base + e.lane_addr + e.d_col0();
  Buf nxt;
  pre(CW * e.sub, nxt);
  auto body = [&](int pass) {
    const Buf cur = nxt;
    if (pass == 0) epi_wait_q0_skewed(s, e);
    else if (pass == 1) epi_wait_q(s, e, 1);
    const int col = 64 * pass + CW * e.sub;
    uint32_t r[CW];
    tmem_ld16_issue(base + (uint32_t)col, r);
    if (pass < 3) pre(col + 64, nxt);
    tmem_ld16_wait(r);
    float v[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]);
    f(pass, col, v, cur);
  };
  if (UNROLL) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) body(pass);
  } else {
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) body(pass);
  }
}
// Same walk with the TMEM chunk of pass p + 1 requested BEFORE pass p is processed (passes 1..3: their columns are complete once
// d_q[.][1] has fired), so the tcgen05.ld latency of every pass but the first two hides under the arithmetic of its predecessor.
// Costs one more 16-register buffer: only the single-pass (CHEAP) occupancy program, whose packed-fp16 epilogue is short on registers
// to spare, uses it - the full epilogue spills with it (profiles/README.md item 11).
template <class Buf, class Pre, class F>
__device__ __forceinline__ void epi_for_chunks_pf_ld(const Smem& s, const EpiCtx& e, Pre&& pre, F&& f) {
  const uint32_t base = e.tmem_base + e.lane_addr + e.d_col0() + (uint32_t)(CW * e.sub);
  Buf nxt;
  pre(CW * e.sub, nxt);
  uint32_t r0[CW], r1[CW];
  epi_wait_q(s, e, 0);
  tmem_ld16_issue(base, r0);
  {  // pass 0 (the remaining columns may still be accumulating: nothing to prefetch yet)
    const Buf cur = nxt;
    pre(CW * e.sub + 64, nxt);
    tmem_ld16_wait(r0);
    float v[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r0[i]);
    f(0, CW * e.sub, v, cur);
  }
  epi_wait_q(s, e, 1);
  tmem_ld16_issue(base + 64u, r1);
#pragma unroll
  for (int pass = 1; pass < 4; ++pass) {
    const Buf cur = nxt;
    const int col = 64 * pass + CW * e.sub;
    if (pass < 3) pre(col + 64, nxt);
    uint32_t (&rc)[CW] = (pass & 1) ? r1 : r0;
    uint32_t (&rn)[CW] = (pass & 1) ? r0 : r1;
    tmem_ld16_wait(rc);
    if (pass < 3) tmem_ld16_issue(base + 64u * (uint32_t)(pass + 1), rn);
    float v[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(rc[i]);
    f(pass, col, v, cur);
  }
}
struct Bias16 { float4 b[4]; };
__device__ __forceinline__ void load_bias16(const float* __restrict__ bias, int col, Bias16& o) {
  const float4* b4 = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
  for (int t = 0; t < 4; ++t) o.b[t] = __ldg(b4 + t);
}
// v <- v * c + b (feature head of the geo net: its input is the scaled activation s_7 = P h_7, c = 1 / P)
__device__ __forceinline__ void fma16(float (&v)[CW], float c, const float4 (&b)[4]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    v[4 * t] = fmaf(v[4 * t], c, b[t].x); v[4 * t + 1] = fmaf(v[4 * t + 1], c, b[t].y);
    v[4 * t + 2] = fmaf(v[4 * t + 2], c, b[t].z); v[4 * t + 3] = fmaf(v[4 * t + 3], c, b[t].w);
  }
}
__device__ __forceinline__ void add16(float (&v)[CW], const float4 (&b)[4]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) { v[4 * t] += b[t].x; v[4 * t + 1] += b[t].y; v[4 * t + 2] += b[t].z; v[4 * t + 3] += b[t].w; }
}
// Write 16 consecutive activation values (K elements col..col+15 of this thread's row) as the A operand of K-step col/16 of the
// region starting at TMEM column col0: 8 columns of packed fp16 hi, then 8 columns of packed fp16 lo (see the file header).
// Packed cvt.rn.f16x2.f32 (F2FP, ALU pipe) - scalar F2F would queue on the XU pipe with the MUFUs.
__device__ __forceinline__ void epi_store_a16(const EpiCtx& e, uint32_t col0, int col, const float (&v)[CW]) {
  uint32_t r[16];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
#ifdef PSN_TC_SWAP_HALVES  // bring-up only: the other packing order of the two K elements of a column
    const float x0 = v[2 * u + 1], x1 = v[2 * u];
#else
    const float x0 = v[2 * u], x1 = v[2 * u + 1];
#endif
    const __half2 hh = __floats2half2_rn(x0, x1);  // .x (low half) = x0 = the even K element
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    r[u] = *reinterpret_cast<const uint32_t*>(&hh);
    r[8 + u] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  tmem_st16(e.tmem_base + e.lane_addr + col0 + (uint32_t)col, r);
}
// The same 16 activation values for a SINGLE-PASS consumer step (Step::single): only the packed fp16 hi half, at the columns
// where the full layout has it (the 8 lo columns keep stale data that no MMA of that step reads): half the conversions, half the
// tcgen05.st traffic.
__device__ __forceinline__ void epi_store_a16_hi(const EpiCtx& e, uint32_t col0, int col, const float (&v)[CW]) {
  uint32_t r[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const __half2 hh = __floats2half2_rn(v[2 * u], v[2 * u + 1]);
    r[u] = *reinterpret_cast<const uint32_t*>(&hh);
  }
  tmem_st8(e.tmem_base + e.lane_addr + col0 + (uint32_t)col, r);
}
__device__ __forceinline__ void epi_store_a16(const EpiCtx& e, uint32_t col0, int col, const float (&v)[CW], bool hi_only) {
  if (hi_only) epi_store_a16_hi(e, col0, col, v);
  else epi_store_a16(e, col0, col, v);
}
// this warp's 16 columns of K-block kb are written: publish to the MMA warp (one arrival per epilogue warp, 16 per block)
__device__ __forceinline__ void epi_signal_a(const Smem& s, int kb) {
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&s.c->a_ready[kb]);
}

// Point encoding [x, sin(2^0 x), cos(2^0 x), ...] (network.py:141-150) of this thread's row into the smem table pe[k][row]:
// sub s evaluates the sincos pairs m = s, s+4, ... (m = 3*octave + coordinate).  Callers follow with a barrier over the
// epilogue threads; the table then serves the layer-0 operand, the skip-layer concat and the encoding Jacobian.
#ifndef PSN_PE_DOUBLING
#define PSN_PE_DOUBLING 1  // octaves 1, 2 and 4, 5 from octaves 0 and 3 by angle doubling: 2 sincosf per thread and tile instead of up to 5 (the tensor pipe idles
                           // during the encoding).  r2 A/B: march 157.1 -> 153.7 ms, relit step 518.2 -> 513.2 ms; every measured parity error unchanged
                           // (117 tests, profiles/r2_parity_errlog_pe_doubling.jsonl).  0 = one sincosf per octave and coordinate
#endif
__device__ __forceinline__ void epi_write_pe(const Smem& s, int row, int sub, const float (&x)[3], int octaves) {
  if (sub == 0) { s.pe[row] = x[0]; s.pe[TILE_M + row] = x[1]; s.pe[2 * TILE_M + row] = x[2]; }
  if (PSN_PE_DOUBLING && octaves == 6) {
    // sub c (0..2) owns coordinate c: sincosf at octaves 0 and 3 (their arguments 2^k x are exact in fp32), the two octaves above
    // each by  sin 2a = 2 sin a cos a,  cos 2a = 1 - 2 sin^2 a  (|error| <= ~4 ulp of the base value after two doublings)
    if (sub < 3) {
      const float xc = sub == 0 ? x[0] : (sub == 1 ? x[1] : x[2]);
#pragma unroll
      for (int base = 0; base < 6; base += 3) {
        float sn, cs;
        sincosf((float)(1 << base) * xc, &sn, &cs);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int oct = base + d;
          s.pe[(3 + 6 * oct + sub) * TILE_M + row] = sn;
          s.pe[(6 + 6 * oct + sub) * TILE_M + row] = cs;
          const float s2 = 2.f * sn * cs, c2 = fmaf(-2.f * sn, sn, 1.f);
          sn = s2; cs = c2;
        }
      }
    }
    return;
  }
#pragma unroll 1
  for (int m = sub; m < 3 * octaves; m += EPI_SUBS) {
    const int oct = m / 3, c = m - 3 * oct;
    float sn, cs;
    sincosf((float)(1 << oct) * (c == 0 ? x[0] : (c == 1 ? x[1] : x[2])), &sn, &cs);
    s.pe[(3 + 6 * oct + c) * TILE_M + row] = sn;
    s.pe[(6 + 6 * oct + c) * TILE_M + row] = cs;
  }
}
// d pe[k] / d x[coord] from the table: 1 for k < 3; f cos(f x) = f * pe[k+3] for a sin entry; -f sin(f x) = -f * pe[k-3] for a cos entry
__device__ __forceinline__ float pe_jac_tab(const Smem& s, int row, int k, int* coord) {
  if (k < 3) { *coord = k; return 1.f; }
  const int j = k - 3, oct = j / 6, r = j - 6 * oct;
  const float f = (float)(1 << oct);
  if (r < 3) { *coord = r; return f * s.pe[(k + 3) * TILE_M + row]; }
  *coord = r - 3;
  return -f * s.pe[(k - 3) * TILE_M + row];
}

// softplus(beta=100) in the scaled domain (tc_pack.cu): zs = P z with P = 100 log2(e) is the accumulator, the activation handed on is
// s = P softplus(z) = max(zs, lg2(1 + 2^min(zs, 40))).  2 MUFU + 3 ALU ops and no multiplication by ln2 / 100 (it is folded into the
// packed weights of whatever consumes s); for zs > ~25 the lg2 term equals zs in fp32, so the max reproduces PyTorch's linear branch
// (threshold 20) to <1e-9 without a compare/select.
// (Measured alternatives, profiles/README.md: one MUFU + a degree-6 FMA-pipe polynomial for lg2(1 + u) halves the XU load but
// costs five more issue slots per activation - the march kernel got 8 % slower; applying it to every 2nd / 4th element was
// within run-to-run noise of this form, so the two-MUFU form stays.)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#define PSN_SOFTPLUS_C 0.0069314718055994531f  /* ln2 / 100 = 1 / P: s -> h (bring-up dumps only) */
__device__ __forceinline__ float softplus_scaled(float zs) {
  const float e = ex2_approx(fminf(zs, 40.f));
  return fmaxf(zs, lg2_approx(1.f + e));
}
// First level of the two-level march only (k_tc_occ<.., CHEAP>): s = max(zs, 0) + lg2(1 + u), u = 2^-|zs| in (0, 1], with
// lg2(1 + u) ~ u (C1 + C2 u + C3 u^2) (no constant term; max error 7.7e-4, i.e. 5e-6 on the activation): ONE MUFU.
// The values this produces are only trusted away from the occupancy threshold
// (tests/precision_study.py: max |cheap - full| = 1e-3 on alpha against a refine margin of 0.02).
__device__ __forceinline__ float softplus_scaled_cheap(float zs) {
  const float u = ex2_approx(-fabsf(zs));
  const float q = fmaf(fmaf(0.165381165f, u, -0.589203729f), u, 1.42459315f);
  return fmaf(q, u, fmaxf(zs, 0.f));
}
// The same activation for TWO accumulators entirely in packed fp16 arithmetic (ex2.approx.f16x2, HFMA2, HMNMX2): the result
// is directly one 32-bit column of the single-pass A operand, so neither a second MUFU nor a conversion follows.  An fp16 epilogue
// adds about as much error as the fp16 operands themselves (emulated: max |cheap - full| = 1.5e-3 on alpha, margin 0.02).
__device__ __forceinline__ uint32_t softplus_scaled_cheap_h2(float zs0, float zs1) {
  const __half2 zs = __floats2half2_rn(zs0, zs1);
  const __half2 na = __hneg2(__habs2(zs));
  uint32_t ub;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(ub) : "r"(*reinterpret_cast<const uint32_t*>(&na)));
  const __half2 u = *reinterpret_cast<const __half2*>(&ub);
  __half2 q = __hfma2(__float2half2_rn(0.165381165f), u, __float2half2_rn(-0.589203729f));
  q = __hfma2(q, u, __float2half2_rn(1.42459315f));
  const __half2 r = __hfma2(q, u, __hmax2(zs, __float2half2_rn(0.f)));
  return *reinterpret_cast<const uint32_t*>(&r);
}

}  // namespace tc
}  // namespace psn
