// fp32 GEMM + element-wise pieces shared by the train-step translation units (stage2_train.cu, stage1_train.cu).
#pragma once
#include "common.cuh"

namespace psn {

// ---------------------------------------------------------------------------------------------------------------------
// fp32 GEMM, 64x64x16 tiles, 256 threads, 4x4 per thread.
//   FORM 0 (NT): C[m,n] = sum_k A[m,k] B[n,k]      forward Linear        (A = X [M,K], B = W [N,K])
//   FORM 1 (NN): C[m,n] = sum_k A[m,k] B[k,n]      input gradient        (A = dZ [M,K], B = W [K,N])
//   FORM 2 (TN): C[m,n] += sum_k A[k,m] B[k,n]     weight gradient       (A = dZ [K,M], B = X [K,N]); split over k, atomicAdd
// EPI: 0 none, 1 +bias, 2 +bias relu, 3 +bias sigmoid
// ---------------------------------------------------------------------------------------------------------------------
template <int FORM>
__global__ void __launch_bounds__(256)
k_gemm(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
       const float* __restrict__ bias, int M, int N, int K, int epi, int k_per_split) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg = (FORM == 2) ? blockIdx.z * k_per_split : 0;
  const int kend = (FORM == 2) ? min(K, kbeg + k_per_split) : K;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = threadIdx.x + 256 * j;
      if (FORM == 2) {
        const int m = idx & 63, k = idx >> 6;
        As[k][m] = (k0 + k < kend && m0 + m < M) ? A[(size_t)(k0 + k) * lda + m0 + m] : 0.f;
      } else {
        const int k = idx & 15, m = idx >> 4;
        As[k][m] = (k0 + k < kend && m0 + m < M) ? A[(size_t)(m0 + m) * lda + k0 + k] : 0.f;
      }
      if (FORM == 0) {
        const int k = idx & 15, n = idx >> 4;
        Bs[k][n] = (k0 + k < kend && n0 + n < N) ? B[(size_t)(n0 + n) * ldb + k0 + k] : 0.f;
      } else {
        const int n = idx & 63, k = idx >> 6;
        Bs[k][n] = (k0 + k < kend && n0 + n < N) ? B[(size_t)(k0 + k) * ldb + n0 + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (FORM == 2) {
        atomicAdd(&C[(size_t)m * ldc + n], v);
      } else {
        if (epi >= 1) v += bias[n];
        if (epi == 2) v = fmaxf(v, 0.f);
        if (epi == 3) v = 1.f / (1.f + expf(-v));
        C[(size_t)m * ldc + n] = v;
      }
    }
  }
}

// dZ[r, c] = dY[r, c] * act'(Y[r, c]);  kind 0: identity, 2: relu (Y > 0), 3: sigmoid (Y (1 - Y))
static __global__ void k_act_bwd(const float* __restrict__ dY, int lddy, const float* __restrict__ Y, int ldy, float* __restrict__ dZ, int lddz,
                          long long rows, int ncols, int kind) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ncols) return;
  const long long r = i / ncols;
  const int c = (int)(i - r * ncols);
  const float g = dY[r * lddy + c], y = Y[r * ldy + c];
  dZ[r * lddz + c] = kind == 2 ? (y > 0.f ? g : 0.f) : kind == 3 ? g * y * (1.f - y) : g;
}
// Bias gradient db[c] += sum_r dZ[r, c].  A block sums COLSUM_ROWS rows of a 64-column strip: thread (tx, ty) walks rows ty, ty + 4,
// ... with eight independent accumulators (the round-1 version walked 1024 rows per thread in one dependent chain on 32 blocks:
// 58 us for a 32768 x 256 matrix, longer than the weight-gradient GEMM next to it), the four row lanes are combined in shared memory
// and every column costs one atomic per block.
constexpr int COLSUM_ROWS = 256;
static __global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ dZ, int ld, long long rows, int ncols, float* __restrict__ db) {
  __shared__ float part[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + tx;
  const long long r0 = (long long)blockIdx.y * COLSUM_ROWS, r1 = min(rows, r0 + COLSUM_ROWS);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < ncols) {
    const float* src = dZ + c;
    long long r = r0 + ty;
    for (; r + 28 < r1; r += 32) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += __ldg(src + (r + 4 * u) * ld);
    }
    for (; r < r1; r += 4) acc[0] += __ldg(src + r * ld);
  }
  part[ty][tx] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (ty == 0 && c < ncols) atomicAdd(&db[c], (part[0][tx] + part[1][tx]) + (part[2][tx] + part[3][tx]));
}
static inline void colsum(const float* dZ, int ld, long long rows, int ncols, float* db, cudaStream_t st) {
  if (rows <= 0 || ncols <= 0) return;
  k_colsum<<<dim3((unsigned)((ncols + 63) / 64), (unsigned)((rows + COLSUM_ROWS - 1) / COLSUM_ROWS)), 256, 0, st>>>(dZ, ld, rows, ncols, db);
}
static inline unsigned nblk(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// tc_gemm.cu: the same three products on the tensor cores (tcgen05 kind::tf32, x = hi + lo split, three passes, fp32 accumulate).
// It is what the train steps run; PSNERF_B200_TRAIN_GEMM=ffma selects the FFMA kernel above (A/B measurements, cross-check).
// Fused element-wise epilogues of the tensor-core kernel (FORM 0 / 1; all matrices [M, N] with leading dimension lde):
//   EPI 4  C = softplus_100(acc + bias), C2 = sigmoid(100 (acc + bias))                       (k_s1_softplus)
//   EPI 5  C = acc * E1, E2 := acc * E2 * 100 E1 (1 - E1)                                     (k_s1_second)
//   EPI 6  C = acc * scale * E1 + E2                                                          (k_s1_zbar)
//   EPI 7  C = acc, C2 = acc * E1                                                             (k_s1_mul on the GEMM's own output)
int tc_gemm(int form, const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc, const float* bias, long long M,
            int N, long long K, int epi, cudaStream_t st, const GemmFuse* fz);
bool train_gemm_use_tc();
bool train_gemm_fused();  // the fused epilogues are available (tensor-core GEMM, second generation)

static int gemm(int form, const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias, long long M, int N,
                long long K, int epi, cudaStream_t st, const GemmFuse* fz = nullptr) {
  if (M == 0 || N == 0 || K == 0) return PSN_OK;
  if (train_gemm_use_tc()) return tc_gemm(form, A, lda, B, ldb, C, ldc, bias, M, N, K, epi, st, fz);
  PSN_REQUIRE(epi <= 3, PSN_ERR_ARG, "gemm: fused epilogue %d on the FFMA path", epi);
  count_launch();
  if (form == 0) {
    dim3 g((N + 63) / 64, (unsigned)((M + 63) / 64));
    k_gemm<0><<<g, 256, 0, st>>>(A, lda, B, ldb, C, ldc, bias, (int)M, N, (int)K, epi, 0);
  } else if (form == 1) {
    dim3 g((N + 63) / 64, (unsigned)((M + 63) / 64));
    k_gemm<1><<<g, 256, 0, st>>>(A, lda, B, ldb, C, ldc, bias, (int)M, N, (int)K, epi, 0);
  } else {
    const int kps = 512;
    dim3 g((N + 63) / 64, (unsigned)((M + 63) / 64), (unsigned)((K + kps - 1) / kps));
    k_gemm<2><<<g, 256, 0, st>>>(A, lda, B, ldb, C, ldc, bias, (int)M, N, (int)K, 0, kps);
  }
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}


}  // namespace psn
