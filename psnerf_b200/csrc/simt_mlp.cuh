// fp32 FFMA building block: one dense layer over a 64-row tile held in shared memory.
//
// Layout: activations are feature-major  X[k][LDX]  (LDX = 68: 64 rows + pad so that the epilogue's
// per-column float4 stores are bank-conflict free).  Weights are global, k-major [K_pad][N_pad], staged
// 16 k-rows at a time through a 3-stage cp.async ring.  256 threads: lane tx owns NC output
// columns, warp ty owns rows 8*ty..8*ty+7 (A-reads are warp broadcasts, W-reads are conflict-free
// float4).  Accumulators stay in registers until every input row has been consumed, so the epilogue may
// overwrite the input buffer (layers run in place).
#pragma once
#include "common.cuh"

namespace psn {

constexpr int TM = 64;    // rows (samples) per tile
constexpr int LDX = 68;   // leading dimension of feature-major smem buffers
constexpr int NT = 256;   // threads per CTA
constexpr int KC = 16;    // k rows per weight stage
constexpr int WSTAGE_FLOATS = KC * 256;  // one stage holds up to N_pad = 256 columns
constexpr int WRING_FLOATS = 3 * WSTAGE_FLOATS;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// column owned by (lane tx, slot j)
template <int NC>
__device__ __forceinline__ int simt_col(int tx, int j) {
  if (NC == 8) return tx * 4 + (j & 3) + 128 * (j >> 2);
  if (NC == 4) return tx * 4 + j;
  if (NC == 2) return tx * 2 + j;
  return tx;
}

__device__ __forceinline__ void stage_weights(const float* __restrict__ wt, int n_pad, int chunk, float* dst) {
  const float4* src = reinterpret_cast<const float4*>(wt + (size_t)chunk * KC * n_pad);
  const int n4 = KC * n_pad / 4;
  for (int i = threadIdx.x; i < n4; i += NT) cp_async16(reinterpret_cast<float4*>(dst) + i, src + i);
}

// acc[i][j] = sum_k xs[k][8*ty+i] * wt[k][col(j)]   (+ bias).  Ends with a __syncthreads(): on return every
// thread may overwrite xs.  The caller must __syncthreads() after its epilogue before the next dense().
// wstage: WRING_FLOATS floats of shared memory.
template <int NC>
__device__ __forceinline__ void dense(const SimtLayer& L, const float* __restrict__ xs, float* wstage,
                                      float (&acc)[8][NC]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n_pad = L.N_pad;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const float b = L.bias ? L.bias[simt_col<NC>(tx, j)] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][j] = b;
  }
  const int nchunk = L.K_pad / KC;
  // 3-stage cp.async ring, one barrier per chunk: the barrier of iteration c proves every thread is done
  // with chunk c-1, whose stage is the one chunk c+2 is then streamed into.
  stage_weights(L.wt, n_pad, 0, wstage);
  cp_async_commit();
  if (nchunk > 1) stage_weights(L.wt, n_pad, 1, wstage + WSTAGE_FLOATS);
  cp_async_commit();
  for (int c = 0; c < nchunk; ++c) {
    cp_async_wait<1>();
    __syncthreads();
    if (c + 2 < nchunk) stage_weights(L.wt, n_pad, c + 2, wstage + ((c + 2) % 3) * WSTAGE_FLOATS);
    cp_async_commit();
    const float* ws = wstage + (c % 3) * WSTAGE_FLOATS;
    const float* xk = xs + (size_t)c * KC * LDX + ty * 8;
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(xk + kk * LDX);
      const float4 a1 = *reinterpret_cast<const float4*>(xk + kk * LDX + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float w[NC];
      if constexpr (NC == 8) {
        const float4 w0 = *reinterpret_cast<const float4*>(ws + kk * n_pad + tx * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + kk * n_pad + 128 + tx * 4);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
        w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
      } else if constexpr (NC == 4) {
        const float4 w0 = *reinterpret_cast<const float4*>(ws + kk * n_pad + tx * 4);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
      } else if constexpr (NC == 2) {
        const float2 w0 = *reinterpret_cast<const float2*>(ws + kk * n_pad + tx * 2);
        w[0] = w0.x; w[1] = w0.y;
      } else {
        w[0] = ws[kk * n_pad + tx];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// Store a thread's 8 rows of column `col` into a feature-major buffer (two float4).
__device__ __forceinline__ void store_col8(float* dst, int col, int ty, const float v[8]) {
  float4* p = reinterpret_cast<float4*>(dst + (size_t)col * LDX + ty * 8);
  p[0] = make_float4(v[0], v[1], v[2], v[3]);
  p[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// nn.Softplus(beta=100, threshold=20): x if 100x > 20 else log1p(exp(100x))/100   (network.py:68)
__device__ __forceinline__ float softplus100(float z) {
  const float v = z * 100.f;
  return v > 20.f ? z : log1pf(expf(v)) / 100.f;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace psn
