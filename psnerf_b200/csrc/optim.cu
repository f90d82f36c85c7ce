// Fused optimizer steps of the two train loops: torch.optim.Adam over every parameter of a model in ONE
// launch (stage1/train.py:62, stage2/trainer.py:116) and torch.optim.SparseAdam over the rows of the light embeddings that a
// batch touched (stage2/trainer.py:165).  Both are HBM/L2-bound element-wise passes over <= 3.2 MB of state, so the cost is
// launches, not bytes: the torch optimizers issue several kernels per parameter group (foreach) or per tensor, these issue one.
//
// Adam (torch/optim/_functional.py adam(), amsgrad = False), per element with t = step after the increment:
//     g  = grad + weight_decay * p
//     m  = beta1 m + (1 - beta1) g ;  v = beta2 v + (1 - beta2) g g
//     p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// SparseAdam (torch/optim/_functional.py sparse_adam()), only for the rows present in the (coalesced) sparse gradient:
//     m += (1 - beta1)(g - m) ;  v += (1 - beta2)(g g - v) ;  p -= lr sqrt(1 - beta2^t) / (1 - beta1^t) * m / (sqrt(v) + eps)
// The two differ in where eps sits relative to the bias correction; both are kept as the reference's optimizers have them.
// The hyper-parameters arrive as doubles and every derived scalar (1 - beta, the bias corrections, lr / bc1) is evaluated on the
// host in double before it is rounded to fp32, as torch does: (float)(1 - 0.999) is 1.3e-5 away from 1.f - (float)0.999.
#include "launch.cuh"
#include "internal.cuh"

#include <math.h>

namespace psn {

constexpr int kAdamMaxTensors = 64;  // per launch; the table travels in the kernel parameter space (2.6 KB of the 4 KB)

struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  long long first[kAdamMaxTensors + 1];  // first float4-unit of tensor i in the launch's unit numbering
  long long numel[kAdamMaxTensors];
  int n;
};

struct AdamScalars {
  float beta1, beta2, one_m_beta1, one_m_beta2, eps, weight_decay, step_size, inv_bc2_sqrt;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamScalars& s) {
  g = fmaf(s.weight_decay, p, g);
  m = fmaf(s.beta1, m, s.one_m_beta1 * g);
  v = fmaf(s.beta2, v, s.one_m_beta2 * g * g);
  const float denom = fmaf(sqrtf(v), s.inv_bc2_sqrt, s.eps);
  p = fmaf(-s.step_size, __fdiv_rn(m, denom), p);
}

// One thread per float4 unit; a unit never straddles two tensors (each tensor owns ceil(numel / 4) units).
__global__ void __launch_bounds__(256) k_adam(const __grid_constant__ AdamTable tab, const AdamScalars s) {
  const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= tab.first[tab.n]) return;
  int lo = 0, hi = tab.n - 1;  // last tensor whose first unit <= u
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.first[mid] <= u) lo = mid; else hi = mid - 1;
  }
  const long long e0 = (u - tab.first[lo]) * 4;
  const long long left = tab.numel[lo] - e0;
  float* __restrict__ p = tab.p[lo] + e0;
  const float* __restrict__ g = tab.g[lo] + e0;
  float* __restrict__ m = tab.m[lo] + e0;
  float* __restrict__ v = tab.v[lo] + e0;
  const bool vec = left >= 4 && ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
  if (vec) {
    float4 P = *reinterpret_cast<float4*>(p), M = *reinterpret_cast<float4*>(m), V = *reinterpret_cast<float4*>(v);
    const float4 G = *reinterpret_cast<const float4*>(g);
    adam_elem(P.x, G.x, M.x, V.x, s);
    adam_elem(P.y, G.y, M.y, V.y, s);
    adam_elem(P.z, G.z, M.z, V.z, s);
    adam_elem(P.w, G.w, M.w, V.w, s);
    *reinterpret_cast<float4*>(p) = P;
    *reinterpret_cast<float4*>(m) = M;
    *reinterpret_cast<float4*>(v) = V;
  } else {
    const int n = left < 4 ? (int)left : 4;
    for (int i = 0; i < n; ++i) {
      float P = p[i], M = m[i], V = v[i];
      adam_elem(P, g[i], M, V, s);
      p[i] = P;
      m[i] = M;
      v[i] = V;
    }
  }
}

// One thread per (entry k, column d) of the sparse gradient.  Duplicate row indices are summed in entry order by the thread of
// the FIRST occurrence (what coalesce() does), the others retire: deterministic, no workspace, one launch.  K is the number of
// lights of a batch (<= a few hundred), so the O(K) scan per thread is a few hundred L1 hits.
__global__ void __launch_bounds__(128) k_sparse_adam(float* __restrict__ param, float* __restrict__ exp_avg,
                                                     float* __restrict__ exp_avg_sq, long long R, int D,
                                                     const long long* __restrict__ rows, const float* __restrict__ gv, long long K,
                                                     float one_m_beta1, float one_m_beta2, float eps, float step_size) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= K * D) return;
  const long long k = t / D;
  const int d = (int)(t - k * D);
  const long long r = rows[k];
  if (r < 0 || r >= R) return;
  for (long long j = 0; j < k; ++j)
    if (rows[j] == r) return;
  float g = gv[k * D + d];
  for (long long j = k + 1; j < K; ++j)
    if (rows[j] == r) g += gv[j * D + d];
  const long long o = r * D + d;
  const float m0 = exp_avg[o], v0 = exp_avg_sq[o];
  const float m = m0 + (g - m0) * one_m_beta1;
  const float v = v0 + (g * g - v0) * one_m_beta2;
  exp_avg[o] = m;
  exp_avg_sq[o] = v;
  param[o] -= step_size * __fdiv_rn(m, sqrtf(v) + eps);
}

static int check_hyper(const psn_adam_hyper* h, const char* who) {
  PSN_REQUIRE(h, PSN_ERR_ARG, "%s: null hyper-parameters", who);
  PSN_REQUIRE(h->step >= 1, PSN_ERR_ARG, "%s: step must be >= 1 (the count AFTER this update), got %lld", who, (long long)h->step);
  PSN_REQUIRE(h->lr >= 0.0 && h->eps >= 0.0 && h->weight_decay >= 0.0, PSN_ERR_ARG, "%s: negative lr / eps / weight_decay", who);
  PSN_REQUIRE(h->beta1 >= 0.0 && h->beta1 < 1.0 && h->beta2 >= 0.0 && h->beta2 < 1.0, PSN_ERR_ARG, "%s: betas must be in [0, 1)", who);
  return PSN_OK;
}

}  // namespace psn

using namespace psn;

extern "C" int psn_adam_step(const psn_adam_tensor* tensors, int n_tensors, const psn_adam_hyper* h, void* stream) {
  int rc = check_hyper(h, "psn_adam_step");
  if (rc) return rc;
  PSN_REQUIRE(n_tensors >= 0 && (tensors || n_tensors == 0), PSN_ERR_ARG, "psn_adam_step: bad tensor list");
  const double bc1 = 1.0 - pow(h->beta1, (double)h->step);
  const double bc2 = 1.0 - pow(h->beta2, (double)h->step);
  AdamScalars s;
  s.beta1 = (float)h->beta1;
  s.beta2 = (float)h->beta2;
  s.one_m_beta1 = (float)(1.0 - h->beta1);
  s.one_m_beta2 = (float)(1.0 - h->beta2);
  s.eps = (float)h->eps;
  s.weight_decay = (float)h->weight_decay;
  s.step_size = (float)(h->lr / bc1);
  s.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n_tensors; ++i) {
    const psn_adam_tensor& t = tensors[i];
    PSN_REQUIRE(t.numel >= 0, PSN_ERR_ARG, "psn_adam_step: tensor %d has negative numel", i);
    PSN_REQUIRE(t.numel == 0 || (t.param && t.grad && t.exp_avg && t.exp_avg_sq), PSN_ERR_ARG, "psn_adam_step: tensor %d has a null pointer", i);
  }
  int i0 = 0;
  bool launched = false;
  while (i0 < n_tensors) {
    AdamTable tab;
    memset(&tab, 0, sizeof(tab));
    long long units = 0;
    int n = 0;
    for (; i0 < n_tensors && n < kAdamMaxTensors; ++i0) {
      const psn_adam_tensor& t = tensors[i0];
      if (t.numel == 0) continue;
      tab.p[n] = t.param;
      tab.g[n] = t.grad;
      tab.m[n] = t.exp_avg;
      tab.v[n] = t.exp_avg_sq;
      tab.numel[n] = t.numel;
      tab.first[n] = units;
      units += (t.numel + 3) / 4;
      ++n;
    }
    if (n == 0) break;
    tab.first[n] = units;
    tab.n = n;
    const long long blocks = (units + 255) / 256;
    PSN_REQUIRE(blocks <= 0x7fffffffLL, PSN_ERR_SHAPE, "psn_adam_step: %lld elements in one launch", units * 4);
    k_adam<<<(unsigned)blocks, 256, 0, st>>>(tab, s);
    count_launch();
    launched = true;
  }
  if (launched) PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

extern "C" int psn_sparse_adam_step(float* param, float* exp_avg, float* exp_avg_sq, int64_t R, int D, const int64_t* rows,
                                    const float* grad_values, int64_t K, const psn_adam_hyper* h, void* stream) {
  int rc = check_hyper(h, "psn_sparse_adam_step");
  if (rc) return rc;
  PSN_REQUIRE(R >= 0 && D >= 1 && K >= 0, PSN_ERR_ARG, "psn_sparse_adam_step: bad shape R=%lld D=%d K=%lld", (long long)R, D, (long long)K);
  PSN_REQUIRE(h->weight_decay == 0.0, PSN_ERR_ARG, "psn_sparse_adam_step: SparseAdam has no weight decay");
  PSN_REQUIRE(K <= 65536, PSN_ERR_SHAPE, "psn_sparse_adam_step: %lld gradient rows > 65536 (coalesce on the caller's side)", (long long)K);
  if (K == 0 || R == 0) return PSN_OK;
  PSN_REQUIRE(param && exp_avg && exp_avg_sq && rows && grad_values, PSN_ERR_ARG, "psn_sparse_adam_step: null pointer");
  const double bc1 = 1.0 - pow(h->beta1, (double)h->step);
  const double bc2 = 1.0 - pow(h->beta2, (double)h->step);
  const float step_size = (float)(h->lr * sqrt(bc2) / bc1);
  const long long total = (long long)K * D;
  k_sparse_adam<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      param, exp_avg, exp_avg_sq, R, D, (const long long*)rows, grad_values, K, (float)(1.0 - h->beta1), (float)(1.0 - h->beta2), (float)h->eps,
      step_size);
  count_launch();
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
