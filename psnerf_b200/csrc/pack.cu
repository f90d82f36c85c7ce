// Library state, error reporting and network packing (psn_mlp_create / psn_mlp_free).
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>

#include "common.cuh"

namespace psn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// dst[k][n] = W[(row0+n)*in + k]  for k < in, n < rows; zero elsewhere.  (k-major copy used by the forward GEMM)
__global__ void pack_kmajor(const float* __restrict__ W, int in, int row0, int rows, int k_pad, int n_pad,
                            float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k_pad * n_pad) return;
  const int k = idx / n_pad, n = idx % n_pad;
  dst[idx] = (k < in && n < rows) ? W[(size_t)(row0 + n) * in + k] : 0.f;
}
// dst[k][n] = W[k*in + n] for k < out, n < in; zero elsewhere.  (native copy used by the reverse GEMM dx = dz W)
__global__ void pack_native(const float* __restrict__ W, int out, int in, int k_pad, int n_pad,
                            float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k_pad * n_pad) return;
  const int k = idx / n_pad, n = idx % n_pad;
  dst[idx] = (k < out && n < in) ? W[(size_t)k * in + n] : 0.f;
}
__global__ void pack_vec(const float* __restrict__ src, int off, int n, int n_pad, float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad) return;
  dst[idx] = idx < n ? src[off + idx] : 0.f;
}

static int n_pad_of(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : -1; }

int tc_pack_bytes(const psn_mlp* net);                                     // tc_pack.cu
int tc_pack_fill(psn_mlp* net, const float* const* W, const float* const* b, char* base, size_t off,
                 cudaStream_t st);                                         // tc_pack.cu

}  // namespace psn

using namespace psn;

extern "C" int psn_version(void) { return 100; }
extern "C" const char* psn_last_error(void) { return g_err; }
extern "C" int64_t psn_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int psn_device_check(int* sm_count) {
  int dev = 0;
  PSN_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  PSN_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  PSN_REQUIRE(prop.major == 10, PSN_ERR_CUDA, "psnerf_b200 needs a compute-capability 10.x GPU (sm_100a), found %d.%d",
              prop.major, prop.minor);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  return PSN_OK;
}

extern "C" int psn_mlp_create(const psn_mlp_desc* desc, const int* in_dims, const int* out_dims,
                              const float* const* W, const float* const* b, void* stream, psn_mlp** out) {
  PSN_REQUIRE(desc && in_dims && out_dims && W && b && out, PSN_ERR_ARG, "psn_mlp_create: null argument");
  const int nl = desc->n_layers;
  PSN_REQUIRE(nl >= 1 && nl <= kMaxLayers - 1, PSN_ERR_SHAPE, "psn_mlp_create: n_layers=%d unsupported", nl);
  PSN_REQUIRE(desc->kind >= 0 && desc->kind <= 2, PSN_ERR_ARG, "psn_mlp_create: bad kind %d", desc->kind);
  cudaStream_t st = (cudaStream_t)stream;
  psn_mlp* net = (psn_mlp*)calloc(1, sizeof(psn_mlp));
  net->kind = desc->kind;
  net->desc = *desc;
  net->n_layers = nl;
  for (int l = 0; l < nl; ++l) {
    net->in_dims[l] = in_dims[l];
    net->out_dims[l] = out_dims[l];
  }
  const bool geo = desc->kind == PSN_NET_GEO;
  // ---- shapes -------------------------------------------------------------------------------------------
  size_t floats = 0;
  auto reserve = [&](size_t n) { size_t o = floats; floats += (n + 63) / 64 * 64; return o; };
  size_t off_wt[kMaxLayers], off_b[kMaxLayers], off_rev[kMaxLayers], off_logit_w = 0, off_logit_b = 0, off_row = 0;
  for (int l = 0; l < nl; ++l) {
    int N = out_dims[l];
    if (geo && l == nl - 1) N -= 1;  // feature head; row 0 (logit) is packed separately
    const int np = n_pad_of(N), kp = pad_to(in_dims[l], 16);
    if (np < 0 || in_dims[l] > 384 || N < 1) {
      free(net);
      PSN_REQUIRE(false, PSN_ERR_SHAPE, "psn_mlp_create: layer %d shape [%d,%d] unsupported (out<=256, in<=384)", l,
                  out_dims[l], in_dims[l]);
    }
    net->fwd[l].K = in_dims[l];
    net->fwd[l].N = N;
    net->fwd[l].K_pad = kp;
    net->fwd[l].N_pad = np;
    off_wt[l] = reserve((size_t)kp * np);
    off_b[l] = reserve(np);
  }
  if (geo) {
    const int kp = pad_to(in_dims[nl - 1], 16);
    net->logit_head.K = in_dims[nl - 1];
    net->logit_head.N = 1;
    net->logit_head.K_pad = kp;
    net->logit_head.N_pad = 32;
    off_logit_w = reserve((size_t)kp * 32);
    off_logit_b = reserve(32);
    off_row = reserve(kp);
    for (int l = 0; l < nl - 1; ++l) {
      const int kp2 = pad_to(out_dims[l], 16), np2 = n_pad_of(in_dims[l]);
      if (np2 < 0) {
        free(net);
        PSN_REQUIRE(false, PSN_ERR_SHAPE, "psn_mlp_create: geo layer %d input %d > 256 unsupported", l, in_dims[l]);
      }
      net->rev[l].K = out_dims[l];
      net->rev[l].N = in_dims[l];
      net->rev[l].K_pad = kp2;
      net->rev[l].N_pad = np2;
      off_rev[l] = reserve((size_t)kp2 * np2);
    }
  }
  const int tc_bytes = tc_pack_bytes(net);
  const size_t fp32_bytes = floats * sizeof(float);
  net->blob_bytes = fp32_bytes + (tc_bytes > 0 ? (size_t)tc_bytes + 1024 : 0);
  cudaError_t e = cudaMalloc(&net->device_blob, net->blob_bytes);
  if (e != cudaSuccess) {
    const size_t want = net->blob_bytes;
    free(net);
    PSN_REQUIRE(false, PSN_ERR_CUDA, "psn_mlp_create: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  float* base = (float*)net->device_blob;
  // ---- fill ---------------------------------------------------------------------------------------------
  for (int l = 0; l < nl; ++l) {
    SimtLayer& L = net->fwd[l];
    const int row0 = (geo && l == nl - 1) ? 1 : 0;
    const int tot = L.K_pad * L.N_pad;
    psn::count_launch();
    pack_kmajor<<<(tot + 255) / 256, 256, 0, st>>>(W[l], in_dims[l], row0, L.N, L.K_pad, L.N_pad, base + off_wt[l]);
    psn::count_launch();
    pack_vec<<<(L.N_pad + 255) / 256, 256, 0, st>>>(b[l], row0, L.N, L.N_pad, base + off_b[l]);
    L.wt = base + off_wt[l];
    L.bias = base + off_b[l];
  }
  if (geo) {
    SimtLayer& H = net->logit_head;
    psn::count_launch();
    pack_kmajor<<<(H.K_pad * 32 + 255) / 256, 256, 0, st>>>(W[nl - 1], in_dims[nl - 1], 0, 1, H.K_pad, 32,
                                                             base + off_logit_w);
    psn::count_launch();
    pack_vec<<<1, 256, 0, st>>>(b[nl - 1], 0, 1, 32, base + off_logit_b);
    psn::count_launch();
    pack_vec<<<(H.K_pad + 255) / 256, 256, 0, st>>>(W[nl - 1], 0, in_dims[nl - 1], H.K_pad, base + off_row);
    H.wt = base + off_logit_w;
    H.bias = base + off_logit_b;
    net->w_logit_row = base + off_row;
    for (int l = 0; l < nl - 1; ++l) {
      SimtLayer& R = net->rev[l];
      const int tot = R.K_pad * R.N_pad;
      psn::count_launch();
      pack_native<<<(tot + 255) / 256, 256, 0, st>>>(W[l], out_dims[l], in_dims[l], R.K_pad, R.N_pad, base + off_rev[l]);
      R.wt = base + off_rev[l];
      R.bias = nullptr;
    }
  }
  net->tc_ok = 0;
  if (tc_bytes > 0) {
    size_t off = (fp32_bytes + 1023) / 1024 * 1024;
    if (tc_pack_fill(net, W, b, (char*)net->device_blob, off, st) == PSN_OK) net->tc_ok = 1;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(net->device_blob);
    free(net);
    PSN_REQUIRE(false, PSN_ERR_CUDA, "psn_mlp_create: pack kernels failed: %s", cudaGetErrorString(e));
  }
  *out = net;
  return PSN_OK;
}

extern "C" int psn_mlp_free(psn_mlp* net) {
  if (!net) return PSN_OK;
  if (net->device_blob) cudaFree(net->device_blob);
  free(net);
  return PSN_OK;
}
