// Shared declarations for the psnerf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/psnerf_b200.h"

namespace psn {

void set_error(const char* fmt, ...);
void count_launch();  // every kernel launch of this library is counted (psn_launch_count)

#define PSN_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      (void)cudaGetLastError(); /* do not leave a stale error for the caller's runtime (torch) */ \
      psn::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return PSN_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define PSN_REQUIRE(cond, code, ...)                                                          \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      psn::set_error(__VA_ARGS__);                                                            \
      return (code);                                                                          \
    }                                                                                         \
  } while (0)

inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }

// operands of the fused element-wise epilogues of the train-step GEMM (train_gemm.cuh / tc_gemm.cu): [M, N] matrices, leading dimension lde
struct GemmFuse { float* C2; const float* E1; float* E2; long long lde; float scale; };

// One Linear layer packed for the SIMT fp32 path: wt is [K_pad][N_pad] (k-major, i.e. the transpose of
// torch's [out,in] weight), zero padded; bias is [N_pad].
struct SimtLayer {
  const float* wt;
  const float* bias;
  int K, N, K_pad, N_pad;
};

constexpr int kMaxLayers = 12;

// One tcgen05 step: weight tiles [kb][hi,lo][n_pad rows][64] fp16, pre-swizzled (tc_mlp.cuh, tc_pack.cu).
struct TcStepW {
  unsigned int w_off;  // byte offset inside the net's tile blob
  unsigned short nkb, n_pad;
};
constexpr int kMaxTcSteps = 20;
// step indices per net kind
enum { TCG_FWD0 = 0, TCG_FEAT = 8, TCG_REV_TOP = 9 /* rev step of layer l is TCG_REV_TOP + (n_hidden-1-l) */, TCG_REV0 = 16 };
enum { TCA_L0F = 0, TCA_L0R = 1, TCA_L1 = 2, TCA_L4 = 5 };
enum { TCV_L1 = 0, TCV_L5Y = 4, TCV_L6 = 5, TCV_L7 = 6 };

}  // namespace psn

// Opaque handle behind the C ABI.  kind: 0 = stage-1 geo, 1 = stage-1 app, 2 = stage-2 MLP.
struct psn_mlp {
  int kind;
  psn_mlp_desc desc;
  int n_layers;
  int in_dims[psn::kMaxLayers], out_dims[psn::kMaxLayers];
  // forward layers (SIMT).  For kind 0 the last Linear is split: fwd[n_layers-1] = feature head
  // (rows 1..feat), logit_head = row 0.
  psn::SimtLayer fwd[psn::kMaxLayers];
  psn::SimtLayer logit_head;
  // reverse layers (kind 0 only): rev[l].wt is W_l in its native [out][in] layout, zero padded.
  psn::SimtLayer rev[psn::kMaxLayers];
  const float* w_logit_row;  // [hidden] row 0 of the last geo layer (d logit / d x_last)
  float b_logit;
  // tcgen05 packs (filled when the shape is one the tensor path supports)
  int tc_ok;
  psn::TcStepW tc_step[psn::kMaxTcSteps];
  const unsigned char* tc_blob;
  const float* tc_bias_scaled[8];  // GEO: biases of the softplus layers times 100*log2(e)
  const float* tc_w_logit_row_scaled;  // GEO: w_logit_row / (100 log2 e): the logit dot over the scaled activations of the tensor path
  // stage-2 visibility net restructuring (tc path): layer 0 and the skip layer are split into per-point and
  // per-light partial products; vis_aux = {W0 point part, W0 light part, Wskip point part, Wskip light part}
  psn::SimtLayer vis_aux[4];
  const float* w_last_row;  // S2: row 0 of the last layer (dot-product head), [K_pad]
  // stage-2 visibility restructuring: columns of layer 0 / skip layer split into point / light parts
  void* device_blob;  // single allocation that owns every packed array
  size_t blob_bytes;
};
