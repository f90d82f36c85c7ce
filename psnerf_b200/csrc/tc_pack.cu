// Packing of effective fp32 weights into fp16 hi/lo, 128-byte-swizzled UMMA B-operand tiles (see tc_mlp.cuh).
#include "tc_mlp.cuh"
#include "internal.cuh"

namespace psn {

// value(n, k) = transpose ? W[(row0 + k) * ld + col0 + n] : W[(row0 + n) * ld + col0 + k]
__global__ void k_tc_pack(const float* __restrict__ W, int ld, int row0, int col0, int n_valid, int k_valid, int transpose,
                          int n_pad, int nkb, float scale, int k_split, float scale2, unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = nkb * n_pad * 64;
  if (idx >= total) return;
  const int kk = idx & 63, n = (idx >> 6) % n_pad, kb = idx / (64 * n_pad);
  const int k = kb * 64 + kk;
  float v = 0.f;
  if (n < n_valid && k < k_valid) v = (k < k_split ? scale : scale2) * (transpose ? W[(size_t)(row0 + k) * ld + col0 + n] : W[(size_t)(row0 + n) * ld + col0 + k]);
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const size_t tile = (size_t)n_pad * 128;
  const size_t off = (size_t)n * 128 + ((((kk >> 3) ^ (n & 7))) << 4) + (kk & 7) * 2;
  *reinterpret_cast<__half*>(dst + (size_t)(kb * 2 + 0) * tile + off) = hi;
  *reinterpret_cast<__half*>(dst + (size_t)(kb * 2 + 1) * tile + off) = lo;
}

struct TcPlanItem { int step, layer, row0, col0, n_valid, k_valid, transpose; float scale; int k_split; float scale2; };  // columns k >= k_split: scale2

// The softplus(beta=100) stack runs in a SCALED domain: with P = 100 log2(e) the accumulator of every layer is zs = P z, directly the ex2
// argument, and the epilogue hands on  s = max(zs, lg2(1 + 2^zs)) = P h  WITHOUT multiplying by ln2 / 100 (r2: one FMUL less per
// activation in epilogue-bound kernels).  So layer 0 (fed with the unscaled point encoding) is packed times P, the layers fed with s
// are packed as they are, the skip layer (input cat[h, pe] / sqrt2) gets 1 / sqrt2 on its h columns and P / sqrt2 on its pe columns,
// the biases are times P; the fp32 logit row over s_7 is packed times 1 / P and the feature head (no activation) applies 1 / P in its
// epilogue as an FFMA with the bias.
// |s| = P |h| must stay below the fp16 range: h < 454.
#define PSN_SOFTPLUS_PRESCALE 144.26950408889634f
#define PSN_INV_SQRT2_F 0.70710678118654752440f

// Decide which steps exist for this net (0 items => shape unsupported by the tensor path).
static int tc_plan(const psn_mlp* net, TcPlanItem* items) {
  int n = 0;
  const int nl = net->n_layers;
  if (net->kind == PSN_NET_GEO) {
    if (nl != 9 || net->in_dims[0] > tc::PE_K) return 0;  // the kernels keep the point encoding in a [PE_K][128] smem table
    for (int l = 0; l < nl; ++l) {
      if (l > 0 && net->in_dims[l] != 256) return 0;
      if (l < nl - 1 && (net->out_dims[l] > 256 || net->out_dims[l] <= 128)) return 0;
    }
    if (net->out_dims[nl - 1] != 257) return 0;
    const int BIG = 1 << 30;
    for (int l = 0; l < 8; ++l) {
      if (l == 0) items[n++] = {TCG_FWD0, 0, 0, 0, net->out_dims[0], net->in_dims[0], 0, PSN_SOFTPLUS_PRESCALE, BIG, 0.f};
      else if (l == net->desc.skip)
        items[n++] = {TCG_FWD0 + l, l, 0, 0, net->out_dims[l], net->in_dims[l], 0, PSN_INV_SQRT2_F, net->out_dims[l - 1],
                      PSN_SOFTPLUS_PRESCALE * PSN_INV_SQRT2_F};
      else items[n++] = {TCG_FWD0 + l, l, 0, 0, net->out_dims[l], net->in_dims[l], 0, 1.f, BIG, 0.f};
    }
    items[n++] = {TCG_FEAT, 8, 1, 0, 256, 256, 0, 1.f, BIG, 0.f};  // fed with s_7 = P h_7: its epilogue multiplies by 1 / P (scaling the fp16 tiles down would push them into the subnormals)
    for (int l = 7; l >= 1; --l) items[n++] = {TCG_REV_TOP + (7 - l), l, 0, 0, net->in_dims[l], net->out_dims[l], 1, 1.f, BIG, 0.f};
    items[n++] = {TCG_REV0, 0, 0, 0, net->in_dims[0], net->out_dims[0], 1, 1.f, BIG, 0.f};
  } else if (net->kind == PSN_NET_APP) {
    if (nl != 5) return 0;
    const int rest = net->in_dims[0] - 256;
    if (rest < 1 || rest > 64 || net->out_dims[4] > 16) return 0;
    for (int l = 0; l < 4; ++l)
      if (net->out_dims[l] != 256 || (l > 0 && net->in_dims[l] != 256)) return 0;
    if (net->in_dims[4] != 256) return 0;
    items[n++] = {TCA_L0F, 0, 0, rest, 256, 256, 0, 1.f, 1 << 30, 0.f};
    items[n++] = {TCA_L0R, 0, 0, 0, 256, rest, 0, 1.f, 1 << 30, 0.f};
    for (int l = 1; l <= 3; ++l) items[n++] = {TCA_L1 + (l - 1), l, 0, 0, 256, 256, 0, 1.f, 1 << 30, 0.f};
    items[n++] = {TCA_L4, 4, 0, 0, net->out_dims[4], 256, 0, 1.f, 1 << 30, 0.f};
  } else {  // stage-2: only the visibility-net shape (9 layers, width 256, skip after layer 4, scalar output)
    if (nl != 9 || net->desc.skip != 4 || net->out_dims[8] != 1 || net->in_dims[0] > 128 || (net->in_dims[0] & 1)) return 0;
    for (int l = 0; l < 8; ++l)
      if (net->out_dims[l] != 256) return 0;
    for (int l = 1; l < 9; ++l)
      if (net->in_dims[l] != (l == 5 ? 256 + net->in_dims[0] : 256)) return 0;
    for (int l = 1; l <= 4; ++l) items[n++] = {TCV_L1 + (l - 1), l, 0, 0, 256, 256, 0, 1.f, 1 << 30, 0.f};
    items[n++] = {TCV_L5Y, 5, 0, 0, 256, 256, 0, 1.f, 1 << 30, 0.f};
    items[n++] = {TCV_L6, 6, 0, 0, 256, 256, 0, 1.f, 1 << 30, 0.f};
    items[n++] = {TCV_L7, 7, 0, 0, 256, 256, 0, 1.f, 1 << 30, 0.f};
  }
  return n;
}

static void step_shape(const TcPlanItem& it, int* nkb, int* n_pad) {
  *nkb = (it.k_valid + 63) / 64;
  *n_pad = it.n_valid <= 16 ? 16 : it.n_valid <= 48 ? 48 : it.n_valid <= 64 ? 64 : 256;
}

int tc_pack_bytes(const psn_mlp* net) {
  TcPlanItem items[kMaxTcSteps];
  const int n = tc_plan(net, items);
  size_t bytes = 0;
  for (int i = 0; i < n; ++i) {
    int nkb, n_pad;
    step_shape(items[i], &nkb, &n_pad);
    bytes += (size_t)nkb * 2 * n_pad * 128;
  }
  if (net->kind == PSN_NET_S2 && n > 0) {
    const int half = net->in_dims[0] / 2;
    bytes += (size_t)(4 * pad_to(half, 16) * 256 + 256) * sizeof(float) + 1024;
  }
  if (net->kind == PSN_NET_GEO && n > 0) bytes += 9 * 256 * sizeof(float) + 1024;  // pre-scaled biases + the scaled logit row
  return (int)bytes;
}

__global__ void k_pack_kmajor_cols(const float* __restrict__ W, int ld, int col0, int k_valid, int k_pad, int n_valid,
                                   float* __restrict__ dst) {  // dst[k][n] (N_pad = 256) = W[n][col0 + k]
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k_pad * 256) return;
  const int k = idx / 256, n = idx % 256;
  dst[idx] = (k < k_valid && n < n_valid) ? W[(size_t)n * ld + col0 + k] : 0.f;
}
__global__ void k_scale_vec(const float* __restrict__ src, int n, float scale, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = scale * src[i];
}
__global__ void k_copy_row(const float* __restrict__ W, int n, int n_pad, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = i < n ? W[i] : 0.f;
}

int tc_pack_fill(psn_mlp* net, const float* const* W, const float* const* b, char* base, size_t off, cudaStream_t st) {
  (void)b;
  TcPlanItem items[kMaxTcSteps];
  const int n = tc_plan(net, items);
  if (n == 0) return PSN_ERR_SHAPE;
  unsigned char* blob = (unsigned char*)base + off;
  net->tc_blob = blob;
  size_t cur = 0;
  for (int i = 0; i < n; ++i) {
    const TcPlanItem& it = items[i];
    int nkb, n_pad;
    step_shape(it, &nkb, &n_pad);
    net->tc_step[it.step].w_off = (unsigned int)cur;
    net->tc_step[it.step].nkb = (unsigned short)nkb;
    net->tc_step[it.step].n_pad = (unsigned short)n_pad;
    const int total = nkb * n_pad * 64;
    count_launch();
    k_tc_pack<<<(total + 255) / 256, 256, 0, st>>>(W[it.layer], net->in_dims[it.layer], it.row0, it.col0, it.n_valid, it.k_valid,
                                                   it.transpose, n_pad, nkb, it.scale, it.k_split, it.scale2, blob + cur);
    cur += (size_t)nkb * 2 * n_pad * 128;
  }
  if (net->kind == PSN_NET_GEO) {  // biases of the softplus layers, pre-scaled like their weights
    float* f = (float*)(blob + (cur + 255) / 256 * 256);
    for (int l = 0; l < 8; ++l) {
      count_launch();
      k_scale_vec<<<1, 256, 0, st>>>(net->fwd[l].bias, 256, PSN_SOFTPLUS_PRESCALE, f + l * 256);
      net->tc_bias_scaled[l] = f + l * 256;
    }
    count_launch();  // row 0 of the last layer for the fp32 logit dot over s_7 = P h_7
    k_scale_vec<<<1, 256, 0, st>>>(net->w_logit_row, 256, 1.f / PSN_SOFTPLUS_PRESCALE, f + 8 * 256);
    net->tc_w_logit_row_scaled = f + 8 * 256;
  }
  if (net->kind == PSN_NET_S2) {  // fp32 partial-product weights of layer 0 and the skip layer + the scalar head row
    float* f = (float*)(blob + (cur + 255) / 256 * 256);
    const int half = net->in_dims[0] / 2, kp = pad_to(half, 16);
    const int specs[4][2] = {{0, 0}, {0, half}, {5, 256}, {5, 256 + half}};  // {layer, first column}
    for (int a = 0; a < 4; ++a) {
      count_launch();
      k_pack_kmajor_cols<<<(kp * 256 + 255) / 256, 256, 0, st>>>(W[specs[a][0]], net->in_dims[specs[a][0]], specs[a][1], half, kp,
                                                                 256, f);
      net->vis_aux[a].wt = f;
      net->vis_aux[a].bias = (a == 0) ? net->fwd[0].bias : (a == 2) ? net->fwd[5].bias : nullptr;
      net->vis_aux[a].K = half;
      net->vis_aux[a].N = 256;
      net->vis_aux[a].K_pad = kp;
      net->vis_aux[a].N_pad = 256;
      f += (size_t)kp * 256;
    }
    count_launch();
    k_copy_row<<<1, 256, 0, st>>>(W[8], 256, 256, f);
    net->w_last_row = f;
  }
  return PSN_OK;
}

}  // namespace psn
