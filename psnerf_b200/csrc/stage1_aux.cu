// Stage-1 non-MLP kernels: ray generation, sphere test, first-sign-change scan, secant bookkeeping, interval
// sampling plan, alpha compositing, shadow transmittance.  All are HBM-/latency-bound streaming kernels
// (coalesced loads, warp shuffles); the MLP work between them lives in stage1_simt.cu / tc_*.cu.
#include <math_constants.h>

#include "stage1_simt.cuh"
#include "launch.cuh"
#include "internal.cuh"

namespace psn {

// ---- rays ------------------------------------------------------------------------------------------------
// cam = R[9] row-major, origin[3], fx, fy, cx, cy.
struct CamDev { float v[16]; };

__global__ void k_rays_from_pixels(const float* __restrict__ pix, long long N, CamDev cam, int stage2, float* __restrict__ dirs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* R = cam.v;
  const float fx = cam.v[12], fy = cam.v[13], cx = cam.v[14], cy = cam.v[15];
  const float x = __fdiv_rn(__fsub_rn(pix[i * 2 + 0], cx), fx);
  const float y = __fdiv_rn(__fsub_rn(pix[i * 2 + 1], cy), fy);
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) d[r] = R[r * 3 + 0] * x + R[r * 3 + 1] * y + R[r * 3 + 2];
  float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (stage2) n = fmaxf(n, 1e-12f);  // F.normalize eps (rend_util.py:118)
  dirs[i * 3 + 0] = d[0] / n; dirs[i * 3 + 1] = d[1] / n; dirs[i * 3 + 2] = d[2] / n;
}

// get_sphere_intersection far depth, clamped at 0, 0 for misses (rendering.py:576-596)
__global__ void k_sphere_far(const float* __restrict__ dirs, long long N, float ox, float oy, float oz, float r,
                             float* __restrict__ far) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float rc = dirs[i * 3 + 0] * ox + dirs[i * 3 + 1] * oy + dirs[i * 3 + 2] * oz;
  const float cn = sqrtf(ox * ox + oy * oy + oz * oz);
  const float under = __fsub_rn(__fmul_rn(rc, rc), __fsub_rn(__fmul_rn(cn, cn), __fmul_rn(r, r)));
  float f = 0.f;
  if (under > 0.f) f = fmaxf(__fsub_rn(sqrtf(under), rc), 0.f);
  far[i] = f;
}

// ---- ray marching: first sign change + secant state ---------------------------------------------------------
__device__ __forceinline__ float secant_estimate(float f_low, float f_high, float d_low, float d_high) {
  // - f_low * (d_high - d_low) / (f_high - f_low) + d_low   (rendering.py:538,554)
  return __fadd_rn(__fdiv_rn(__fmul_rn(-f_low, __fsub_rn(d_high, d_low)), __fsub_rn(f_high, f_low)), d_low);
}

// One warp per ray.  occ is [N, S] occupancy probabilities; val = occ - tau (rendering.py:457-462).
__global__ void k_march_scan(const float* __restrict__ occ, const float* __restrict__ far, long long N, int S, float near_,
                             float tau, SecantState st, float* __restrict__ depth) {
  const long long ray = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= N) return;
  const float* v = occ + ray * S;
  int idx = -1;
  for (int b = 0; b < S && idx < 0; b += 32) {
    const int s = b + lane;
    const float cur = (s < S) ? v[s] - tau : 0.f;
    float nxt = __shfl_down_sync(0xffffffffu, cur, 1);
    if (lane == 31) nxt = (s + 1 < S) ? v[s + 1] - tau : 0.f;
    const bool flag = (s + 1 < S) && (cur * nxt < 0.f);
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (m) idx = b + __ffs(m) - 1;
  }
  if (lane != 0) return;
  const float v0 = v[0] - tau;
  const bool first_free = v0 < 0.f;
  bool mask = false;
  if (idx >= 0) {
    const float f_low = v[idx] - tau;
    mask = first_free && (f_low < 0.f);
    if (mask) {
      const int i2 = min(idx + 1, S - 1);
      const float f_high = v[i2] - tau;
      const float fr = far[ray];
      const float d_low = lerp_depth(near_, fr, linspace01(idx, S));
      const float d_high = lerp_depth(near_, fr, linspace01(i2, S));
      const int slot = atomicAdd(st.count, 1);
      st.ray[slot] = (int)ray;
      st.d_low[slot] = d_low; st.d_high[slot] = d_high; st.f_low[slot] = f_low; st.f_high[slot] = f_high;
      st.d_pred[slot] = secant_estimate(f_low, f_high, d_low, d_high);
    }
  }
  depth[ray] = first_free ? CUDART_INF_F : 0.f;  // masked rays are overwritten by k_march_finalize
}

// Two-level march, selection: one thread per proposal point.  A point is "visible" to k_march_scan beyond its sign if it can be
// the crossing or its successor; it must be re-evaluated by the full program if its level-1 value is within `margin` of the
// threshold (sign not trusted) or its sign differs from a neighbour's (crossing candidate) - or if that holds for one of its two
// neighbours (a refined neighbour may change sign, which makes this point a crossing candidate).
__device__ __forceinline__ bool refine_seed(const float* __restrict__ v, int j, int S, float tau, float margin) {
  const float c = v[j] - tau;
  bool sel = fabsf(c) < margin;
  if (j > 0) sel |= (c < 0.f) != (v[j - 1] - tau < 0.f);
  if (j + 1 < S) sel |= (c < 0.f) != (v[j + 1] - tau < 0.f);
  return sel;
}
__global__ void k_march_refine_select(const float* __restrict__ occ, const float* __restrict__ far, long long N, int S, float near_,
                                      float tau, float margin, RefineList rl) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * S) return;
  const long long ray = i / S;
  const int s = (int)(i - ray * S);
  const float* v = occ + ray * S;
  bool need = refine_seed(v, s, S, tau, margin);
  if (!need && s > 0) need = refine_seed(v, s - 1, S, tau, margin);
  if (!need && s + 1 < S) need = refine_seed(v, s + 1, S, tau, margin);
  if (!need) return;
  const int slot = atomicAdd(&rl.count[1], 1);
  if (slot < rl.cap) {
    rl.ray[slot] = (int)ray;
    rl.depth[slot] = lerp_depth(near_, far[ray], linspace01(s, S));  // the depth GEN_MARCH gives this point
    rl.pos[slot] = (int)i;
  }
}
__global__ void k_march_refine_close(RefineList rl, int total) {
  const int raw = rl.count[1];
  rl.count[0] = raw < rl.cap ? raw : rl.cap;
  rl.count[2] = raw > rl.cap ? total : 0;
}
__global__ void k_march_refine_scatter(RefineList rl, const float* __restrict__ refined, float* __restrict__ occ) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= rl.count[0]) return;
  occ[rl.pos[slot]] = refined[slot];
}

__global__ void k_secant_update(SecantState st, const float* __restrict__ occ_mid, float tau) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *st.count) return;
  const float f_mid = occ_mid[slot] - tau;
  float d_low = st.d_low[slot], d_high = st.d_high[slot], f_low = st.f_low[slot], f_high = st.f_high[slot];
  const float d_pred = st.d_pred[slot];
  if (f_mid < 0.f) { d_low = d_pred; f_low = f_mid; } else { d_high = d_pred; f_high = f_mid; }
  st.d_low[slot] = d_low; st.d_high[slot] = d_high; st.f_low[slot] = f_low; st.f_high[slot] = f_high;
  st.d_pred[slot] = secant_estimate(f_low, f_high, d_low, d_high);
}

__global__ void k_march_finalize(SecantState st, float* __restrict__ depth) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *st.count) return;
  depth[st.ray[slot]] = st.d_pred[slot];
}

// ---- unisurf sampling plan (rendering.py:88-168) -----------------------------------------------------------
// One warp per ray.  Writes sample_depth[N, S] (S = steps_in + steps_out), mask[N], and appends object rays to
// the surface list (ray id + surface depth) used for the normal pass.

constexpr int PLAN_MAX_S = 256;

__global__ void __launch_bounds__(128)
k_sample_plan(const float* __restrict__ d_i, const float* __restrict__ far, long long N, psn_unisurf_params prm,
              const float* __restrict__ noise, float* __restrict__ sample_depth, uint8_t* __restrict__ mask,
              SurfList sl) {
  __shared__ float sv[4][PLAN_MAX_S];
  __shared__ float so[4][PLAN_MAX_S];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * 4 + w;
  if (ray >= N) return;
  const int S_in = prm.steps_in, S_out = prm.steps_out, S = S_in + S_out;
  const float d = d_i[ray];
  const bool obj = (fabsf(d) != CUDART_INF_F) && !(d != d) && (d != 0.f);
  const float fr = far[ray];
  float* out = sample_depth + ray * S;
  float* v = sv[w];
  float* o = so[w];
  if (!obj) {
    for (int s = lane; s < S; s += 32) o[s] = lerp_depth(prm.near_, fr, linspace01(s, S));
  } else {
    float dnp = __fsub_rn(d, prm.delta), dfp = __fadd_rn(d, prm.delta);
    if (dnp < prm.near_) dnp = prm.near_;
    if (dfp > fr) dfp = fr;
    for (int s = lane; s < S; s += 32) {
      v[s] = (s < S_out) ? lerp_depth(prm.near_, dnp, linspace01(s, S_out))
                         : lerp_depth(dnp, dfp, linspace01(s - S_out, S_in));
    }
    __syncwarp();
    if (S_out > 0) {  // torch.sort(cat[d_binterval, d_interval]) (rendering.py:155): stable rank sort
      for (int s = lane; s < S; s += 32) {
        const float x = v[s];
        int rank = 0;
        for (int t = 0; t < S; ++t) {
          const float y = v[t];
          rank += (y < x) || (y == x && t < s);
        }
        o[rank] = x;
      }
    } else {
      for (int s = lane; s < S; s += 32) o[s] = v[s];
    }
  }
  __syncwarp();
  if (noise) {  // stratified jitter with caller-supplied uniforms (rendering.py:135-140,159-164)
    for (int s = lane; s < S; s += 32) {
      const float lo = (s == 0) ? o[0] : __fmul_rn(0.5f, __fadd_rn(o[s], o[s - 1]));
      const float hi = (s == S - 1) ? o[S - 1] : __fmul_rn(0.5f, __fadd_rn(o[s + 1], o[s]));
      out[s] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), noise[ray * S + s]));
    }
  } else {
    for (int s = lane; s < S; s += 32) out[s] = o[s];
  }
  if (lane == 0) {
    mask[ray] = obj ? 1 : 0;
    if (obj) {
      const int slot = atomicAdd(sl.count, 1);
      sl.ray[slot] = (int)ray;
      sl.depth[slot] = d;
    }
  }
}

// ---- alpha compositing (rendering.py:196-197, 214-216) ----------------------------------------------------
// One warp per ray; exclusive product scan of (1 - a + 1e-6) by shuffles, carried across 32-sample blocks.
__global__ void k_composite(const float* __restrict__ rgb_s, const float* __restrict__ alpha, long long N, int S, int white,
                            float* __restrict__ rgb, float* __restrict__ acc) {
  const long long ray = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= N) return;
  float carry = 1.f, sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f;
  for (int b = 0; b < S; b += 32) {
    const int s = b + lane;
    const float a = (s < S) ? alpha[ray * S + s] : 0.f;
    float t = (s < S) ? __fadd_rn(__fsub_rn(1.f, a), 1e-6f) : 1.f;
    float incl = t;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float u = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl *= u;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    const float wgt = a * (carry * excl);
    if (s < S) {
      const float* c = rgb_s + (ray * S + s) * 3;
      sr += wgt * c[0]; sg += wgt * c[1]; sb += wgt * c[2]; sw += wgt;
    }
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, off);
    sg += __shfl_xor_sync(0xffffffffu, sg, off);
    sb += __shfl_xor_sync(0xffffffffu, sb, off);
    sw += __shfl_xor_sync(0xffffffffu, sw, off);
  }
  if (lane == 0) {
    if (white) { const float bg = 1.f - sw; sr += bg; sg += bg; sb += bg; }
    rgb[ray * 3 + 0] = sr; rgb[ray * 3 + 1] = sg; rgb[ray * 3 + 2] = sb;
    acc[ray] = sw;
  }
}

// ---- shadow transmittance (rendering.py:402-408) ------------------------------------------------------------
// One warp per (light, point) pair of the current light chunk; occ is [pairs, S].
__global__ void k_shadow_composite(const float* __restrict__ occ, const float* __restrict__ surf, const float* __restrict__ lights,
                                   long long Ns, long long pairs, int S, float lnear, float lfar, float box,
                                   float* __restrict__ vis) {
  const long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pair >= pairs) return;
  const long long l = pair / Ns, n = pair - l * Ns;
  const float p0[3] = {surf[n * 3], surf[n * 3 + 1], surf[n * 3 + 2]};
  const float ld[3] = {lights[l * 3], lights[l * 3 + 1], lights[l * 3 + 2]};
  float carry = 1.f, sw = 0.f;
  for (int b = 0; b < S; b += 32) {
    const int s = b + lane;
    float a = 0.f;
    if (s < S) {
      const float dd = lerp_depth(lnear, lfar, linspace01(s, S));
      bool inside = true;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float pc = madd_rn(p0[c], ld[c], dd);
        inside = inside && (pc <= box) && (pc >= -box);
      }
      a = inside ? occ[pair * S + s] : 0.f;
    }
    const float t = (s < S) ? __fadd_rn(__fsub_rn(1.f, a), 1e-6f) : 1.f;
    float incl = t;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float u = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl *= u;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    sw += a * (carry * excl);
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, off);
  if (lane == 0) vis[pair] = 1.f - sw;
}

// ---- box-culled shadow pass (rendering.py:378-408) ------------------------------------------------------------
// The reference evaluates the occupancy MLP at all S steps of every shadow ray and then zeroes alpha outside the +-box cube
// (rendering.py:402-404): those evaluations cannot influence the result.  A shadow ray starts on the surface (inside the cube) and
// the cube is convex, so the in-box steps of a ray are one contiguous span - typically a quarter of the 128 steps at lfar = 3.5.
// k_shadow_plan finds that span per (light, point) pair with the SAME per-step predicate and point arithmetic the composite uses,
// and appends one packed (pair, step) entry per step to a list; the MLP kernels evaluate only the lists (GEN_SHADOW_LIST), writing
// alpha over the entry it came from; k_shadow_composite_list then runs the reference's transmittance product over all S steps with
// alpha = 0 outside the box, exactly as the unculled pass does.  (Contiguity follows from the monotonicity of every rounded
// operation in t -> o + d t; the composite re-evaluates the predicate per step, so correctness does not rest on it.)
// Second cull - dead rays.  vis = 1 - sum_j alpha_j T_j with T_j the transmittance in front of step j, so everything behind step k
// contributes at most T_k.  Rays that enter the object (lights below the local horizon) reach T < 1e-6 within a few steps - about a
// dozen in the soft field of the geometric initialisation, one or two once the occupancy has sharpened - but stay inside the cube
// for another 60-80.  The span is therefore evaluated in two lists: A = its first `lead` (16) steps for every pair, B = the rest,
// only for the pairs with T after list A >= kShadowDeadT (k_shadow_plan_b).  The dropped terms change vis by < 1e-6, an order of
// magnitude below the rounding of the split-operand occupancy itself (1.4e-5 measured on vis) and two below the 1e-4 gate.
// One block = 64 pairs (8 warps x 8 pairs); one atomicAdd per block reserves the block's list range (the list order is therefore
// not reproducible between runs, the values are: every sample is evaluated independently of its tile neighbours).
constexpr int PLAN_PAIRS_PER_BLOCK = 64;
constexpr float kShadowDeadT = 1e-6f;

__device__ __forceinline__ bool shadow_step_inside(const float (&p0)[3], const float (&ld)[3], float dd, float box) {
  bool inside = true;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float pc = madd_rn(p0[c], ld[c], dd);
    inside = inside && (pc <= box) && (pc >= -box);
  }
  return inside;
}

// Reserve `mine[slot]` entries per pair slot of this block in list `which`; returns the block's base offset through s_off / s_base.
__device__ __forceinline__ void plan_reserve(const int* s_cnt, int* s_off, unsigned* s_base, ShadowList sl, int which) {
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < PLAN_PAIRS_PER_BLOCK; ++i) { s_off[i] = run; run += s_cnt[i]; }
    *s_base = run ? atomicAdd(sl.total + which, (unsigned)run) : 0u;
    if (run) atomicAdd(sl.evaluated, (unsigned long long)run);
    if (blockIdx.x == 0) sl.total[4] = 1u;  // marks "the box-culled pass ran" next to the counters (read by the host for statistics)
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
k_shadow_plan(const float* __restrict__ surf, const float* __restrict__ lights, long long Ns, long long pairs, int S, float lnear,
              float lfar, float box, ShadowList sl) {
  __shared__ int s_first[PLAN_PAIRS_PER_BLOCK], s_cnt[PLAN_PAIRS_PER_BLOCK], s_lead[PLAN_PAIRS_PER_BLOCK], s_off[PLAN_PAIRS_PER_BLOCK];
  __shared__ unsigned s_base;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pair0 = (long long)blockIdx.x * PLAN_PAIRS_PER_BLOCK;
  for (int k = 0; k < 8; ++k) {
    const int slot = w * 8 + k;
    const long long pair = pair0 + slot;
    int first = 0, cnt = 0;
    if (pair < pairs) {
      const long long l = pair / Ns, n = pair - l * Ns;
      const float p0[3] = {surf[n * 3], surf[n * 3 + 1], surf[n * 3 + 2]};
      const float ld[3] = {lights[l * 3], lights[l * 3 + 1], lights[l * 3 + 2]};
      int lo = S, hi = -1;
      for (int b = 0; b < S; b += 32) {
        const int s = b + lane;
        const bool in = (s < S) && shadow_step_inside(p0, ld, lerp_depth(lnear, lfar, linspace01(s, S)), box);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (m) {
          if (lo == S) lo = b + __ffs(m) - 1;
          hi = b + 31 - __clz(m);
        }
      }
      if (hi >= 0) { first = lo; cnt = hi - lo + 1; }
    }
    if (lane == 0) { s_first[slot] = first; s_cnt[slot] = cnt; s_lead[slot] = cnt < sl.lead ? cnt : sl.lead; }
  }
  plan_reserve(s_lead, s_off, &s_base, sl, 0);
  for (int k = 0; k < 8; ++k) {
    const int slot = w * 8 + k;
    const long long pair = pair0 + slot;
    if (pair >= pairs) break;
    const unsigned off = s_base + (unsigned)s_off[slot];
    const int first = s_first[slot], cnt = s_cnt[slot], lead = s_lead[slot];
    if (lane == 0) {
      sl.meta[pair] = (unsigned long long)off | ((unsigned long long)first << 32) | ((unsigned long long)cnt << 40);
      sl.off_b[pair] = 0xffffffffu;
    }
    for (int i = lane; i < lead; i += 32) sl.entry[off + i] = ((unsigned)pair << SHADOW_LIST_STEP_BITS) | (unsigned)(first + i);
  }
}

// After the MLP ran on list A: pairs with more in-box steps than `lead` whose transmittance is still >= kShadowDeadT get their
// remaining steps appended to list B.  occ = the alpha values of list A (written over its entries).
__global__ void __launch_bounds__(256) k_shadow_plan_b(const float* __restrict__ occ, long long pairs, ShadowList sl) {
  __shared__ int s_cnt[PLAN_PAIRS_PER_BLOCK], s_off[PLAN_PAIRS_PER_BLOCK];
  __shared__ unsigned s_base;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pair0 = (long long)blockIdx.x * PLAN_PAIRS_PER_BLOCK;
  for (int k = 0; k < 8; ++k) {
    const int slot = w * 8 + k;
    const long long pair = pair0 + slot;
    int rest = 0;
    if (pair < pairs) {
      const unsigned long long meta = sl.meta[pair];
      const int cnt = (int)((meta >> 40) & 0x1ffu);
      if (cnt > sl.lead) {
        const unsigned off = (unsigned)meta;
        float t = lane < sl.lead ? __fadd_rn(__fsub_rn(1.f, occ[off + lane]), 1e-6f) : 1.f;  // lead <= 32
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t *= __shfl_xor_sync(0xffffffffu, t, o);
        if (t >= kShadowDeadT) rest = cnt - sl.lead;
      }
    }
    if (lane == 0) s_cnt[slot] = rest;
  }
  plan_reserve(s_cnt, s_off, &s_base, sl, 1);
  for (int k = 0; k < 8; ++k) {
    const int slot = w * 8 + k;
    const long long pair = pair0 + slot;
    if (pair >= pairs) break;
    const int rest = s_cnt[slot];
    if (rest == 0) continue;
    const unsigned off = s_base + (unsigned)s_off[slot];
    const int first = (int)((sl.meta[pair] >> 32) & 0xffu);
    if (lane == 0) sl.off_b[pair] = off;
    unsigned* dst = sl.entry + sl.cap_a;
    for (int i = lane; i < rest; i += 32) dst[off + i] = ((unsigned)pair << SHADOW_LIST_STEP_BITS) | (unsigned)(first + sl.lead + i);
  }
}

// One warp per pair: the transmittance product of k_shadow_composite with alpha read through the pair's list spans (occ = the entry
// buffer after both MLP launches).  Steps of a dead pair behind list A keep alpha = 0.
__global__ void k_shadow_composite_list(const float* __restrict__ occ, ShadowList sl, const float* __restrict__ surf,
                                        const float* __restrict__ lights, long long Ns, long long pairs, int S, float lnear,
                                        float lfar, float box, float* __restrict__ vis) {
  const long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pair >= pairs) return;
  const unsigned long long meta = sl.meta[pair];
  const unsigned off = (unsigned)meta, off_b = sl.off_b[pair];
  const int first = (int)((meta >> 32) & 0xffu), cnt = (int)((meta >> 40) & 0x1ffu);
  const int lead = cnt < sl.lead ? cnt : sl.lead;
  const int last = off_b != 0xffffffffu ? first + cnt : first + lead;  // steps >= last have alpha = 0: they add nothing to the sum
  const float* occ_b = occ + sl.cap_a;
  const long long l = pair / Ns, n = pair - l * Ns;
  const float p0[3] = {surf[n * 3], surf[n * 3 + 1], surf[n * 3 + 2]};
  const float ld[3] = {lights[l * 3], lights[l * 3 + 1], lights[l * 3 + 2]};
  float carry = 1.f, sw = 0.f;
  for (int b = 0; b < S && b < last; b += 32) {
    const int s = b + lane;
    float a = 0.f;
    if (s < S && s >= first && s < last) {
      if (shadow_step_inside(p0, ld, lerp_depth(lnear, lfar, linspace01(s, S)), box))
        a = s < first + lead ? occ[off + (unsigned)(s - first)] : occ_b[off_b + (unsigned)(s - first - lead)];
    }
    const float t = (s < S) ? __fadd_rn(__fsub_rn(1.f, a), 1e-6f) : 1.f;
    float incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= u;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    sw += a * (carry * excl);
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, o);
  if (lane == 0) vis[pair] = 1.f - sw;
}

// ---- surface normals (rendering.py:208-211) -----------------------------------------------------------------
__global__ void k_scatter_normals(const float* __restrict__ grad, SurfList sl, float* __restrict__ normal) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *sl.count) return;
  const float gx = grad[slot * 3], gy = grad[slot * 3 + 1], gz = grad[slot * 3 + 2];
  const float n = sqrtf(gx * gx + gy * gy + gz * gz) + 1e-5f;
  const long long r = sl.ray[slot];
  normal[r * 3] = gx / n; normal[r * 3 + 1] = gy / n; normal[r * 3 + 2] = gz / n;
}

// ---- host wrappers --------------------------------------------------------------------------------------------
static int g_num_ctas = 0;
int num_ctas() {
  if (g_num_ctas == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      g_num_ctas = sms;
    else
      g_num_ctas = 148;
  }
  return g_num_ctas;
}

int launch_rays(const float* pix, long long N, const float* cam, int stage2, float* dirs, cudaStream_t st) {
  if (N == 0) return PSN_OK;
  CamDev c;
  memcpy(c.v, cam, sizeof(c.v));
  psn::count_launch();
  k_rays_from_pixels<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(pix, N, c, stage2, dirs);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_sphere_far(const float* dirs, long long N, const float* o, float r, float* far, cudaStream_t st) {
  if (N == 0) return PSN_OK;
  psn::count_launch();
  k_sphere_far<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(dirs, N, o[0], o[1], o[2], r, far);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_march_scan(const float* occ, const float* far, long long N, int S, float near_, float tau, SecantState s,
                      float* depth, cudaStream_t st) {
  psn::count_launch();
  k_march_scan<<<(unsigned)((N * 32 + 255) / 256), 256, 0, st>>>(occ, far, N, S, near_, tau, s, depth);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_secant_update(SecantState s, const float* occ_mid, float tau, long long N, cudaStream_t st) {
  psn::count_launch();
  k_secant_update<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(s, occ_mid, tau);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_march_refine_select(const float* occ, const float* far, long long N, int S, float near_, float tau, float margin,
                               RefineList rl, cudaStream_t st) {
  const long long total = N * S;
  psn::count_launch();
  k_march_refine_select<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(occ, far, N, S, near_, tau, margin, rl);
  psn::count_launch();
  k_march_refine_close<<<1, 1, 0, st>>>(rl, (int)total);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_march_refine_scatter(RefineList rl, const float* refined, float* occ, cudaStream_t st) {
  psn::count_launch();
  k_march_refine_scatter<<<(unsigned)((rl.cap + 255) / 256), 256, 0, st>>>(rl, refined, occ);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_march_finalize(SecantState s, float* depth, long long N, cudaStream_t st) {
  psn::count_launch();
  k_march_finalize<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(s, depth);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_sample_plan(const float* d_i, const float* far, long long N, const psn_unisurf_params& prm, const float* noise,
                       float* sample_depth, uint8_t* mask, SurfList sl, cudaStream_t st) {
  PSN_REQUIRE(prm.steps_in + prm.steps_out <= PLAN_MAX_S && prm.steps_in >= 2 && prm.steps_out != 1, PSN_ERR_SHAPE,
              "unisurf: steps_in=%d steps_out=%d unsupported (need 2 <= total <= %d)", prm.steps_in, prm.steps_out,
              PLAN_MAX_S);
  psn::count_launch();
  k_sample_plan<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(d_i, far, N, prm, noise, sample_depth, mask, sl);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_composite(const float* rgb_s, const float* alpha, long long N, int S, int white, float* rgb, float* acc,
                     cudaStream_t st) {
  if (N == 0) return PSN_OK;
  psn::count_launch();
  k_composite<<<(unsigned)((N * 32 + 255) / 256), 256, 0, st>>>(rgb_s, alpha, N, S, white, rgb, acc);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_shadow_composite(const float* occ, const float* surf, const float* lights, long long Ns, long long pairs, int S,
                            float lnear, float lfar, float box, float* vis, cudaStream_t st) {
  if (pairs == 0) return PSN_OK;
  psn::count_launch();
  k_shadow_composite<<<(unsigned)((pairs * 32 + 255) / 256), 256, 0, st>>>(occ, surf, lights, Ns, pairs, S, lnear, lfar,
                                                                            box, vis);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_shadow_plan(const float* surf, const float* lights, long long Ns, long long pairs, int S, float lnear, float lfar,
                       float box, ShadowList sl, cudaStream_t st) {
  psn::count_launch();
  k_shadow_plan<<<(unsigned)((pairs + PLAN_PAIRS_PER_BLOCK - 1) / PLAN_PAIRS_PER_BLOCK), 256, 0, st>>>(surf, lights, Ns, pairs, S,
                                                                                                      lnear, lfar, box, sl);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_shadow_plan_b(const float* occ, long long pairs, ShadowList sl, cudaStream_t st) {
  psn::count_launch();
  k_shadow_plan_b<<<(unsigned)((pairs + PLAN_PAIRS_PER_BLOCK - 1) / PLAN_PAIRS_PER_BLOCK), 256, 0, st>>>(occ, pairs, sl);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_shadow_composite_list(const float* occ, ShadowList sl, const float* surf, const float* lights, long long Ns,
                                 long long pairs, int S, float lnear, float lfar, float box, float* vis, cudaStream_t st) {
  psn::count_launch();
  k_shadow_composite_list<<<(unsigned)((pairs * 32 + 255) / 256), 256, 0, st>>>(occ, sl, surf, lights, Ns, pairs, S, lnear, lfar,
                                                                                 box, vis);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}
int launch_scatter_normals(const float* grad, SurfList sl, float* normal, long long N, cudaStream_t st) {
  psn::count_launch();
  k_scatter_normals<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(grad, sl, normal);
  PSN_CUDA_CHECK(cudaGetLastError());
  return PSN_OK;
}

}  // namespace psn
