// Wide policies on tensor cores: the PPO minibatch gradient (row U, src/ppo.py:220-267) for `--hidden_dim` 128 / 256
// (src/run_ppo.py:36) and any `--num_layers` >= 2, layer by layer.  The fused kernels (update_tc.cu) keep a 128-sample tile of
// every activation in shared memory, which stops fitting beyond 64 units; the shape-generic SIMT kernel (update_generic.cu)
// runs these widths at 27 % of the fp32 FMA peak.  Here every H x H contraction is a tcgen05 GEMM over two-plane bf16 operands
// (v = hi + mid, three products hi*hi + hi*mid + mid*hi, fp32 accumulation in TMEM: the scheme of the fused kernel, within the
// reference's 1e-4), and the activations of a SUB-BATCH of WD_MS samples live in HBM as operand planes:
//
//   wide_first_kernel   gather + first layer + tanh                  -> h_1 planes [2][ms][H]
//   tc GEMM             z = h_l W_l^T                                -> fp32 [ms][H]       (tc_gemm.cu)
//   wide_act_kernel     tanh(z + b_l)                                -> h_{l+1} planes     (num_layers >= 3 only)
//   wide_head_kernel    tanh, output layer, loss (ppo.py:225-264), dz_L = (W_L^T dout)(1 - h_L^2) -> delta planes,
//                       dW_L, db_L, db_{L-1}, log-std gradient, statistics
//   tc_gemm_tn_kernel   dW_l = dz_{l+1}^T h_l, a split-K GEMM over the SAMPLES: both operands are read straight from the
//                       [ms][H] planes as MN-major UMMA operands (TMA boxes of 64 samples x 64 features)
//   tc GEMM             dh_l = dz_{l+1} W_l                          -> fp32 [ms][H]
//   wide_dact_kernel    dz_l = dh_l (1 - h_l^2), db_{l-1}; first layer: dW_0 = dz_1^T x from the re-gathered observations
//
// Sums over samples are kept per ROW of WD_KC samples (one CTA owns a row in every kernel, fixed order: deterministic) in the
// same [net][row][pstride] partial layout the fused kernels write, so grad_reduce_kernel / the data-parallel exchange / Adam
// (update.cu) are shared.  Sub-batches accumulate into the same rows, so the workspace does not depend on the minibatch size.
#include <stdlib.h>

#include "tc.cuh"
#include "update.cuh"

namespace aur {

namespace tc {
int launch_tc_gemm(int64_t M, int64_t N, int64_t K, const void* A, size_t a_plane, const void* B, size_t b_plane, float* C, int ldc,
                   int planes, cudaStream_t stream);
int launch_wide_gemm(int64_t M, int H, const void* A, size_t a_plane, const void* B, size_t b_plane, float* C, int planes, cudaStream_t s,
                     bool mid_mid = false);
}
int gen_pstride(const aur_policy_desc& p);      // update_generic.cu: the partial stride of every non-headline shape

constexpr int WD_P = 2;                          // operand planes
constexpr int WD_MS = 262144;                    // samples per sub-batch
constexpr int WD_KC = 1024;                      // samples per partial row
constexpr int WD_ROWS = WD_MS / WD_KC;
constexpr int WD_THREADS = 256;
constexpr int WD_OUT = 4, WD_IN = 8;             // act_dim <= 4, obs_dim <= 8

bool wide_eligible(const aur_policy_desc& p) {
  return (p.hidden_dim == 128 || p.hidden_dim == 256) && p.num_layers >= 2 && p.num_layers <= 16 && p.obs_dim >= 1 &&
         p.obs_dim <= WD_IN && p.act_dim >= 1 && p.act_dim <= WD_OUT;
}

// ---- workspace layout (bytes, every region 1024-aligned) ----------------------------------------------------------------
struct WideLayout {
  size_t part, wp, hb, dz, z, total;
  size_t plane;                                  // elements between the planes of an activation stack
  size_t wplane;                                 // elements between the planes of a weight stack
};
static WideLayout wide_layout(const aur_policy_desc& p) {
  WideLayout L;
  const size_t H = (size_t)p.hidden_dim, NL = (size_t)p.num_layers;
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  L.plane = (size_t)WD_MS * H;
  L.wplane = H * H;
  size_t o = 0;
  L.part = o; o += up((size_t)2 * WD_ROWS * gen_pstride(p) * 4);
  L.wp = o;   o += up((size_t)2 * (NL - 1) * 2 * WD_P * L.wplane * 2);          // [net][layer][W | W^T][plane][H][H] bf16
  L.hb = o;   o += up((NL - 1) * WD_P * L.plane * 2);                           // h_1 .. h_{L-1}
  L.dz = o;   o += up((size_t)2 * WD_P * L.plane * 2);                          // two delta stacks (ping-pong)
  L.z = o;    o += up(L.plane * 4);
  L.total = o;
  return L;
}
size_t wide_workspace_bytes(const aur_policy_desc& p) { return wide_eligible(p) ? wide_layout(p).total + 1024 : 0; }

// ---- weights: W_l and W_l^T as operand planes ------------------------------------------------------------------------------
__global__ void wide_prep_kernel(const float* __restrict__ wh, int nlayers, int H, __nv_bfloat16* __restrict__ wp, size_t wplane) {
  const size_t hstride = (size_t)H * H + H;
  const long long total = (long long)nlayers * H * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i / ((long long)H * H));
    const int r = (int)(i - (long long)l * H * H);
    const int j = r / H, k = r - j * H;
    const float v = wh[(size_t)l * hstride + r];
    __nv_bfloat16* base = wp + (size_t)l * 2 * WD_P * wplane;
    tc::store_planes(base + (size_t)j * H + k, wplane, WD_P, v);                          // W   [j][k]
    tc::store_planes(base + WD_P * wplane + (size_t)k * H + j, wplane, WD_P, v);          // W^T [k][j]
  }
}

__device__ __forceinline__ long long wide_row(const UpdDev& u, long long gi) {
  return u.idx ? (long long)__ldg(u.idx + gi) : u.idx_offset + gi;
}

// ---- first layer ---------------------------------------------------------------------------------------------------------------
// thread = (sample, 8-feature chunk); W0 [H][obs_dim] | b0 [H] are this net's first parameters.  The grid stride is a multiple
// of the chunks per sample, so a thread keeps ONE chunk for its whole loop and holds that chunk's weights in registers (read
// per item they cost 16 L1 wavefronts per load: 281 us per sub-batch instead of 40).
template <int IN>
__global__ void __launch_bounds__(WD_THREADS, IN <= 4 ? 3 : 2) wide_first_kernel(UpdDev u, const float* __restrict__ W0, int H, long long g0, int ms,
                                                                   __nv_bfloat16* __restrict__ hb, size_t plane) {
  const int LPS = H >> 3, obs_dim = u.obs_dim;
  const float* b0 = W0 + (size_t)H * obs_dim;
  const int c = threadIdx.x % LPS;
  float w[8][IN], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    b[e] = __ldg(b0 + 8 * c + e);
#pragma unroll
    for (int k = 0; k < IN; ++k) w[e][k] = k < obs_dim ? __ldg(W0 + (size_t)(8 * c + e) * obs_dim + k) : 0.0f;
  }
  const int SPP = WD_THREADS / LPS;
  const long long stride = (long long)gridDim.x * SPP;
  // the index -> observation chain of the NEXT sample is in flight while this one is computed and stored
  float xn[IN];
  auto fetch = [&](long long s) {
    if (s < ms) {
      const long long row = wide_row(u, g0 + s);
#pragma unroll
      for (int k = 0; k < IN; ++k) xn[k] = k < obs_dim ? __ldg(u.obs + row * obs_dim + k) : 0.0f;
    }
  };
  long long s = (long long)blockIdx.x * SPP + threadIdx.x / LPS;
  fetch(s);
  for (; s < ms; s += stride) {
    float x[IN];
#pragma unroll
    for (int k = 0; k < IN; ++k) x[k] = xn[k];
    fetch(s + stride);
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float z = b[e];
#pragma unroll
      for (int k = 0; k < IN; ++k) z = fmaf(w[e][k], x[k], z);
      h[e] = tanh_fast(z);
    }
    tc::store8_planes(hb + (size_t)s * H + 8 * c, plane, WD_P, h);
  }
}

// ---- middle layers: h = tanh(z + b) -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WD_THREADS) wide_act_kernel(const float* __restrict__ Z, const float* __restrict__ bias, int H, int ms,
                                                              __nv_bfloat16* __restrict__ hb, size_t plane) {
  const int LPS = H >> 3;
  const long long items = (long long)ms * LPS;
  for (long long it = (long long)blockIdx.x * WD_THREADS + threadIdx.x; it < items; it += (long long)gridDim.x * WD_THREADS) {
    const int c = (int)(it % LPS);
    const float4 za = __ldcs(reinterpret_cast<const float4*>(Z + it * 8)), zb = __ldcs(reinterpret_cast<const float4*>(Z + it * 8 + 4));
    const float* b = bias + 8 * c;                      // the parameters of a net are only 4-byte aligned inside the flat buffer
    float h[8] = {tanh_fast(za.x + __ldg(b)), tanh_fast(za.y + __ldg(b + 1)), tanh_fast(za.z + __ldg(b + 2)), tanh_fast(za.w + __ldg(b + 3)),
                  tanh_fast(zb.x + __ldg(b + 4)), tanh_fast(zb.y + __ldg(b + 5)), tanh_fast(zb.z + __ldg(b + 6)), tanh_fast(zb.w + __ldg(b + 7))};
    tc::store8_planes(hb + it * 8, plane, WD_P, h);
  }
}

// ---- block reductions over the threads that share a feature chunk ------------------------------------------------------------
// v[8] of thread (so, c) -> sum over so (fixed order) of feature f = 8 c + e, returned in thread f < H
__device__ __forceinline__ float chunk_col_sum(const float (&v)[8], float* red, int H) {
  const int tid = threadIdx.x, LPS = H >> 3, SPP = WD_THREADS / LPS;
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) red[tid * 8 + e] = v[e];
  __syncthreads();
  float s = 0.0f;
  if (tid < H) {
    const int c = tid >> 3, e = tid & 7;
    for (int so = 0; so < SPP; ++so) s += red[(so * LPS + c) * 8 + e];
  }
  return s;
}
__device__ __forceinline__ float block_sum_wide(float v, float* sred) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x == 0)
    for (int w = 0; w < WD_THREADS / 32; ++w) t += sred[w];
  return t;   // valid in thread 0
}
__device__ __forceinline__ void put(float* p, float v, int beta) { *p = beta ? *p + v : v; }

struct WideHead {
  const float* Z;            // [ms][H] pre-activation of the last hidden layer, without its bias
  const float* bias;         // b_{L-1}
  const float* WL;           // [OUT][H] | bL [OUT]
  const float* logstd;       // actor_logstd (continuous actor) or nullptr
  int H, OUT, ms, beta, pstride;
  long long g0;
  __nv_bfloat16* dz;         // delta planes out
  size_t plane;
  float* part;               // this net's rows
  size_t oBh, oWL, oBL, oLS; // offsets inside a row
};

// KIND: 0 actor (Categorical), 1 actor (Normal), 2 critic.  One CTA per row of WD_KC samples; the LPS = H / 8 lanes of a sample
// sit side by side in a warp, the head is a butterfly sum over them (every lane ends up with the same bits), and all of them
// evaluate the loss redundantly.
template <int KIND, int OUT>
__global__ void __launch_bounds__(WD_THREADS, (KIND == 2 || OUT == 1) ? 3 : 2) wide_head_kernel(UpdDev u, WideHead a) {
  constexpr bool ACTOR = KIND != 2;
  __shared__ __align__(16) float red[WD_THREADS * 8];
  __shared__ __align__(16) float sWL[4 * 256];
  __shared__ float sred[8];
  const int H = a.H, LPS = H >> 3, SPP = WD_THREADS / LPS;
  const int tid = threadIdx.x, c = tid % LPS, so = tid / LPS;
  const int s_begin = blockIdx.x * WD_KC, s_end = min(a.ms, s_begin + WD_KC);
  float* prow = a.part + (size_t)blockIdx.x * a.pstride;
  for (int i = tid; i < OUT * H; i += WD_THREADS) sWL[i] = __ldg(a.WL + i);
  float bL[OUT], sd[OUT], ls[OUT];
#pragma unroll
  for (int k = 0; k < OUT; ++k) {
    bL[k] = k < OUT ? __ldg(a.WL + (size_t)OUT * H + k) : 0.0f;
    const float l = (KIND == 1 && k < OUT) ? __ldg(a.logstd + k) : 0.0f;
    sd[k] = expf(l); ls[k] = logf(sd[k]);                 // torch Normal: log(exp(logstd))
  }
  float adv_mean = 0.0f, adv_den = 1.0f;
  if (ACTOR && u.norm_adv) {           // ONE thread reads the moments (data-parallel: waits for the peers' flags), the CTA shares them
    if (tid == 0) { adv_norm_consts(u, adv_mean, adv_den); sred[0] = adv_mean; sred[1] = adv_den; }
    __syncthreads();
    adv_mean = sred[0]; adv_den = sred[1];
  }
  float bias[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) bias[e] = __ldg(a.bias + 8 * c + e);
  float dWL[OUT][8], dbh[8], dbL[OUT], gls[OUT];
#pragma unroll
  for (int k = 0; k < OUT; ++k) {
    dbL[k] = 0.0f; gls[k] = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) dWL[k][e] = 0.0f;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) dbh[e] = 0.0f;
  float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f, st4 = 0.f;
  __syncthreads();

  // software pipeline: the pre-activations and the row index of the NEXT sample are requested before this one is processed
  // (an iteration is a chain of dependent global loads: index -> per-sample scalars -> loss; 16 warps per SM do not hide it)
  float4 za_n = make_float4(0.f, 0.f, 0.f, 0.f), zb_n = za_n;
  long long row_n = 0;
  auto prefetch = [&](int s) {
    if (s < s_end) {
      za_n = __ldcs(reinterpret_cast<const float4*>(a.Z + (size_t)s * H + 8 * c));
      zb_n = __ldcs(reinterpret_cast<const float4*>(a.Z + (size_t)s * H + 8 * c + 4));
      row_n = wide_row(u, a.g0 + s);
    }
  };
  prefetch(s_begin + so);
  for (int base = s_begin; base < s_end; base += SPP) {          // uniform trip count: the shuffles below need every lane
    const int s = base + so;
    const bool valid = s < s_end;
    const float4 za = za_n, zb = zb_n;
    const long long row = row_n;
    // this sample's scalars (their address needs the row index, which arrived an iteration ago)
    float e0 = 0.0f, e1 = 0.0f, eact[OUT];
#pragma unroll
    for (int k = 0; k < OUT; ++k) eact[k] = 0.0f;
    if (valid) {
      if (ACTOR) {
        e0 = __ldg(u.logprobs + row); e1 = __ldg(u.advantages + row);
        if (KIND == 0) eact[0] = __ldg(u.actions + row);
        else {
#pragma unroll
          for (int k = 0; k < OUT; ++k) if (k < OUT) eact[k] = __ldg(u.actions + row * OUT + k);
        }
      } else {
        e0 = __ldg(u.returns + row); e1 = __ldg(u.values + row);
      }
    }
    prefetch(s + SPP);
    float h[8];
    h[0] = tanh_fast(za.x + bias[0]); h[1] = tanh_fast(za.y + bias[1]); h[2] = tanh_fast(za.z + bias[2]); h[3] = tanh_fast(za.w + bias[3]);
    h[4] = tanh_fast(zb.x + bias[4]); h[5] = tanh_fast(zb.y + bias[5]); h[6] = tanh_fast(zb.z + bias[6]); h[7] = tanh_fast(zb.w + bias[7]);
    float out[OUT], dout[OUT];
#pragma unroll
    for (int k = 0; k < OUT; ++k) {
      float o = 0.0f;
      if (k < OUT) {
        const float4 wa = lds4(sWL + k * H + 8 * c), wb = lds4(sWL + k * H + 8 * c + 4);
        o = fmaf(wa.x, h[0], o); o = fmaf(wa.y, h[1], o); o = fmaf(wa.z, h[2], o); o = fmaf(wa.w, h[3], o);
        o = fmaf(wb.x, h[4], o); o = fmaf(wb.y, h[5], o); o = fmaf(wb.z, h[6], o); o = fmaf(wb.w, h[7], o);
        for (int off = LPS >> 1; off > 0; off >>= 1) o += __shfl_xor_sync(0xffffffffu, o, off);
      }
      out[k] = o + bL[k];
      dout[k] = 0.0f;
    }
    if (valid) {
      if (ACTOR) {
        const float oldlp = e0, adv = e1;
        float newlogp, entropy;
        float dlp[OUT], dH[OUT];
        if (KIND == 0) {
          float m = out[0];
#pragma unroll
          for (int k = 1; k < OUT; ++k) if (k < OUT) m = fmaxf(m, out[k]);
          float se = 0.0f;
#pragma unroll
          for (int k = 0; k < OUT; ++k) if (k < OUT) se += expf(out[k] - m);
          const float lse = m + logf(se);
          const int act = (int)eact[0];
          float lp[OUT], pr[OUT];
          entropy = 0.0f; newlogp = 0.0f;
#pragma unroll
          for (int k = 0; k < OUT; ++k) {
            lp[k] = out[k] - lse;
            pr[k] = k < OUT ? expf(lp[k]) : 0.0f;
            if (k < OUT) entropy -= pr[k] * lp[k];
            if (k == act) newlogp = lp[k];
          }
#pragma unroll
          for (int k = 0; k < OUT; ++k) {
            dlp[k] = k < OUT ? (k == act ? 1.0f : 0.0f) - pr[k] : 0.0f;
            dH[k] = k < OUT ? -pr[k] * (lp[k] + entropy) : 0.0f;
          }
        } else {
          const float LOG_SQRT_2PI = 0.91893853320467267f;
          newlogp = 0.0f; entropy = 0.0f;
#pragma unroll
          for (int k = 0; k < OUT; ++k) {
            dlp[k] = 0.0f; dH[k] = 0.0f;
            if (k < OUT) {
              const float d = eact[k] - out[k], var = sd[k] * sd[k];
              newlogp += -(d * d) / (2.0f * var) - ls[k] - LOG_SQRT_2PI;
              entropy += 0.5f + LOG_SQRT_2PI + ls[k];
              dlp[k] = d / var;
            }
          }
        }
        const float logr = newlogp - oldlp;
        const float ratio = expf(logr);
        const float advn = u.norm_adv ? (adv - adv_mean) / adv_den : adv;
        const float l1 = -advn * ratio;
        const float l2 = -advn * fminf(fmaxf(ratio, u.clip_lo), u.clip_hi);
        const float w1 = l1 > l2 ? 1.0f : (l1 == l2 ? 0.5f : 0.0f);
        const float inr = (ratio >= u.clip_lo && ratio <= u.clip_hi) ? 1.0f : 0.0f;
        const float g_logp = -advn * (w1 + (1.0f - w1) * inr) * ratio * u.inv_m;
        const float g_H = -u.ent_c * u.inv_m;
#pragma unroll
        for (int k = 0; k < OUT; ++k) dout[k] = g_logp * dlp[k] + g_H * dH[k];
        if (c == 0) {
          if (KIND == 1) {
#pragma unroll
            for (int k = 0; k < OUT; ++k)
              if (k < OUT) {
                const float d = out[k] - eact[k];
                gls[k] += g_logp * (d * d / (sd[k] * sd[k]) - 1.0f) + g_H;
              }
          }
          st0 += fmaxf(l1, l2); st1 += entropy; st2 += -logr; st3 += (ratio - 1.0f) - logr;
          st4 += fabsf(ratio - 1.0f) > u.clip ? 1.0f : 0.0f;
        }
      } else {
        const float R = e0, vold = e1, v = out[0];
        float l;
        if (u.clip_vloss) {
          const float du = v - R, vu = du * du;
          const float d = v - vold, vc = vold + fminf(fmaxf(d, -u.clip), u.clip);
          const float dc = vc - R, lc = dc * dc;
          const float w1 = vu > lc ? 1.0f : (vu == lc ? 0.5f : 0.0f);
          const float inr = (d >= -u.clip && d <= u.clip) ? 1.0f : 0.0f;
          dout[0] = (w1 * du + (1.0f - w1) * dc * inr) * u.vf_c * u.inv_m;
          l = 0.5f * fmaxf(vu, lc);
        } else {
          const float d = v - vold;                      // reference quirk ppo.py:261: b_values, not b_returns
          dout[0] = d * u.vf_c * u.inv_m;
          l = 0.5f * d * d;
        }
        if (c == 0) st0 += l;
      }
      // ---- delta of the last hidden layer, output-layer gradients
      float dzv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) dzv[e] = 0.0f;
#pragma unroll
      for (int k = 0; k < OUT; ++k) {
        if (k < OUT) {
          const float4 wa = lds4(sWL + k * H + 8 * c), wb = lds4(sWL + k * H + 8 * c + 4);
          dzv[0] = fmaf(wa.x, dout[k], dzv[0]); dzv[1] = fmaf(wa.y, dout[k], dzv[1]);
          dzv[2] = fmaf(wa.z, dout[k], dzv[2]); dzv[3] = fmaf(wa.w, dout[k], dzv[3]);
          dzv[4] = fmaf(wb.x, dout[k], dzv[4]); dzv[5] = fmaf(wb.y, dout[k], dzv[5]);
          dzv[6] = fmaf(wb.z, dout[k], dzv[6]); dzv[7] = fmaf(wb.w, dout[k], dzv[7]);
#pragma unroll
          for (int e = 0; e < 8; ++e) dWL[k][e] = fmaf(dout[k], h[e], dWL[k][e]);
          if (c == 0) dbL[k] += dout[k];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { dzv[e] *= fmaf(-h[e], h[e], 1.0f); dbh[e] += dzv[e]; }
      tc::store8_planes(a.dz + (size_t)s * H + 8 * c, a.plane, WD_P, dzv);
    }
  }

  // ---- row sums
  float v = chunk_col_sum(dbh, red, H);
  if (tid < H) put(prow + a.oBh + tid, v, a.beta);
#pragma unroll
  for (int k = 0; k < OUT; ++k) {
    if (k < OUT) {
      v = chunk_col_sum(dWL[k], red, H);
      if (tid < H) put(prow + a.oWL + (size_t)k * H + tid, v, a.beta);
    }
  }
#pragma unroll
  for (int k = 0; k < OUT; ++k) {
    if (k < OUT) {
      v = block_sum_wide(dbL[k], sred);
      if (tid == 0) put(prow + a.oBL + k, v, a.beta);
      if (KIND == 1) {
        v = block_sum_wide(gls[k], sred);
        if (tid == 0) put(prow + a.oLS + k, v, a.beta);
      }
    }
  }
  float* stat = prow + (a.pstride - AUR_NUM_STATS);
  v = block_sum_wide(st0, sred); if (tid == 0) put(stat + (ACTOR ? AUR_STAT_POLICY_LOSS : AUR_STAT_VALUE_LOSS), v, a.beta);
  if (ACTOR) {
    v = block_sum_wide(st1, sred); if (tid == 0) put(stat + AUR_STAT_ENTROPY, v, a.beta);
    v = block_sum_wide(st2, sred); if (tid == 0) put(stat + AUR_STAT_OLD_APPROX_KL, v, a.beta);
    v = block_sum_wide(st3, sred); if (tid == 0) put(stat + AUR_STAT_APPROX_KL, v, a.beta);
    v = block_sum_wide(st4, sred); if (tid == 0) put(stat + AUR_STAT_CLIPFRAC, v, a.beta);
  }
}

// ---- dz_l = dh_l (1 - h_l^2), bias gradient; FIRST: the first layer's weight gradient from the re-gathered observations ----
struct WideDact {
  const float* Z;                 // dh_l [ms][H]
  const __nv_bfloat16* hb;        // h_l planes
  __nv_bfloat16* dz;              // delta planes out (not FIRST)
  size_t plane;
  int H, ms, beta, pstride;
  long long g0;
  float* part;
  size_t oB, oW0;
};
__device__ __forceinline__ void add_bf16x8(const uint4& q, float (&h)[8]) {
  const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { h[2 * j] += __uint_as_float(w[j] << 16); h[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u); }
}
template <bool FIRST, int IN>
__global__ void __launch_bounds__(WD_THREADS, IN <= 4 ? 3 : 2) wide_dact_kernel(UpdDev u, WideDact a) {
  __shared__ __align__(16) float red[WD_THREADS * 8];
  const int H = a.H, LPS = H >> 3, SPP = WD_THREADS / LPS, obs_dim = u.obs_dim;
  const int tid = threadIdx.x, c = tid % LPS, so = tid / LPS;
  const int s_begin = blockIdx.x * WD_KC, s_end = min(a.ms, s_begin + WD_KC);
  float* prow = a.part + (size_t)blockIdx.x * a.pstride;
  float dbh[8];
  float dW0[FIRST ? IN : 1][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    dbh[e] = 0.0f;
#pragma unroll
    for (int k = 0; k < (FIRST ? IN : 1); ++k) dW0[k][e] = 0.0f;
  }
  for (int s = s_begin + so; s < s_end; s += SPP) {
    const size_t o = (size_t)s * H + 8 * c;
    const float4 da = __ldcs(reinterpret_cast<const float4*>(a.Z + o)), db = __ldcs(reinterpret_cast<const float4*>(a.Z + o + 4));
    float h[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = WD_P - 1; p >= 0; --p) add_bf16x8(__ldg(reinterpret_cast<const uint4*>(a.hb + (size_t)p * a.plane + o)), h);
    float dzv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) { dzv[e] *= fmaf(-h[e], h[e], 1.0f); dbh[e] += dzv[e]; }
    if (FIRST) {
      const long long row = wide_row(u, a.g0 + s);
#pragma unroll
      for (int k = 0; k < IN; ++k) {
        if (k < obs_dim) {
          const float x = __ldg(u.obs + row * obs_dim + k);
#pragma unroll
          for (int e = 0; e < 8; ++e) dW0[FIRST ? k : 0][e] = fmaf(dzv[e], x, dW0[FIRST ? k : 0][e]);
        }
      }
    } else {
      tc::store8_planes(a.dz + o, a.plane, WD_P, dzv);
    }
  }
  float v = chunk_col_sum(dbh, red, H);
  if (tid < H) put(prow + a.oB + tid, v, a.beta);
  if (FIRST) {
#pragma unroll
    for (int k = 0; k < IN; ++k) {
      if (k < obs_dim) {
        v = chunk_col_sum(dW0[FIRST ? k : 0], red, H);
        if (tid < H) put(prow + a.oW0 + (size_t)tid * obs_dim + k, v, a.beta);
      }
    }
  }
}

// ---- split-K GEMM over the samples: C_r[j][i] (+)= sum_{s in row r} A[s][j] B[s][i] ---------------------------------------------
// A = delta planes, B = activation planes, both [ms][H] row-major (features contiguous): MN-major UMMA operands, a stage holds
// for BOTH planes of both operands the boxes of 64 samples x 64 features (8 KB, 128-B rows, 128-B swizzle; 8 sample rows form one
// 1024-B swizzle atom, the 64-feature groups of an operand are 8192 B apart), and the three plane products are issued from it.
// grid = (H / 128, rows): CTA (mb, r) owns C_r[128 mb .. +128][0 .. N); the M blocks of a row are neighbours in launch order, so
// the B tiles they both read come from DRAM once (with the row index fastest the 256-wide kernel read 872 MB for 536 MB of
// operands).  Warp 0: TMA, warp 1: MMA, warps 2..5: epilogue.
constexpr int TN_BK = 64;
template <int N>
struct TnCfg {
  static constexpr int A_BYTES = WD_P * 2 * 8192, B_BYTES = WD_P * (N / 64) * 8192, STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = N == 128 ? 3 : 2;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE + 1024 + 256;
};
template <int N>
__global__ void __launch_bounds__(192, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                  size_t c_row_stride, int ldc, int K, int beta) {
  using namespace tc;
  using Cfg = TnCfg<N>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tmem_full = empty + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.y * WD_KC, k1 = min(K, k0 + WD_KC);
  const int nkb = k1 > k0 ? (k1 - k0 + TN_BK - 1) / TN_BK : 0;
  const int j0 = blockIdx.x * 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, N);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  if (nkb > 0) {
    if (warp == 0 && lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (i / Cfg::STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE);
        unsigned char* sa = smem + s * Cfg::STAGE;
        unsigned char* sb = sa + Cfg::A_BYTES;
        const int kk = k0 + i * TN_BK;
        for (int p = 0; p < WD_P; ++p) {
          for (int g = 0; g < 2; ++g) tma_load_3d(sa + (p * 2 + g) * 8192, &tmA, j0 + 64 * g, kk, p, &full[s]);
          for (int g = 0; g < N / 64; ++g) tma_load_3d(sb + (p * (N / 64) + g) * 8192, &tmB, 64 * g, kk, p, &full[s]);
        }
      }
    } else if (warp == 1 && lane == 0) {
      constexpr uint32_t idesc = instr_desc(FMT_BF16, 128, N, 1, 1);           // both operands MN-major
      for (int i = 0; i < nkb; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (i / Cfg::STAGES) & 1u;
        mbar_wait(&full[s], ph);
        fence_after_sync();
        unsigned char* sa = smem + s * Cfg::STAGE;
        unsigned char* sb = sa + Cfg::A_BYTES;
#pragma unroll
        for (int t = 0; t < 3; ++t) {                                          // hi*hi, hi*mid, mid*hi
          const uint64_t ad = smem_desc_mn_sw128(sa + term_plane_a(t) * 2 * 8192, 8192, 1024);
          const uint64_t bd = smem_desc_mn_sw128(sb + term_plane_b(t) * (N / 64) * 8192, 8192, 1024);
#pragma unroll
          for (int k = 0; k < TN_BK / 16; ++k)                                  // 16 sample rows per MMA = 2048 B = 128 x 16 B
            mma_f16(tmem_d, ad + (uint64_t)(128 * k), bd + (uint64_t)(128 * k), idesc, (i | t | k) != 0);
        }
        mma_commit(&empty[s]);
      }
      mma_commit(tmem_full);
    }
  }
  if (warp >= 2) {
    const int q = warp & 3;                                                    // TMEM lane quarter this warp may read
    const int j = j0 + 32 * q + lane;
    float* dst = C + (size_t)blockIdx.y * c_row_stride + (size_t)j * ldc;
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      fence_after_sync();
#pragma unroll 1
      for (int cc = 0; cc < N; cc += 32) {
        float v[32];
        tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)cc, v);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {                                      // rows are 16-byte aligned (H, pstride multiples of 4)
          float4* d4 = reinterpret_cast<float4*>(dst + cc + i);
          float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          if (beta) { const float4 c4 = *d4; o.x += c4.x; o.y += c4.y; o.z += c4.z; o.w += c4.w; }
          *d4 = o;
        }
      }
    } else if (!beta) {
      for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, N);
}

template <int N>
static int launch_gemm_tn(const __nv_bfloat16* A, const __nv_bfloat16* B, size_t plane, int ms, float* C, size_t c_row_stride, int rows,
                          int beta, cudaStream_t s) {
  using Cfg = TnCfg<N>;
  CUtensorMap tmA, tmB;
  const uint64_t d[3] = {(uint64_t)N, (uint64_t)ms, (uint64_t)WD_P};
  const uint64_t st[2] = {(uint64_t)N * 2, (uint64_t)plane * 2};
  const uint32_t box[3] = {64, TN_BK, 1};
  int rc;
  if ((rc = tc::make_tensor_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, d, st, box))) return rc;
  if ((rc = tc::make_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, B, d, st, box))) return rc;
  static DeviceOnce attr;
  if (attr.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(tc_gemm_tn_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr.done();
  }
  tc_gemm_tn_kernel<N><<<dim3(N / 128, (unsigned)rows), 192, Cfg::SMEM, s>>>(tmA, tmB, C, c_row_stride, N, ms, beta);
  AUR_LAUNCH_OK("tc_gemm_tn_kernel");
  return 0;
}

// ---- persistent GEMM for H = 128: C[M][128] = A[M][128] B[128][128]^T over two-plane operands --------------------------------
// The per-tile GEMM (tc_gemm.cu: one CTA per 128 x 128 tile) spends most of a tile on its own prologue and epilogue when K is
// only 128 (93 us for 268 MB of traffic).  Here a CTA keeps both planes of B (the layer's weights, 64 KB) in shared memory and
// walks row tiles: a 64 KB stage holds both planes of a 128-row A tile (all of K), the three plane products go into one of two
// TMEM accumulators, and the four epilogue warps store tile t while the tensor core works on t + 1.
constexpr int SG_ATOM = 128 * 128;                       // 128 rows x 64 K values (128-B rows)
constexpr int SG_STAGE = WD_P * 2 * SG_ATOM, SG_STAGES = 2;
constexpr int SG_PATCH = 32 * 33 * 4;                    // per epilogue warp: a 32 x 32 fp32 block, rows padded to 33 words
constexpr size_t SG_SMEM = (size_t)SG_STAGE * (1 + SG_STAGES) + 4 * SG_PATCH + 1024 + 256;
__global__ void __launch_bounds__(192, 1)
wide_gemm128_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C, int M) {
  using namespace tc;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sB = smem;
  unsigned char* sA = smem + SG_STAGE;
  float* patches = reinterpret_cast<float*>(sA + SG_STAGES * SG_STAGE);
  uint64_t* bfull = reinterpret_cast<uint64_t*>(sA + SG_STAGES * SG_STAGE + 4 * SG_PATCH);
  uint64_t* full = bfull + 1;
  uint64_t* empty = full + SG_STAGES;
  uint64_t* tfull = empty + SG_STAGES;      // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(bfull, 1);
    for (int s = 0; s < SG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_fence_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(bfull, SG_STAGE);
    for (int p = 0; p < WD_P; ++p)
      for (int g = 0; g < 2; ++g) tma_load_3d(sB + (p * 2 + g) * SG_ATOM, &tmB, 64 * g, 0, p, bfull);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % SG_STAGES;
      mbar_wait(&empty[s], ((it / SG_STAGES) & 1u) ^ 1u);
      mbar_arrive_expect_tx(&full[s], SG_STAGE);
      for (int p = 0; p < WD_P; ++p)
        for (int g = 0; g < 2; ++g) tma_load_3d(sA + s * SG_STAGE + (p * 2 + g) * SG_ATOM, &tmA, 64 * g, tile * 128, p, &full[s]);
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = instr_desc(FMT_BF16, 128, 128, 0, 0);
    mbar_wait(bfull, 0);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % SG_STAGES;
      const uint32_t acc = it & 1u;
      mbar_wait(&tempty[acc], ((it >> 1) & 1u) ^ 1u);
      mbar_wait(&full[s], (it / SG_STAGES) & 1u);
      fence_after_sync();
#pragma unroll
      for (int t = 0; t < 3; ++t) {                                            // hi*hi, hi*mid, mid*hi
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint64_t ad = smem_desc_k_sw128(sA + s * SG_STAGE + (term_plane_a(t) * 2 + g) * SG_ATOM);
          const uint64_t bd = smem_desc_k_sw128(sB + (term_plane_b(t) * 2 + g) * SG_ATOM);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16(tmem_d + acc * 128, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (t | g | k) != 0);
        }
      }
      mma_commit(&empty[s]);
      mma_commit(&tfull[acc]);
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      mbar_wait(&tfull[acc], (it >> 1) & 1u);
      fence_after_sync();
      // a thread holds 32 consecutive columns of ITS row; stored like that a warp instruction would touch 32 rows (32 LSU
      // wavefronts).  The warp transposes each 32 x 32 block through a padded patch instead: every store is one 128-B row segment.
      const long long row0 = (long long)tile * 128 + 32 * q;
      float* patch = patches + (warp - 2) * (SG_PATCH / 4);
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        float v[32];
        tmem_ld32(tmem_d + acc * 128 + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * g), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) patch[lane * 33 + i] = v[i];
        __syncwarp();
#pragma unroll 8
        for (int r = 0; r < 32; ++r)
          if (row0 + r < M) C[(row0 + r) * 128 + 32 * g + lane] = patch[r * 33 + lane];
        __syncwarp();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, 256);
}

// Z[ms][H] = A planes [ms][H] x B planes [H][H]^T
static int wide_gemm(int H, const __nv_bfloat16* A, size_t a_plane, const __nv_bfloat16* B, size_t b_plane, int ms, float* Z, cudaStream_t s) {
  if (H != 128) return tc::launch_wide_gemm(ms, H, A, a_plane, B, b_plane, Z, WD_P, s);
  CUtensorMap tmA, tmB;
  const uint64_t dA[3] = {128, (uint64_t)ms, (uint64_t)WD_P}, dB[3] = {128, 128, (uint64_t)WD_P};
  const uint64_t stA[2] = {256, (uint64_t)a_plane * 2}, stB[2] = {256, (uint64_t)b_plane * 2};
  const uint32_t box[3] = {64, 128, 1};
  int rc;
  if ((rc = tc::make_tensor_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, dA, stA, box))) return rc;
  if ((rc = tc::make_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, B, dB, stB, box))) return rc;
  static DeviceOnce attr;
  if (attr.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(wide_gemm128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM));
    attr.done();
  }
  const int ntiles = (ms + 127) / 128;
  const int grid = ntiles < sm_count() ? ntiles : sm_count();
  wide_gemm128_kernel<<<grid, 192, SG_SMEM, s>>>(tmA, tmB, Z, ms);
  AUR_LAUNCH_OK("wide_gemm128_kernel");
  return 0;
}

// 1 (default): hidden 128 / 256 run this path; 0: the shape-generic SIMT kernel (cross-check); AUR_UPDATE_WIDE=0|1
static int g_wide = -1;
int update_wide_enabled() {
  if (g_wide < 0) {
    const char* e = getenv("AUR_UPDATE_WIDE");
    g_wide = (e && e[0] == '0') ? 0 : 1;
  }
  return g_wide;
}
void set_update_wide(int on) { g_wide = on ? 1 : 0; }

// Launches the whole layer-wise gradient pass; partial rows end up at *part_out as [2][*gx_out][gen_pstride].
int launch_ppo_grad_wide(const UpdDev& d, const aur_policy_desc& p, float* ws, int* gx_out, float** part_out, cudaStream_t s) {
  const WideLayout L = wide_layout(p);
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 1023) / 1024 * 1024);
  const int H = p.hidden_dim, NL = p.num_layers, obs_dim = p.obs_dim;
  const int pstride = gen_pstride(p);
  const long long m = d.m_local;
  const int gx = (int)((m < WD_MS ? m : (long long)WD_MS) + WD_KC - 1) / WD_KC;
  float* part = reinterpret_cast<float*>(base + L.part);
  *gx_out = gx;
  *part_out = part;
  if (m <= 0) return 0;
  __nv_bfloat16* wp_all = reinterpret_cast<__nv_bfloat16*>(base + L.wp);
  __nv_bfloat16* hb_all = reinterpret_cast<__nv_bfloat16*>(base + L.hb);
  __nv_bfloat16* dzb[2] = {reinterpret_cast<__nv_bfloat16*>(base + L.dz), reinterpret_cast<__nv_bfloat16*>(base + L.dz) + WD_P * L.plane};
  float* Z = reinterpret_cast<float*>(base + L.z);
  const int64_t nA = net_param_count(obs_dim, H, NL, p.act_dim), nC = net_param_count(obs_dim, H, NL, 1);
  const size_t hstride = (size_t)H * H + H;
  const int ew_grid = sm_count() * 8;
  int rc;

  for (int net = 0; net < 2; ++net) {
    const int OUT = net == 0 ? p.act_dim : 1;
    const float* np = d.params + (net == 0 ? 0 : nA);
    float* npart = part + (size_t)net * gx * pstride;
    const size_t oB0 = (size_t)H * obs_dim, oWh = oB0 + H, oWL = oWh + (size_t)(NL - 1) * hstride, oBL = oWL + (size_t)OUT * H,
                 oLS = oBL + OUT;
    __nv_bfloat16* wp = wp_all + (size_t)net * (NL - 1) * 2 * WD_P * L.wplane;
    wide_prep_kernel<<<(unsigned)(((size_t)(NL - 1) * H * H + 255) / 256), 256, 0, s>>>(np + oWh, NL - 1, H, wp, L.wplane);
    AUR_LAUNCH_OK("wide_prep_kernel");
    auto Wp = [&](int l) { return wp + (size_t)(l - 1) * 2 * WD_P * L.wplane; };                   // W_l planes, l = 1 .. NL-1
    auto WTp = [&](int l) { return Wp(l) + WD_P * L.wplane; };
    auto hb = [&](int l) { return hb_all + (size_t)(l - 1) * WD_P * L.plane; };                    // h_l planes, l = 1 .. NL-1

    for (long long g0 = 0; g0 < m; g0 += WD_MS) {
      const int ms = (int)((m - g0) < WD_MS ? (m - g0) : WD_MS);
      const int rows = (ms + WD_KC - 1) / WD_KC;
      const int beta = g0 > 0 ? 1 : 0;
      if (obs_dim <= 4) wide_first_kernel<4><<<ew_grid, WD_THREADS, 0, s>>>(d, np, H, g0, ms, hb(1), L.plane);
      else wide_first_kernel<8><<<ew_grid, WD_THREADS, 0, s>>>(d, np, H, g0, ms, hb(1), L.plane);
      AUR_LAUNCH_OK("wide_first_kernel");
      for (int l = 1; l <= NL - 1; ++l) {
        if ((rc = wide_gemm(H, hb(l), L.plane, Wp(l), L.wplane, ms, Z, s))) return rc;
        if (l < NL - 1) {
          wide_act_kernel<<<ew_grid, WD_THREADS, 0, s>>>(Z, np + oWh + (size_t)(l - 1) * hstride + (size_t)H * H, H, ms, hb(l + 1), L.plane);
          AUR_LAUNCH_OK("wide_act_kernel");
        }
      }
      WideHead h;
      h.Z = Z; h.bias = np + oWh + (size_t)(NL - 2) * hstride + (size_t)H * H; h.WL = np + oWL;
      h.logstd = (net == 0 && p.continuous) ? d.params + nA + nC : nullptr;
      h.H = H; h.OUT = OUT; h.ms = ms; h.beta = beta; h.pstride = pstride; h.g0 = g0;
      h.dz = dzb[0]; h.plane = L.plane; h.part = npart;
      h.oBh = oWh + (size_t)(NL - 2) * hstride + (size_t)H * H; h.oWL = oWL; h.oBL = oBL; h.oLS = oLS;
      {
        void (*hk)(UpdDev, WideHead) = wide_head_kernel<2, 1>;
        if (net == 0) {
          if (p.continuous) hk = OUT == 1 ? wide_head_kernel<1, 1> : OUT == 2 ? wide_head_kernel<1, 2> : OUT == 3 ? wide_head_kernel<1, 3> : wide_head_kernel<1, 4>;
          else hk = OUT == 1 ? wide_head_kernel<0, 1> : OUT == 2 ? wide_head_kernel<0, 2> : OUT == 3 ? wide_head_kernel<0, 3> : wide_head_kernel<0, 4>;
        }
        hk<<<rows, WD_THREADS, 0, s>>>(d, h);
      }
      AUR_LAUNCH_OK("wide_head_kernel");
      int cur = 0;
      for (int l = NL - 1; l >= 1; --l) {
        float* cw = npart + oWh + (size_t)(l - 1) * hstride;                                       // dW_l inside row 0
        if (H == 128) rc = launch_gemm_tn<128>(dzb[cur], hb(l), L.plane, ms, cw, pstride, rows, beta, s);
        else rc = launch_gemm_tn<256>(dzb[cur], hb(l), L.plane, ms, cw, pstride, rows, beta, s);
        if (rc) return rc;
        if ((rc = wide_gemm(H, dzb[cur], L.plane, WTp(l), L.wplane, ms, Z, s))) return rc;
        WideDact a;
        a.Z = Z; a.hb = hb(l); a.dz = dzb[cur ^ 1]; a.plane = L.plane; a.H = H; a.ms = ms; a.beta = beta; a.pstride = pstride;
        a.g0 = g0; a.part = npart; a.oW0 = 0;
        if (l > 1) {
          a.oB = oWh + (size_t)(l - 2) * hstride + (size_t)H * H;
          wide_dact_kernel<false, 4><<<rows, WD_THREADS, 0, s>>>(d, a);
          cur ^= 1;
        } else {
          a.oB = oB0;
          if (obs_dim <= 4) wide_dact_kernel<true, 4><<<rows, WD_THREADS, 0, s>>>(d, a);
          else wide_dact_kernel<true, 8><<<rows, WD_THREADS, 0, s>>>(d, a);
        }
        AUR_LAUNCH_OK("wide_dact_kernel");
      }
    }
  }
  return 0;
}

}  // namespace aur
