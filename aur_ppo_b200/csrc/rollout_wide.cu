// Rollout for the policy widths whose second-layer weights do not fit shared memory beside an env tile (`--hidden_dim 256`, or
// 128 with more than two layers; src/run_ppo.py:36,38): the actor runs LAYER BY LAYER over all N envs each step, its H x H
// contractions as tcgen05 GEMMs over two-plane bf16 operands with all four plane products (the precision of rollout_tc_kernel: the
// log-probs are compared with the reference at 2e-5; the value pass uses three planes / six products, values are compared at 2e-6), and the env / sampling / bookkeeping part is the tensor-core rollout kernel itself in
// its external-actor mode (rollout_tc_kernel<ENV, 0>: same fp64 physics, Philox streams, autoreset and episode log, so the
// replay tests do not change).  Per step: rows_first_kernel (first layer -> h planes), tc_gemm, [rows_act_kernel per middle
// layer], rows_head_kernel (tanh + output layer -> logits [N][4]), rollout step.  The T*N critic values are the same three
// kernels over the observation buffer afterwards, in sub-batches of 262,144 rows.  Replaces the runtime-width SIMT kernel for
// these shapes (694 ms per 65536 x 128 rollout at 256 units).  Scratch: library-owned (aur_rollout has no workspace argument).
#include "envs.cuh"
#include "tc.cuh"

namespace aur {

namespace tc {
int launch_wide_gemm(int64_t M, int H, const void* A, size_t a_plane, const void* B, size_t b_plane, float* C, int planes, cudaStream_t s,
                     bool mid_mid = false);
}
int launch_rollout_step_ext(const RolloutDev& d, int env_kind, cudaStream_t s);

constexpr int RW_P = 3;                      // planes of the value pass (values are compared at 2e-6) and of the scratch layout
constexpr int RW_P_ACTOR = 2;                // the per-step actor: two planes, four products (hi*hi, hi*mid, mid*hi, mid*mid) like rollout_tc_kernel
constexpr int RW_MS = 262144;
constexpr int RW_THREADS = 256;

bool rollout_wide_eligible(const aur_policy_desc& p, int env_kind) {
  return (p.hidden_dim == 128 || p.hidden_dim == 256) && p.num_layers >= 2 && p.num_layers <= 16 && p.obs_dim <= POL_IN_PAD &&
         p.act_dim <= POL_OUT_MAX && (env_kind == AUR_ENV_CARTPOLE || env_kind == AUR_ENV_PENDULUM || env_kind == AUR_ENV_MOUNTAINCAR);
}

// hidden-layer weights W_1 .. W_{NL-1} of one net -> operand planes [layer][plane][H][H]
__global__ void rows_prep_kernel(const float* __restrict__ wh, int nlayers, int H, __nv_bfloat16* __restrict__ wp, int P) {
  const size_t hstride = (size_t)H * H + H, wplane = (size_t)H * H;
  const long long total = (long long)nlayers * H * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i / (long long)wplane);
    const size_t r = (size_t)(i - (long long)l * (long long)wplane);
    tc::store_planes(wp + (size_t)l * P * wplane + r, wplane, P, wh[(size_t)l * hstride + r]);
  }
}

// first layer over contiguous rows: h = tanh(W0 x + b0) -> planes [P][.][H]; thread = (row, 8-feature chunk), the chunk's weights
// in registers for the whole loop
__global__ void __launch_bounds__(RW_THREADS) rows_first_kernel(const float* __restrict__ X, int obs_dim, const float* __restrict__ W0, int H,
                                                                long long M, __nv_bfloat16* __restrict__ hb, size_t plane, int P) {
  const int LPS = H >> 3, SPP = RW_THREADS / LPS;
  const float* b0 = W0 + (size_t)H * obs_dim;
  const int c = threadIdx.x % LPS;
  float w[8][POL_IN_PAD], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    b[e] = __ldg(b0 + 8 * c + e);
#pragma unroll
    for (int k = 0; k < POL_IN_PAD; ++k) w[e][k] = k < obs_dim ? __ldg(W0 + (size_t)(8 * c + e) * obs_dim + k) : 0.0f;
  }
  for (long long r = (long long)blockIdx.x * SPP + threadIdx.x / LPS; r < M; r += (long long)gridDim.x * SPP) {
    float x[POL_IN_PAD];
#pragma unroll
    for (int k = 0; k < POL_IN_PAD; ++k) x[k] = k < obs_dim ? __ldg(X + r * obs_dim + k) : 0.0f;
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float z = b[e];
#pragma unroll
      for (int k = 0; k < POL_IN_PAD; ++k) z = fmaf(w[e][k], x[k], z);
      h[e] = tanh_fast(z);
    }
    tc::store8_planes(hb + (size_t)r * H + 8 * c, plane, P, h);
  }
}

// middle layers: h = tanh(z + b) -> planes
__global__ void __launch_bounds__(RW_THREADS) rows_act_kernel(const float* __restrict__ Z, const float* __restrict__ bias, int H, long long M,
                                                              __nv_bfloat16* __restrict__ hb, size_t plane, int P) {
  const int LPS = H >> 3;
  const long long items = M * LPS;
  for (long long it = (long long)blockIdx.x * RW_THREADS + threadIdx.x; it < items; it += (long long)gridDim.x * RW_THREADS) {
    const int c = (int)(it % LPS);
    const float4 za = __ldcs(reinterpret_cast<const float4*>(Z + it * 8)), zb = __ldcs(reinterpret_cast<const float4*>(Z + it * 8 + 4));
    const float* b = bias + 8 * c;
    float h[8] = {tanh_fast(za.x + __ldg(b)), tanh_fast(za.y + __ldg(b + 1)), tanh_fast(za.z + __ldg(b + 2)), tanh_fast(za.w + __ldg(b + 3)),
                  tanh_fast(zb.x + __ldg(b + 4)), tanh_fast(zb.y + __ldg(b + 5)), tanh_fast(zb.z + __ldg(b + 6)), tanh_fast(zb.w + __ldg(b + 7))};
    tc::store8_planes(hb + it * 8, plane, P, h);
  }
}

// last hidden activation + output layer: out[row * ldo + k] = bL[k] + sum_f WL[k][f] tanh(Z[row][f] + bias[f]) for k < OUT (zeros up
// to ldo).  The H / 8 lanes of a row sit side by side in a warp; the output is a butterfly sum over them.
__global__ void __launch_bounds__(RW_THREADS) rows_head_kernel(const float* __restrict__ Z, const float* __restrict__ bias, const float* __restrict__ WL,
                                                               int H, int OUT, long long M, float* __restrict__ out, int ldo) {
  const int LPS = H >> 3, SPP = RW_THREADS / LPS;
  const int c = threadIdx.x % LPS, so = threadIdx.x / LPS;
  float b[8], w[POL_OUT_MAX][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    b[e] = __ldg(bias + 8 * c + e);
#pragma unroll
    for (int k = 0; k < POL_OUT_MAX; ++k) w[k][e] = k < OUT ? __ldg(WL + (size_t)k * H + 8 * c + e) : 0.0f;
  }
  const float* bL = WL + (size_t)OUT * H;
  for (long long base = (long long)blockIdx.x * SPP; base < M; base += (long long)gridDim.x * SPP) {     // uniform trip count per warp
    const long long r = base + so;
    const bool valid = r < M;
    float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;
    if (valid) {
      za = __ldcs(reinterpret_cast<const float4*>(Z + (size_t)r * H + 8 * c));
      zb = __ldcs(reinterpret_cast<const float4*>(Z + (size_t)r * H + 8 * c + 4));
    }
    const float h[8] = {tanh_fast(za.x + b[0]), tanh_fast(za.y + b[1]), tanh_fast(za.z + b[2]), tanh_fast(za.w + b[3]),
                        tanh_fast(zb.x + b[4]), tanh_fast(zb.y + b[5]), tanh_fast(zb.z + b[6]), tanh_fast(zb.w + b[7])};
    float o[POL_OUT_MAX];
#pragma unroll
    for (int k = 0; k < POL_OUT_MAX; ++k) {
      float v = 0.0f;
#pragma unroll
      for (int e = 0; e < 8; ++e) v = fmaf(w[k][e], h[e], v);
      for (int off = LPS >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      o[k] = k < OUT ? v + __ldg(bL + k) : 0.0f;
    }
    if (valid && c == 0) {
      if (ldo == 4) *reinterpret_cast<float4*>(out + r * 4) = make_float4(o[0], o[1], o[2], o[3]);
      else {
#pragma unroll
        for (int k = 0; k < POL_OUT_MAX; ++k) if (k < ldo) out[r * ldo + k] = o[k];
      }
    }
  }
}

struct RwScratch {
  __nv_bfloat16 *wp_actor, *wp_critic, *hb[2];
  float *Z, *logits;
  size_t plane;
};
static size_t rw_layout(const aur_policy_desc& p, long long N, unsigned char* base, RwScratch* out) {
  const size_t H = (size_t)p.hidden_dim, NL = (size_t)p.num_layers;
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  const size_t rows = (size_t)RW_MS;                      // per-step actor passes are chunked to the same sub-batch size
  size_t o = 0;
  const size_t wp = up((NL - 1) * RW_P * H * H * 2), hb = up(RW_P * rows * H * 2), z = up(rows * H * 4), lg = up((size_t)N * 16);
  if (out) {
    out->wp_actor = reinterpret_cast<__nv_bfloat16*>(base + o);
    out->wp_critic = reinterpret_cast<__nv_bfloat16*>(base + o + wp);
    out->hb[0] = reinterpret_cast<__nv_bfloat16*>(base + o + 2 * wp);
    out->hb[1] = reinterpret_cast<__nv_bfloat16*>(base + o + 2 * wp + hb);
    out->Z = reinterpret_cast<float*>(base + o + 2 * wp + 2 * hb);
    out->logits = reinterpret_cast<float*>(base + o + 2 * wp + 2 * hb + z);
    out->plane = rows * H;
  }
  o += 2 * wp + 2 * hb + z + lg;
  return o;
}

// out[M][ldo] = MLP(X[M][obs_dim]) for one net of the flat parameter buffer, hidden layers on tensor cores
static int rows_forward(const float* net, const __nv_bfloat16* wp, int P, int obs_dim, int H, int NL, int OUT, const float* X, long long M,
                        float* out, int ldo, const RwScratch& sc, cudaStream_t s) {
  const size_t hstride = (size_t)H * H + H, wplane = (size_t)H * H;
  const float* wh = net + (size_t)H * obs_dim + H;                    // W_1 | b_1 | ...
  const float* WL = wh + (size_t)(NL - 1) * hstride;
  const int grid = sm_count() * 8;
  for (long long m0 = 0; m0 < M; m0 += RW_MS) {
    const long long ms = (M - m0) < RW_MS ? (M - m0) : RW_MS;
    int cur = 0;
    rows_first_kernel<<<grid, RW_THREADS, 0, s>>>(X + m0 * obs_dim, obs_dim, net, H, ms, sc.hb[0], sc.plane, P);
    AUR_LAUNCH_OK("rows_first_kernel");
    for (int l = 1; l <= NL - 1; ++l) {
      int rc = tc::launch_wide_gemm(ms, H, sc.hb[cur], sc.plane, wp + (size_t)(l - 1) * P * wplane, wplane, sc.Z, P, s, P == 2);
      if (rc) return rc;
      if (l < NL - 1) {
        rows_act_kernel<<<grid, RW_THREADS, 0, s>>>(sc.Z, wh + (size_t)(l - 1) * hstride + wplane, H, ms, sc.hb[cur ^ 1], sc.plane, P);
        AUR_LAUNCH_OK("rows_act_kernel");
        cur ^= 1;
      }
    }
    rows_head_kernel<<<grid, RW_THREADS, 0, s>>>(sc.Z, wh + (size_t)(NL - 2) * hstride + wplane, WL, H, OUT, ms, out + m0 * ldo, ldo);
    AUR_LAUNCH_OK("rows_head_kernel");
  }
  return 0;
}

int launch_rollout_wide(const RolloutDev& d, const aur_policy_desc& p, int env_kind, cudaStream_t s) {
  const int H = p.hidden_dim, NL = p.num_layers, obs_dim = p.obs_dim;
  const size_t bytes = rw_layout(p, d.N, nullptr, nullptr) + 1024;
  unsigned char* raw = static_cast<unsigned char*>(stream_scratch(s, bytes));
  if (!raw) return AUR_ERR_ARG;
  RwScratch sc;
  rw_layout(p, d.N, reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) / 1024 * 1024), &sc);
  const int64_t nA = net_param_count(obs_dim, H, NL, p.act_dim);
  const float* actor = d.params;
  const float* critic = d.params + nA;
  const unsigned pgrid = (unsigned)(((size_t)(NL - 1) * H * H + 255) / 256);
  rows_prep_kernel<<<pgrid, 256, 0, s>>>(actor + (size_t)H * obs_dim + H, NL - 1, H, sc.wp_actor, RW_P_ACTOR);
  AUR_LAUNCH_OK("rows_prep_kernel");
  rows_prep_kernel<<<pgrid, 256, 0, s>>>(critic + (size_t)H * obs_dim + H, NL - 1, H, sc.wp_critic, RW_P);
  AUR_LAUNCH_OK("rows_prep_kernel");
  int rc;
  RolloutDev step = d;
  step.T = 1;
  step.ext_logits = sc.logits;
  for (int t = 0; t < d.T; ++t) {
    if ((rc = rows_forward(actor, sc.wp_actor, RW_P_ACTOR, obs_dim, H, NL, p.act_dim, d.next_obs, d.N, sc.logits, 4, sc, s))) return rc;
    step.t0 = t;
    if ((rc = launch_rollout_step_ext(step, env_kind, s))) return rc;
  }
  if ((rc = rows_forward(critic, sc.wp_critic, RW_P, obs_dim, H, NL, 1, d.obs_buf, (long long)d.T * d.N, d.val_buf, 1, sc, s))) return rc;
  if (d.next_value && (rc = rows_forward(critic, sc.wp_critic, RW_P, obs_dim, H, NL, 1, d.next_obs, d.N, d.next_value, 1, sc, s))) return rc;
  return 0;
}

}  // namespace aur
