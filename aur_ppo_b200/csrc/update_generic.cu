// Shape-generic PPO minibatch gradient kernel (row U for every `--hidden_dim` / `--num_layers` the reference CLI accepts,
// src/run_ppo.py:36,38, and for observation / action widths up to 8).  Same contract, statistics and packed output as
// ppo_grad_kernel / ppo_grad_tc_kernel (update.cu, update_tc.cu), which stay the kernels for the 64 x 2 headline shape;
// this one is the CUDA path for everything else, so that no flag combination ends in a fallback.
//
//   stage_params_kernel        flat parameters -> 16-byte aligned, row-padded copy (float4 weight loads everywhere)
//   ppo_grad_generic_kernel    gather -> forward (all layers kept in shared memory) -> loss -> backward -> gradients
//
// A CTA (256 threads) trains one net (blockIdx.y) on tiles of S = 32 / 64 / 128 samples.  Activations are feature-major
// [H][S + 4] in shared memory; a dense layer is a register-tiled product with lane = sample and warp = 4 output rows, the
// weights arriving as warp-uniform read-only loads; backward-data overwrites each activation buffer with its own delta;
// weight gradients are contractions over the tile's samples whose per-CTA sums live in a private slab of global memory
// (L2-resident for small nets), owned element-wise by one thread, so the accumulation order is fixed (deterministic).
#include "update.cuh"

namespace aur {

constexpr int GEN_THREADS = 256;
constexpr int GEN_IO = 8;                 // obs_dim, act_dim <= 8

__host__ __device__ inline int gen_in_pad(int obs_dim) { return (obs_dim + 3) & ~3; }
// staged (aligned) floats of one net: W0 [H][IP] | b0 [H] | (NL-1) x (W [H][H] | b [H]) | Wout [OUT][H] | bout [8]
__host__ __device__ inline int64_t gen_staged_floats(int obs_dim, int H, int NL, int out) {
  return (int64_t)H * gen_in_pad(obs_dim) + H + (int64_t)(NL - 1) * ((int64_t)H * H + H) + (int64_t)out * H + 8;
}

struct GenDev {
  UpdDev u;
  int H, NL, S, LD, pstride;
  const float* staged;      // [actor staged | critic staged]
  float* part;              // [2][gridDim.x][pstride]
};

__global__ void stage_params_kernel(const float* __restrict__ params, int obs_dim, int H, int NL, int act_dim,
                                    float* __restrict__ staged) {
  const int IP = gen_in_pad(obs_dim);
  const int64_t nA = net_param_count(obs_dim, H, NL, act_dim);
  const int64_t sA = gen_staged_floats(obs_dim, H, NL, act_dim), sC = gen_staged_floats(obs_dim, H, NL, 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sA + sC; i += (int64_t)gridDim.x * blockDim.x) {
    const bool critic = i >= sA;
    const int64_t e = critic ? i - sA : i;
    const int out = critic ? 1 : act_dim;
    const float* g = params + (critic ? nA : 0);
    float v;
    const int64_t first = (int64_t)H * IP;
    if (e < first) {
      const int j = (int)(e / IP), c = (int)(e - (int64_t)j * IP);
      v = c < obs_dim ? g[(int64_t)j * obs_dim + c] : 0.0f;
    } else {
      const int64_t r = e - first;                                   // the rest keeps the flat order
      const int64_t rest = H + (int64_t)(NL - 1) * ((int64_t)H * H + H) + (int64_t)out * H + out;
      v = r < rest ? g[(int64_t)H * obs_dim + r] : 0.0f;
    }
    staged[i] = v;
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// out[j][s] = tanh(b[j] + sum_k W[j][k] in[k][s]),  W [H][in_pad] staged, in / out feature-major with row stride LD
template <int TS>
__device__ __forceinline__ void dense_tanh_fwd(const float* __restrict__ W, const float* __restrict__ b, int in_pad,
                                               const float* __restrict__ in, float* __restrict__ out, int H, int LD) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int jb = warp * 4; jb < H; jb += 32) {
    float acc[4][TS];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int t = 0; t < TS; ++t) acc[jj][t] = 0.0f;
#pragma unroll 2
    for (int k = 0; k < in_pad; k += 4) {
      float x[4][TS];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int t = 0; t < TS; ++t) x[kk][t] = in[(k + kk) * LD + lane + 32 * t];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 w = ldg4(W + (size_t)(jb + jj) * in_pad + k);
#pragma unroll
        for (int t = 0; t < TS; ++t)
          acc[jj][t] = fmaf(w.w, x[3][t], fmaf(w.z, x[2][t], fmaf(w.y, x[1][t], fmaf(w.x, x[0][t], acc[jj][t]))));
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float bj = __ldg(b + jb + jj);
#pragma unroll
      for (int t = 0; t < TS; ++t) out[(jb + jj) * LD + lane + 32 * t] = tanh_fast(acc[jj][t] + bj);
    }
  }
}

// h[k][s] <- (sum_j W[j][k] delta[j][s]) * (1 - h[k][s]^2)   (delta of the layer below, in place over its activation)
template <int TS>
__device__ __forceinline__ void dense_bwd_data(const float* __restrict__ W, const float* __restrict__ delta, int R,
                                               float* __restrict__ h, int H, int LD) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int kb = warp * 4; kb < H; kb += 32) {
    float acc[4][TS];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int t = 0; t < TS; ++t) acc[kk][t] = 0.0f;
#pragma unroll 4
    for (int j = 0; j < R; ++j) {
      const float4 w = ldg4(W + (size_t)j * H + kb);
#pragma unroll
      for (int t = 0; t < TS; ++t) {
        const float d = delta[j * LD + lane + 32 * t];
        acc[0][t] = fmaf(w.x, d, acc[0][t]); acc[1][t] = fmaf(w.y, d, acc[1][t]);
        acc[2][t] = fmaf(w.z, d, acc[2][t]); acc[3][t] = fmaf(w.w, d, acc[3][t]);
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int t = 0; t < TS; ++t) {
        float* p = h + (kb + kk) * LD + lane + 32 * t;
        const float hv = *p;
        *p = acc[kk][t] * fmaf(-hv, hv, 1.0f);
      }
  }
}

// part[j][i] += sum_s delta[j][s] h[i][s]  for a R x C matrix (row stride C in the flat parameter order), 64 x 64 blocks of
// 4 x 4 register tiles: thread (tj = tid / 16, ti = tid % 16) owns rows tj + 16 jj and columns ti + 16 ii of every block.
__device__ __forceinline__ void wgrad_big(const float* __restrict__ delta, const float* __restrict__ h, int R, int C, int S,
                                          int LD, float* __restrict__ part) {
  const int tj = threadIdx.x >> 4, ti = threadIdx.x & 15;
  for (int jb = 0; jb < R; jb += 64)
    for (int ib = 0; ib < C; ib += 64) {
      float2 acc[4][4];
      const float *ap[4], *bp[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = jb + tj + 16 * q, i = ib + ti + 16 * q;
        ap[q] = delta + (j < R ? j : R - 1) * LD;
        bp[q] = h + (i < C ? i : C - 1) * LD;
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[q][r] = make_float2(0.f, 0.f);
      }
#pragma unroll 2
      for (int s4 = 0; s4 < S; s4 += 4) {
        float4 av[4], bv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { av[q] = lds4(ap[q] + s4); bv[q] = lds4(bp[q] + s4); }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const float2 s = __ffma2_rn(make_float2(av[jj].x, av[jj].y), make_float2(bv[ii].x, bv[ii].y), acc[jj][ii]);
            acc[jj][ii] = __ffma2_rn(make_float2(av[jj].z, av[jj].w), make_float2(bv[ii].z, bv[ii].w), s);
          }
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int j = jb + tj + 16 * jj, i = ib + ti + 16 * ii;
          if (j < R && i < C) part[(size_t)j * C + i] += acc[jj][ii].x + acc[jj][ii].y;
        }
    }
}

// the same contraction for thin matrices (first layer: H x obs_dim, output layer: OUT x H): one element per thread and pass
__device__ __forceinline__ void wgrad_small(const float* __restrict__ delta, const float* __restrict__ h, int R, int C, int S,
                                            int LD, float* __restrict__ part) {
  for (int e = threadIdx.x; e < R * C; e += GEN_THREADS) {
    const int r = e / C, c = e - r * C;
    const float *ap = delta + r * LD, *bp = h + c * LD;
    float2 s2 = make_float2(0.f, 0.f);
    for (int s4 = 0; s4 < S; s4 += 4) {
      const float4 av = lds4(ap + s4), bv = lds4(bp + s4);
      s2 = __ffma2_rn(make_float2(av.x, av.y), make_float2(bv.x, bv.y), s2);
      s2 = __ffma2_rn(make_float2(av.z, av.w), make_float2(bv.z, bv.w), s2);
    }
    part[e] += s2.x + s2.y;
  }
}

__device__ __forceinline__ void bgrad(const float* __restrict__ delta, int R, int S, int LD, float* __restrict__ part) {
  for (int r = threadIdx.x; r < R; r += GEN_THREADS) {
    const float* p = delta + r * LD;
    float s = 0.0f;
    for (int s4 = 0; s4 < S; s4 += 4) { const float4 v = lds4(p + s4); s += (v.x + v.y) + (v.z + v.w); }
    part[r] += s;
  }
}

__device__ __forceinline__ float block_sum_gen(float v, float* sred) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x == 0)
    for (int w = 0; w < GEN_THREADS / 32; ++w) t += sred[w];
  return t;   // valid in thread 0
}

// KIND: 0 actor (Categorical), 1 actor (Normal), 2 critic.  TS = S / 32.
template <int KIND, int TS>
__device__ void generic_net(const GenDev& g, float* smem) {
  constexpr bool ACTOR = KIND != 2;
  const UpdDev& a = g.u;
  const int tid = threadIdx.x;
  const int H = g.H, NL = g.NL, S = g.S, LD = g.LD;
  const int obs_dim = a.obs_dim, IP = gen_in_pad(obs_dim);
  const int OUT = ACTOR ? a.act_dim : 1;
  const int Q = GEN_THREADS / S;                           // output-layer partial sums per sample

  // ---- shared memory: x [8][LD] | dout [8][LD] | partial heads [Q*8][LD] | activations [NL][H][LD]
  float* sX = smem;
  float* sDout = sX + GEN_IO * LD;
  float* sRed = sDout + GEN_IO * LD;
  float* sAct = sRed + Q * GEN_IO * LD;                    // layer l (1..NL) at sAct + (l-1) * H * LD

  // ---- staged weights / flat gradient offsets of this net
  const int64_t nA = net_param_count(obs_dim, H, NL, a.act_dim);
  const float* W0 = g.staged + (ACTOR ? 0 : gen_staged_floats(obs_dim, H, NL, a.act_dim));
  const float* b0 = W0 + (size_t)H * IP;
  const float* Wh = b0 + H;                                // layer l (1..NL-1): Wh + (l-1) * (H*H + H)
  const size_t hstride = (size_t)H * H + H;
  const float* Wout = Wh + (size_t)(NL - 1) * hstride;
  const float* bout = Wout + (size_t)OUT * H;
  float* part = g.part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * g.pstride;
  const size_t oW0 = 0, oB0 = (size_t)H * obs_dim, oWh = oB0 + H, oWout = oWh + (size_t)(NL - 1) * hstride,
               oBout = oWout + (size_t)OUT * H, oLS = oBout + OUT;
  for (int i = tid; i < g.pstride; i += GEN_THREADS) part[i] = 0.0f;

  float adv_mean = 0.0f, adv_den = 1.0f;
  if (ACTOR && a.norm_adv) adv_norm_consts(a, adv_mean, adv_den);
  float sd[GEN_IO], ls[GEN_IO];
  if (KIND == 1) {
#pragma unroll
    for (int k = 0; k < GEN_IO; ++k) {
      const float l = k < OUT ? a.params[nA + net_param_count(obs_dim, H, NL, 1) + k] : 0.0f;
      sd[k] = expf(l); ls[k] = logf(sd[k]);                // torch Normal: log(exp(logstd))
    }
  }
  float logstd_g[GEN_IO];
#pragma unroll
  for (int k = 0; k < GEN_IO; ++k) logstd_g[k] = 0.0f;
  float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f, st4 = 0.f;
  __syncthreads();

  const long long ntiles = (a.m_local + S - 1) / S;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // ---- gather: thread s < S owns sample tile * S + s
    const long long i = tile * S + tid;
    const bool valid = tid < S && i < a.m_local;
    long long row = 0;
    if (valid) row = a.idx ? (long long)__ldg(a.idx + i) : a.idx_offset + i;
    if (tid < S) {
#pragma unroll
      for (int c = 0; c < GEN_IO; ++c)
        if (c < IP) sX[c * LD + tid] = (valid && c < obs_dim) ? __ldg(a.obs + row * obs_dim + c) : 0.0f;
    }
    __syncthreads();

    // ---- forward
    dense_tanh_fwd<TS>(W0, b0, IP, sX, sAct, H, LD);
    __syncthreads();
    for (int l = 1; l < NL; ++l) {
      const float* W = Wh + (size_t)(l - 1) * hstride;
      dense_tanh_fwd<TS>(W, W + (size_t)H * H, H, sAct + (size_t)(l - 1) * H * LD, sAct + (size_t)l * H * LD, H, LD);
      __syncthreads();
    }
    float* hL = sAct + (size_t)(NL - 1) * H * LD;           // last hidden activation
    {
      // output layer: thread (s, q) sums its slice of the hidden units; a warp shares q, so the weights are uniform loads
      const int s = tid % S, q = tid / S;
      const int per = (H + Q - 1) / Q;
      const int j0 = q * per, j1 = min(H, j0 + per);
      float o[GEN_IO];
#pragma unroll
      for (int k = 0; k < GEN_IO; ++k) o[k] = 0.0f;
      for (int j = j0; j < j1; ++j) {
        const float hv = hL[j * LD + s];
#pragma unroll
        for (int k = 0; k < GEN_IO; ++k)
          if (k < OUT) o[k] = fmaf(__ldg(Wout + (size_t)k * H + j), hv, o[k]);
      }
#pragma unroll
      for (int k = 0; k < GEN_IO; ++k) sRed[(q * GEN_IO + k) * LD + s] = o[k];
    }
    __syncthreads();

    // ---- head, loss and its gradient wrt the head outputs (thread per sample; ppo.py:225-264)
    if (tid < S) {
      float out[GEN_IO], dout[GEN_IO];
#pragma unroll
      for (int k = 0; k < GEN_IO; ++k) {
        float v = k < OUT ? __ldg(bout + k) : 0.0f;
        for (int q = 0; q < Q; ++q) v += sRed[(q * GEN_IO + k) * LD + tid];
        out[k] = v; dout[k] = 0.0f;
      }
      if (valid) {
        if (ACTOR) {
          const float oldlp = __ldg(a.logprobs + row), adv = __ldg(a.advantages + row);
          float newlogp, entropy;
          float dlp[GEN_IO], dH[GEN_IO];
          if (KIND == 0) {
            float m = out[0];
#pragma unroll
            for (int k = 1; k < GEN_IO; ++k) if (k < OUT) m = fmaxf(m, out[k]);
            float se = 0.0f;
#pragma unroll
            for (int k = 0; k < GEN_IO; ++k) if (k < OUT) se += expf(out[k] - m);
            const float lse = m + logf(se);
            const int act = (int)__ldg(a.actions + row);
            float lp[GEN_IO], pr[GEN_IO];
            entropy = 0.0f; newlogp = 0.0f;
#pragma unroll
            for (int k = 0; k < GEN_IO; ++k) {
              lp[k] = out[k] - lse;
              pr[k] = k < OUT ? expf(lp[k]) : 0.0f;
              if (k < OUT) entropy -= pr[k] * lp[k];
              if (k == act) newlogp = lp[k];
            }
#pragma unroll
            for (int k = 0; k < GEN_IO; ++k) {
              dlp[k] = k < OUT ? (k == act ? 1.0f : 0.0f) - pr[k] : 0.0f;
              dH[k] = k < OUT ? -pr[k] * (lp[k] + entropy) : 0.0f;
            }
          } else {
            const float LOG_SQRT_2PI = 0.91893853320467267f;
            newlogp = 0.0f; entropy = 0.0f;
#pragma unroll
            for (int k = 0; k < GEN_IO; ++k) {
              dlp[k] = 0.0f; dH[k] = 0.0f;
              if (k < OUT) {
                const float d = __ldg(a.actions + row * OUT + k) - out[k], var = sd[k] * sd[k];
                newlogp += -(d * d) / (2.0f * var) - ls[k] - LOG_SQRT_2PI;
                entropy += 0.5f + LOG_SQRT_2PI + ls[k];
                dlp[k] = d / var;
              }
            }
          }
          const float logr = newlogp - oldlp;
          const float ratio = expf(logr);
          const float advn = a.norm_adv ? (adv - adv_mean) / adv_den : adv;
          const float l1 = -advn * ratio;
          const float l2 = -advn * fminf(fmaxf(ratio, a.clip_lo), a.clip_hi);
          const float w1 = l1 > l2 ? 1.0f : (l1 == l2 ? 0.5f : 0.0f);
          const float inr = (ratio >= a.clip_lo && ratio <= a.clip_hi) ? 1.0f : 0.0f;
          const float g_logp = -advn * (w1 + (1.0f - w1) * inr) * ratio * a.inv_m;
          const float g_H = -a.ent_c * a.inv_m;
#pragma unroll
          for (int k = 0; k < GEN_IO; ++k) dout[k] = g_logp * dlp[k] + g_H * dH[k];
          if (KIND == 1) {
#pragma unroll
            for (int k = 0; k < GEN_IO; ++k)
              if (k < OUT) {
                const float d = out[k] - __ldg(a.actions + row * OUT + k);
                logstd_g[k] += g_logp * (d * d / (sd[k] * sd[k]) - 1.0f) + g_H;
              }
          }
          st0 += fmaxf(l1, l2); st1 += entropy; st2 += -logr; st3 += (ratio - 1.0f) - logr;
          st4 += fabsf(ratio - 1.0f) > a.clip ? 1.0f : 0.0f;
        } else {
          const float R = __ldg(a.returns + row), vold = __ldg(a.values + row), v = out[0];
          if (a.clip_vloss) {
            const float du = v - R, vu = du * du;
            const float d = v - vold, vc = vold + fminf(fmaxf(d, -a.clip), a.clip);
            const float dc = vc - R, lc = dc * dc;
            const float w1 = vu > lc ? 1.0f : (vu == lc ? 0.5f : 0.0f);
            const float inr = (d >= -a.clip && d <= a.clip) ? 1.0f : 0.0f;
            dout[0] = (w1 * du + (1.0f - w1) * dc * inr) * a.vf_c * a.inv_m;
            st0 += 0.5f * fmaxf(vu, lc);
          } else {
            const float d = v - vold;                  // reference quirk ppo.py:261: b_values, not b_returns
            dout[0] = d * a.vf_c * a.inv_m;
            st0 += 0.5f * d * d;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < GEN_IO; ++k) sDout[k * LD + tid] = dout[k];
    }
    __syncthreads();

    // ---- output layer gradients, then delta of the last hidden layer in place over its activation
    wgrad_small(sDout, hL, OUT, H, S, LD, part + oWout);
    bgrad(sDout, OUT, S, LD, part + oBout);
    __syncthreads();
    dense_bwd_data<TS>(Wout, sDout, OUT, hL, H, LD);
    __syncthreads();
    // ---- hidden layers, top down: buffer l+1 holds delta_{l+1}, buffer l still holds h_l
    for (int l = NL - 1; l >= 1; --l) {
      float* dn = sAct + (size_t)l * H * LD;
      float* hl = sAct + (size_t)(l - 1) * H * LD;
      const float* W = Wh + (size_t)(l - 1) * hstride;
      wgrad_big(dn, hl, H, H, S, LD, part + oWh + (size_t)(l - 1) * hstride);
      bgrad(dn, H, S, LD, part + oWh + (size_t)(l - 1) * hstride + (size_t)H * H);
      __syncthreads();
      dense_bwd_data<TS>(W, dn, H, hl, H, LD);
      __syncthreads();
    }
    // ---- first layer
    wgrad_small(sAct, sX, H, obs_dim, S, LD, part + oW0);
    bgrad(sAct, H, S, LD, part + oB0);
    __syncthreads();                                        // tile buffers free for the next tile
  }

  // ---- statistics and the log-std gradient: block sums into the partial slab
  float* sred8 = sRed;
  float s;
  if (KIND == 1) {
#pragma unroll
    for (int k = 0; k < GEN_IO; ++k) {
      s = block_sum_gen(logstd_g[k], sred8);
      if (tid == 0 && k < OUT) part[oLS + k] = s;
    }
  }
  float* stat = part + (g.pstride - AUR_NUM_STATS);
  s = block_sum_gen(st0, sred8); if (tid == 0) stat[ACTOR ? AUR_STAT_POLICY_LOSS : AUR_STAT_VALUE_LOSS] = s;
  if (ACTOR) {
    s = block_sum_gen(st1, sred8); if (tid == 0) stat[AUR_STAT_ENTROPY] = s;
    s = block_sum_gen(st2, sred8); if (tid == 0) stat[AUR_STAT_OLD_APPROX_KL] = s;
    s = block_sum_gen(st3, sred8); if (tid == 0) stat[AUR_STAT_APPROX_KL] = s;
    s = block_sum_gen(st4, sred8); if (tid == 0) stat[AUR_STAT_CLIPFRAC] = s;
  }
}

template <int TS>
__global__ void __launch_bounds__(GEN_THREADS, 2) ppo_grad_generic_kernel(GenDev g) {
  extern __shared__ __align__(16) float smem[];
  if (blockIdx.y == 0) {
    if (g.u.continuous) generic_net<1, TS>(g, smem);
    else generic_net<0, TS>(g, smem);
  } else {
    generic_net<2, TS>(g, smem);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static size_t gen_smem_bytes(int H, int NL, int S) {
  const int LD = S + 4, Q = GEN_THREADS / S;
  return sizeof(float) * ((size_t)(2 * GEN_IO + Q * GEN_IO) * LD + (size_t)NL * H * LD);
}
static int gen_pick_tile(int H, int NL) {
  // the largest tile that leaves room for two CTAs per SM, else the largest that fits at all
  for (int S : {128, 64, 32}) if (gen_smem_bytes(H, NL, S) <= 100 * 1024) return S;
  for (int S : {128, 64, 32}) if (gen_smem_bytes(H, NL, S) <= 220 * 1024) return S;
  return 0;
}
int gen_pstride(const aur_policy_desc& p) {
  const int64_t nA = net_param_count(p.obs_dim, p.hidden_dim, p.num_layers, p.act_dim) + (p.continuous ? p.act_dim : 0);
  const int64_t nC = net_param_count(p.obs_dim, p.hidden_dim, p.num_layers, 1);
  return (int)(((nA > nC ? nA : nC) + AUR_NUM_STATS + 3) / 4 * 4);
}
int gen_grid_x(const aur_policy_desc& p) {
  const int S = gen_pick_tile(p.hidden_dim, p.num_layers);
  const bool two = S && gen_smem_bytes(p.hidden_dim, p.num_layers, S) <= 100 * 1024;
  const int g = two ? sm_count() : sm_count() / 2;        // x 2 nets (blockIdx.y)
  return g < 1 ? 1 : g;
}
int check_generic_policy(const aur_policy_desc& p, const char* who) {
  if (p.hidden_dim < 4 || p.hidden_dim > 256 || (p.hidden_dim & 3)) {
    set_error("%s: hidden_dim %d outside the compiled range (multiples of 4 in 4..256); no fallback", who, p.hidden_dim);
    return AUR_ERR_UNSUPPORTED;
  }
  if (p.num_layers < 1 || p.num_layers > 16) { set_error("%s: num_layers %d outside 1..16", who, p.num_layers); return AUR_ERR_UNSUPPORTED; }
  if (p.obs_dim < 1 || p.obs_dim > GEN_IO || p.act_dim < 1 || p.act_dim > GEN_IO) {
    set_error("%s: obs_dim %d / act_dim %d outside 1..8", who, p.obs_dim, p.act_dim); return AUR_ERR_UNSUPPORTED;
  }
  if (!gen_pick_tile(p.hidden_dim, p.num_layers)) {
    set_error("%s: %d layers of %d units do not fit the activation tile in shared memory; no fallback", who, p.num_layers,
              p.hidden_dim);
    return AUR_ERR_UNSUPPORTED;
  }
  return 0;
}
// floats of workspace the generic kernel needs: staged parameters + [2][grid][pstride] partial slabs
size_t gen_workspace_floats(const aur_policy_desc& p) {
  const size_t staged = (size_t)(gen_staged_floats(p.obs_dim, p.hidden_dim, p.num_layers, p.act_dim) +
                                 gen_staged_floats(p.obs_dim, p.hidden_dim, p.num_layers, 1) + 3) / 4 * 4;
  return staged + (size_t)2 * gen_grid_x(p) * gen_pstride(p);
}

// Launches staging + gradient kernel; partial slabs end up at ws + staged floats.  Returns the grid width through gx_out.
int launch_ppo_grad_generic(const UpdDev& d, const aur_policy_desc& p, float* ws, int* gx_out, float** part_out, cudaStream_t s) {
  GenDev g;
  g.u = d;
  g.H = p.hidden_dim; g.NL = p.num_layers;
  g.S = gen_pick_tile(g.H, g.NL); g.LD = g.S + 4;
  g.pstride = gen_pstride(p);
  const size_t staged = (size_t)(gen_staged_floats(p.obs_dim, g.H, g.NL, p.act_dim) + gen_staged_floats(p.obs_dim, g.H, g.NL, 1) + 3) / 4 * 4;
  g.staged = ws;
  g.part = ws + staged;
  const int gx = gen_grid_x(p);
  stage_params_kernel<<<(unsigned)((staged + 255) / 256 > 1184 ? 1184 : (staged + 255) / 256), 256, 0, s>>>(
      d.params, p.obs_dim, g.H, g.NL, p.act_dim, ws);
  AUR_LAUNCH_OK("stage_params_kernel");
  const size_t smem = gen_smem_bytes(g.H, g.NL, g.S);
  const int ts = g.S / 32;
  void (*kern)(GenDev) = ts == 4 ? ppo_grad_generic_kernel<4> : ts == 2 ? ppo_grad_generic_kernel<2> : ppo_grad_generic_kernel<1>;
  // per device and per shape: set on every launch (a host-side table update, no device work)
  AUR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(gx, 2), GEN_THREADS, smem, s>>>(g);
  AUR_LAUNCH_OK("ppo_grad_generic_kernel");
  *gx_out = gx;
  *part_out = g.part;
  return 0;
}

}  // namespace aur
