// Tensor-core variant of the fused PPO minibatch update (row U): the three 64x64 contractions per net
//   forward        z2[s][j]  = sum_i h1[s][i] W2[j][i]
//   backward-data  dh1[s][i] = sum_j dz2[s][j] W2[j][i]
//   weight grad    dW2[j][i] = sum_s dz2[s][j] h1[s][i]
// run on tcgen05 with fp32 accumulation in TMEM.  Operands are bf16 two-term splits (x = hi + mid, residual
// <= 2^-18 |x|) and three products are issued per contraction (hi*hi + hi*mid + mid*hi), so a product is
// exact to ~1e-5 relative: the reference's 1e-4 bar on losses and gradients holds (tests/test_update_gpu.py).
//
// One CTA per SM trains BOTH nets on tiles of 128 samples; thread s <-> sample s <-> TMEM lane s, so every
// accumulator row comes back to the thread that owns the sample (tcgen05.ld 32x32b) and the per-sample
// work (first layer, tanh, head, loss, activation derivatives) stays in registers.  The operand rows are
// written by their owner threads straight into the 128-B-swizzled UMMA layout; the SAME tiles serve as
// K-major operands (forward / backward-data A), as MN-major B (W2 read transposed) and as MN-major A/B of
// the weight-gradient MMA, which takes [dz2_actor | dz2_critic]^T [h1_actor | h1_critic] as one 128x128
// accumulator living in TMEM for the whole kernel (its two diagonal 64x64 blocks are the two dW2).
// The small reductions (dW3, db3, db2, dW1, db1) stay SIMT over a feature-major staging buffer.
#include "tc.cuh"
#include "update.cuh"

namespace aur {

constexpr int TCU_S = 128;                 // samples per tile
constexpr int TCU_THREADS = 128;
constexpr int TCU_LD = TCU_S + 4;          // staging row stride (floats)
constexpr int TCU_TILE = TCU_S * 128;      // one operand tile: 128 rows x 128 B
constexpr int TCU_WTILE = 64 * 128;        // one W2 tile: 64 rows x 128 B
// shared memory map (bytes, from a 1024-aligned base)
constexpr int OFF_H1 = 0;                            // [hi: actor, critic][mid: actor, critic]
constexpr int OFF_DZ = OFF_H1 + 4 * TCU_TILE;        // same order
constexpr int OFF_W2 = OFF_DZ + 4 * TCU_TILE;        // [actor hi, actor mid, critic hi, critic mid]
constexpr int OFF_SMALL = OFF_W2 + 4 * TCU_WTILE;    // fp32 small weights, 2 nets x 1024 floats
constexpr int OFF_STAGE = OFF_SMALL + 2 * 1024 * 4;  // fp32 [64][TCU_LD]
constexpr int OFF_X = OFF_STAGE + 64 * TCU_LD * 4;   // fp32 [4][TCU_LD]
constexpr int OFF_DOUT = OFF_X + 4 * TCU_LD * 4;     // fp32 [4][TCU_LD]
constexpr int OFF_RED = OFF_DOUT + 4 * TCU_LD * 4;   // fp32 [256]
constexpr int OFF_BAR = OFF_RED + 256 * 4;           // 8 mbarriers + tmem slot
constexpr int TCU_SMEM_USED = OFF_BAR + 128;
constexpr size_t TCU_SMEM = TCU_SMEM_USED + 1024;

// small-weight block of one net (floats): W1 padded [64][4], b1 [64], b2 [64], W3 [4][64], b3 [4]
constexpr int SW_W1 = 0, SW_B1 = 256, SW_B2 = 320, SW_W3 = 384, SW_B3 = 640;

__device__ __forceinline__ unsigned int pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned int*>(&t);
}
// write one 64-element fp32 row as bf16 hi / mid rows of two 128-B-swizzled tiles (row r of the tile)
__device__ __forceinline__ void store_split_row(unsigned char* tile_hi, unsigned char* tile_mid, int r, const float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    unsigned int hi[4], mid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = v[8 * c + 2 * e], b = v[8 * c + 2 * e + 1];
      const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
      hi[e] = (unsigned int)__bfloat16_as_ushort(ha) | ((unsigned int)__bfloat16_as_ushort(hb) << 16);
      mid[e] = pack_bf16x2(a - __bfloat162float(ha), b - __bfloat162float(hb));
    }
    const int off = r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(tile_mid + off) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
  }
}
// read row r back as hi + mid
__device__ __forceinline__ void load_split_row(const unsigned char* tile_hi, const unsigned char* tile_mid, int r, float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int off = r * 128 + ((c ^ (r & 7)) << 4);
    const uint4 h = *reinterpret_cast<const uint4*>(tile_hi + off), m = *reinterpret_cast<const uint4*>(tile_mid + off);
    const unsigned int hw[4] = {h.x, h.y, h.z, h.w}, mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[8 * c + 2 * e] = __uint_as_float(hw[e] << 16) + __uint_as_float(mw[e] << 16);
      v[8 * c + 2 * e + 1] = __uint_as_float(hw[e] & 0xFFFF0000u) + __uint_as_float(mw[e] & 0xFFFF0000u);
    }
  }
}

struct TcuPtrs {
  unsigned char* base;
  __device__ __forceinline__ unsigned char* h1(int part, int net) const { return base + OFF_H1 + (part * 2 + net) * TCU_TILE; }
  __device__ __forceinline__ unsigned char* dz(int part, int net) const { return base + OFF_DZ + (part * 2 + net) * TCU_TILE; }
  __device__ __forceinline__ unsigned char* w2(int net, int part) const { return base + OFF_W2 + (net * 2 + part) * TCU_WTILE; }
  __device__ __forceinline__ float* small_w(int net) const { return reinterpret_cast<float*>(base + OFF_SMALL) + net * 1024; }
  __device__ __forceinline__ float* stage() const { return reinterpret_cast<float*>(base + OFF_STAGE); }
  __device__ __forceinline__ float* sx() const { return reinterpret_cast<float*>(base + OFF_X); }
  __device__ __forceinline__ float* sdout() const { return reinterpret_cast<float*>(base + OFF_DOUT); }
  __device__ __forceinline__ float* sred() const { return reinterpret_cast<float*>(base + OFF_RED); }
  __device__ __forceinline__ uint64_t* bar(int i) const { return reinterpret_cast<uint64_t*>(base + OFF_BAR) + i; }
  __device__ __forceinline__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(base + OFF_BAR + 96); }
};

// three-product split MMA over K = 64 (4 steps of 16): D (+)= A_hi B_hi + A_hi B_mid + A_mid B_hi
__device__ __forceinline__ void mma_split(uint32_t d, uint64_t a_hi, uint64_t a_mid, uint64_t b_hi, uint64_t b_mid, uint32_t idesc,
                                          int ksteps, uint32_t a_step, uint32_t b_step, bool accumulate) {
  for (int k = 0; k < ksteps; ++k)
    tc::mma_f16(d, a_hi + (uint64_t)(a_step * k), b_hi + (uint64_t)(b_step * k), idesc, (accumulate || k > 0) ? 1u : 0u);
  for (int k = 0; k < ksteps; ++k) tc::mma_f16(d, a_hi + (uint64_t)(a_step * k), b_mid + (uint64_t)(b_step * k), idesc, 1u);
  for (int k = 0; k < ksteps; ++k) tc::mma_f16(d, a_mid + (uint64_t)(a_step * k), b_hi + (uint64_t)(b_step * k), idesc, 1u);
}

__device__ __forceinline__ float block_sum_128(float v, float* sred) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  return (sred[0] + sred[1]) + (sred[2] + sred[3]);
}

__global__ void __launch_bounds__(TCU_THREADS, 1) ppo_grad_tc_kernel(UpdDev a) {
  extern __shared__ unsigned char smem_raw[];
  TcuPtrs P;
  P.base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int obs_dim = a.obs_dim, A = a.act_dim;
  const int64_t gA = net_param_count(obs_dim, UPD_H, 2, A), gC = net_param_count(obs_dim, UPD_H, 2, 1);

  // ---- one-time setup: barriers, TMEM, weights
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(P.bar(i), 1);
    mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc(P.tmem_slot(), 512);
  for (int net = 0; net < 2; ++net) {
    const int OUT = net == 0 ? A : 1;
    const float* g = net == 0 ? a.params : a.params + gA;
    float* sw = P.small_w(net);
    for (int e = tid; e < 256; e += TCU_THREADS) { const int j = e >> 2, c = e & 3; sw[SW_W1 + e] = c < obs_dim ? g[j * obs_dim + c] : 0.0f; }
    const float* gb1 = g + 64 * obs_dim;
    const float* gW2 = gb1 + 64;
    const float* gb2 = gW2 + 4096;
    const float* gW3 = gb2 + 64;
    const float* gb3 = gW3 + OUT * 64;
    for (int e = tid; e < 64; e += TCU_THREADS) { sw[SW_B1 + e] = gb1[e]; sw[SW_B2 + e] = gb2[e]; }
    for (int e = tid; e < 256; e += TCU_THREADS) sw[SW_W3 + e] = e < OUT * 64 ? gW3[e] : 0.0f;
    if (tid < 4) sw[SW_B3 + tid] = tid < OUT ? gb3[tid] : 0.0f;
    if (tid < 64) {                                   // W2 row j = tid -> hi / mid swizzled rows
      float row[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) row[i] = gW2[tid * 64 + i];      // the critic's block is only 4-byte aligned
      store_split_row(P.w2(net, 0), P.w2(net, 1), tid, row);
    }
  }
  NormalConsts nc;
  float logstd_g[POL_OUT_MAX] = {0.f, 0.f, 0.f, 0.f};
  if (a.continuous) nc = normal_consts(a.params + gA + gC, A);
  float adv_mean = 0.0f, adv_den = 1.0f;
  if (a.norm_adv) {
    const double n = a.moments[2], s = a.moments[0], ss = a.moments[1];
    const double mean = s / n;
    double var = (ss - s * mean) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    adv_mean = (float)mean; adv_den = (float)sqrt(var) + 1e-8f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *P.tmem_slot();
  const uint32_t tm_z[2] = {tmem + 0, tmem + 64}, tm_dh[2] = {tmem + 128, tmem + 192}, tm_w = tmem + 256;
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;

  constexpr uint32_t ID_FWD = tc::instr_desc(tc::FMT_BF16, 128, 64, 0, 0);
  constexpr uint32_t ID_BWD = tc::instr_desc(tc::FMT_BF16, 128, 64, 0, 1);
  constexpr uint32_t ID_WG = tc::instr_desc(tc::FMT_BF16, 128, 128, 1, 1);

  // per-thread accumulators of the SIMT reductions (thread t: j = t & 63, half = t >> 6)
  float acc_w3[2][2] = {{0.f, 0.f}, {0.f, 0.f}};    // [net][pass]: dW3[k = half + 2*pass][j]
  float acc_b3[2] = {0.f, 0.f};                      // thread k < OUT
  float acc_b2[2] = {0.f, 0.f}, acc_b1[2] = {0.f, 0.f};
  float acc_w1[2][2] = {{0.f, 0.f}, {0.f, 0.f}};    // [net][pass]: dW1[j][c = half + 2*pass]
  float st_pl = 0.f, st_ent = 0.f, st_okl = 0.f, st_kl = 0.f, st_clip = 0.f, st_vl = 0.f;

  const int j_of = tid & 63, half = tid >> 6;
  const long long ntiles = (a.m_local + TCU_S - 1) / TCU_S;
  uint32_t it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const uint32_t ph = it & 1u;
    // ================= S1: gather, first layers, operand rows of h1 =================
    const long long gi = tile * TCU_S + tid;
    const bool valid = gi < a.m_local;
    const long long row = valid ? (a.idx ? (long long)a.idx[gi] : a.idx_offset + gi) : 0;
    float x[1][POL_IN_PAD];
#pragma unroll
    for (int c = 0; c < POL_IN_PAD; ++c) {
      x[0][c] = (valid && c < obs_dim) ? __ldg(a.obs + row * obs_dim + c) : 0.0f;
      P.sx()[c * TCU_LD + tid] = x[0][c];
    }
#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
      const float* sw = P.small_w(net);
      float2 h1p[1][UPD_H / 2];
      mlp_first_layer<UPD_H, 1>(sw + SW_W1, sw + SW_B1, x, h1p);
      float h1[64];
#pragma unroll
      for (int q = 0; q < 32; ++q) { h1[2 * q] = h1p[0][q].x; h1[2 * q + 1] = h1p[0][q].y; }
      store_split_row(P.h1(0, net), P.h1(1, net), tid, h1);
    }
    tc::fence_proxy_async();
    __syncthreads();
    // ================= S2: forward MMAs (both nets) =================
    if (tid == 0) {
      tc::fence_after_sync();
      for (int net = 0; net < 2; ++net) {
        mma_split(tm_z[net], tc::smem_desc_k_sw128(P.h1(0, net)), tc::smem_desc_k_sw128(P.h1(1, net)),
                  tc::smem_desc_k_sw128(P.w2(net, 0)), tc::smem_desc_k_sw128(P.w2(net, 1)), ID_FWD, 4, 2, 2, false);
        tc::mma_commit(P.bar(net));
      }
    }
    // ================= S3-S5 per net: head, loss, dz2, small reductions =================
#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
      const int OUT = net == 0 ? A : 1;
      const float* sw = P.small_w(net);
      mbar_wait(P.bar(net), ph);
      tc::fence_after_sync();
      float h2[64];
      {
        float v[32];
        tc::tmem_ld32(tm_z[net] + lane_base, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) h2[i] = tanh_fast(v[i] + sw[SW_B2 + i]);
        tc::tmem_ld32(tm_z[net] + lane_base + 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) h2[32 + i] = tanh_fast(v[i] + sw[SW_B2 + 32 + i]);
      }
      float out[POL_OUT_MAX], dout[POL_OUT_MAX];
#pragma unroll
      for (int k = 0; k < POL_OUT_MAX; ++k) {
        float s = 0.0f;
        if (k < OUT) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 w = lds4(sw + SW_W3 + k * 64 + j);
            s = fmaf(w.x, h2[j], s); s = fmaf(w.y, h2[j + 1], s); s = fmaf(w.z, h2[j + 2], s); s = fmaf(w.w, h2[j + 3], s);
          }
          s += sw[SW_B3 + k];
        }
        out[k] = s; dout[k] = 0.0f;
      }
      if (valid) {
        if (net == 0) {
          const float oldlp = __ldg(a.logprobs + row), adv = __ldg(a.advantages + row);
          float newlogp, entropy, dlp[POL_OUT_MAX], dH[POL_OUT_MAX];
          if (!a.continuous) {
            float m = out[0];
#pragma unroll
            for (int k = 1; k < POL_OUT_MAX; ++k) if (k < OUT) m = fmaxf(m, out[k]);
            float se = 0.0f;
#pragma unroll
            for (int k = 0; k < POL_OUT_MAX; ++k) if (k < OUT) se += expf(out[k] - m);
            const float lse = m + logf(se);
            const int act = (int)__ldg(a.actions + row);
            float lp[POL_OUT_MAX], pr[POL_OUT_MAX];
            entropy = 0.0f; newlogp = 0.0f;
#pragma unroll
            for (int k = 0; k < POL_OUT_MAX; ++k) {
              lp[k] = out[k] - lse;
              pr[k] = k < OUT ? expf(lp[k]) : 0.0f;
              if (k < OUT) entropy -= pr[k] * lp[k];
              if (k == act) newlogp = lp[k];
            }
#pragma unroll
            for (int k = 0; k < POL_OUT_MAX; ++k) {
              dlp[k] = (k == act ? 1.0f : 0.0f) - pr[k];
              dH[k] = k < OUT ? -pr[k] * (lp[k] + entropy) : 0.0f;
            }
          } else {
            float act[POL_OUT_MAX];
#pragma unroll
            for (int k = 0; k < POL_OUT_MAX; ++k) act[k] = k < OUT ? __ldg(a.actions + row * OUT + k) : 0.0f;
            normal_logp(out, act, OUT, nc, newlogp, entropy);
#pragma unroll
            for (int k = 0; k < POL_OUT_MAX; ++k) {
              const float d = act[k] - out[k], var = nc.std[k] * nc.std[k];
              dlp[k] = k < OUT ? d / var : 0.0f;
              dH[k] = 0.0f;
            }
          }
          const float logr = newlogp - oldlp, ratio = expf(logr);
          const float advn = a.norm_adv ? (adv - adv_mean) / adv_den : adv;
          const float l1 = -advn * ratio, l2 = -advn * fminf(fmaxf(ratio, a.clip_lo), a.clip_hi);
          const float w1 = l1 > l2 ? 1.0f : (l1 == l2 ? 0.5f : 0.0f);
          const float inr = (ratio >= a.clip_lo && ratio <= a.clip_hi) ? 1.0f : 0.0f;
          const float g_logp = -advn * (w1 + (1.0f - w1) * inr) * ratio * a.inv_m;
          const float g_H = -a.ent_c * a.inv_m;
#pragma unroll
          for (int k = 0; k < POL_OUT_MAX; ++k) dout[k] = g_logp * dlp[k] + g_H * dH[k];
          if (a.continuous) {
#pragma unroll
            for (int k = 0; k < POL_OUT_MAX; ++k)
              if (k < OUT) {
                const float d = out[k] - __ldg(a.actions + row * OUT + k);
                logstd_g[k] += g_logp * (d * d / (nc.std[k] * nc.std[k]) - 1.0f) + g_H;
              }
          }
          st_pl += fmaxf(l1, l2); st_ent += entropy; st_okl += -logr; st_kl += (ratio - 1.0f) - logr;
          st_clip += fabsf(ratio - 1.0f) > a.clip ? 1.0f : 0.0f;
        } else {
          const float R = __ldg(a.returns + row), vold = __ldg(a.values + row), v = out[0];
          if (a.clip_vloss) {
            const float du = v - R, vu = du * du;
            const float d = v - vold, vc = vold + fminf(fmaxf(d, -a.clip), a.clip);
            const float dc = vc - R, lc = dc * dc;
            const float w1 = vu > lc ? 1.0f : (vu == lc ? 0.5f : 0.0f);
            const float inr = (d >= -a.clip && d <= a.clip) ? 1.0f : 0.0f;
            dout[0] = (w1 * du + (1.0f - w1) * dc * inr) * a.vf_c * a.inv_m;
            st_vl += 0.5f * fmaxf(vu, lc);
          } else {
            const float d = v - vold;
            dout[0] = d * a.vf_c * a.inv_m;
            st_vl += 0.5f * d * d;
          }
        }
      }
      // stage h2 (feature-major) and dout for the dW3 / db3 reduction
      float* stg = P.stage();
#pragma unroll
      for (int j = 0; j < 64; ++j) stg[j * TCU_LD + tid] = h2[j];
#pragma unroll
      for (int k = 0; k < POL_OUT_MAX; ++k) P.sdout()[k * TCU_LD + tid] = dout[k];
      __syncthreads();
      // G1: dW3[k][j] += sum_s dout[s][k] h2[s][j]   (thread: j = tid & 63, k = half + 2 * pass)
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int k = half + 2 * pass;
        if (k < OUT) {
          const float* hp = stg + j_of * TCU_LD;
          const float* dp = P.sdout() + k * TCU_LD;
          float2 s2 = make_float2(0.f, 0.f);
#pragma unroll 8
          for (int s4 = 0; s4 < TCU_S / 4; ++s4) {
            const float4 hv = lds4(hp + 4 * s4), dv = lds4(dp + 4 * s4);
            s2 = __ffma2_rn(make_float2(hv.x, hv.y), make_float2(dv.x, dv.y), s2);
            s2 = __ffma2_rn(make_float2(hv.z, hv.w), make_float2(dv.z, dv.w), s2);
          }
          acc_w3[net][pass] += s2.x + s2.y;
        }
      }
      if (tid < OUT) {
        const float* dp = P.sdout() + tid * TCU_LD;
        float s = 0.0f;
        for (int s4 = 0; s4 < TCU_S / 4; ++s4) { const float4 dv = lds4(dp + 4 * s4); s += (dv.x + dv.y) + (dv.z + dv.w); }
        acc_b3[net] += s;
      }
      // dz2 = (W3^T dout) * (1 - h2^2): operand rows for the backward MMAs
      float dz2[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        float dh = 0.0f;
#pragma unroll
        for (int k = 0; k < POL_OUT_MAX; ++k) if (k < OUT) dh = fmaf(sw[SW_W3 + k * 64 + j], dout[k], dh);
        dz2[j] = dh * fmaf(-h2[j], h2[j], 1.0f);
      }
      store_split_row(P.dz(0, net), P.dz(1, net), tid, dz2);
      __syncthreads();                         // G1 done reading the staging buffer
#pragma unroll
      for (int j = 0; j < 64; ++j) stg[j * TCU_LD + tid] = dz2[j];
      __syncthreads();
      {                                        // db2[j] += sum_s dz2[s][j]   (thread: j = tid & 63, samples half*64 ..)
        const float* p = stg + j_of * TCU_LD + half * 64;
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < 16; ++q) { const float4 v = lds4(p + 4 * q); s += (v.x + v.y) + (v.z + v.w); }
        acc_b2[net] += s;
      }
      __syncthreads();
    }
    // ================= S6: backward-data MMAs + weight-gradient MMA =================
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      for (int net = 0; net < 2; ++net) {
        mma_split(tm_dh[net], tc::smem_desc_k_sw128(P.dz(0, net)), tc::smem_desc_k_sw128(P.dz(1, net)),
                  tc::smem_desc_mn_sw128(P.w2(net, 0), 8192, 1024), tc::smem_desc_mn_sw128(P.w2(net, 1), 8192, 1024), ID_BWD, 4, 2,
                  128, false);
        tc::mma_commit(P.bar(2 + net));
      }
      // D_w[m][n] (+)= sum_s [dz2_a | dz2_c][s][m] * [h1_a | h1_c][s][n], K = 128 samples = 8 steps of 16 rows
      mma_split(tm_w, tc::smem_desc_mn_sw128(P.dz(0, 0), TCU_TILE, 1024), tc::smem_desc_mn_sw128(P.dz(1, 0), TCU_TILE, 1024),
                tc::smem_desc_mn_sw128(P.h1(0, 0), TCU_TILE, 1024), tc::smem_desc_mn_sw128(P.h1(1, 0), TCU_TILE, 1024), ID_WG, 8, 128,
                128, it > 0);
      tc::mma_commit(P.bar(4));
    }
    // ================= S7 per net: dz1, dW1, db1 =================
#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
      mbar_wait(P.bar(2 + net), ph);
      tc::fence_after_sync();
      float dz1[64];
      {
        float v[32];
        tc::tmem_ld32(tm_dh[net] + lane_base, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) dz1[i] = v[i];
        tc::tmem_ld32(tm_dh[net] + lane_base + 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) dz1[32 + i] = v[i];
      }
      {
        float h1[64];
        load_split_row(P.h1(0, net), P.h1(1, net), tid, h1);
#pragma unroll
        for (int i = 0; i < 64; ++i) dz1[i] *= fmaf(-h1[i], h1[i], 1.0f);
      }
      float* stg = P.stage();
#pragma unroll
      for (int i = 0; i < 64; ++i) stg[i * TCU_LD + tid] = dz1[i];
      __syncthreads();
      {
        const float* zp = stg + j_of * TCU_LD;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const int c = half + 2 * pass;
          const float* xp = P.sx() + c * TCU_LD;
          float2 s2 = make_float2(0.f, 0.f);
#pragma unroll 8
          for (int s4 = 0; s4 < TCU_S / 4; ++s4) {
            const float4 zv = lds4(zp + 4 * s4), xv = lds4(xp + 4 * s4);
            s2 = __ffma2_rn(make_float2(zv.x, zv.y), make_float2(xv.x, xv.y), s2);
            s2 = __ffma2_rn(make_float2(zv.z, zv.w), make_float2(xv.z, xv.w), s2);
          }
          acc_w1[net][pass] += s2.x + s2.y;
        }
        const float* p = zp + half * 64;
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < 16; ++q) { const float4 v = lds4(p + 4 * q); s += (v.x + v.y) + (v.z + v.w); }
        acc_b1[net] += s;
      }
      __syncthreads();
    }
    // the weight-gradient MMA reads the h1 / dz2 tiles: it must retire before the next tile overwrites them
    mbar_wait(P.bar(4), ph);
    tc::fence_after_sync();
  }

  // ================= epilogue: partials in the nets' flat parameter order =================
  const bool any = (long long)blockIdx.x < ntiles;
  float* sred = P.sred();
  for (int net = 0; net < 2; ++net) {
    const int OUT = net == 0 ? A : 1;
    float* part = a.partials + ((size_t)net * gridDim.x + blockIdx.x) * UPD_PSTRIDE;
    const int oW1 = 0, oB1 = 64 * obs_dim, oW2 = oB1 + 64, oB2 = oW2 + 4096, oW3 = oB2 + 64, oB3 = oW3 + OUT * 64, oLS = oB3 + OUT;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int c = half + 2 * pass;
      if (c < obs_dim) part[oW1 + j_of * obs_dim + c] = acc_w1[net][pass];
      if (c < OUT) part[oW3 + c * 64 + j_of] = acc_w3[net][pass];
    }
    if (tid < OUT) part[oB3 + tid] = acc_b3[net];
    __syncthreads();
    sred[tid] = acc_b2[net];
    sred[128 + tid] = acc_b1[net];
    __syncthreads();
    if (tid < 64) {
      part[oB2 + tid] = sred[tid] + sred[tid + 64];
      part[oB1 + tid] = sred[128 + tid] + sred[128 + tid + 64];
    }
    // dW2: thread t owns accumulator row m = t: rows 0..63 -> actor (columns 0..63), rows 64..127 -> critic (64..127)
    if ((tid >> 6) == net) {
      const int j = tid & 63;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        if (any) tc::tmem_ld32(tm_w + lane_base + (uint32_t)(net * 64 + c), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) part[oW2 + j * 64 + c + i] = any ? v[i] : 0.0f;
      }
    }
    if (net == 0 && a.continuous) {
      for (int k = 0; k < POL_OUT_MAX; ++k) {
        const float s = block_sum_128(logstd_g[k], sred);
        if (tid == 0 && k < OUT) part[oLS + k] = s;
      }
    }
    float* stat = part + UPD_STAT_OFF;
    if (net == 0) {
      float s;
      s = block_sum_128(st_pl, sred); if (tid == 0) stat[AUR_STAT_POLICY_LOSS] = s;
      s = block_sum_128(st_ent, sred); if (tid == 0) stat[AUR_STAT_ENTROPY] = s;
      s = block_sum_128(st_okl, sred); if (tid == 0) stat[AUR_STAT_OLD_APPROX_KL] = s;
      s = block_sum_128(st_kl, sred); if (tid == 0) stat[AUR_STAT_APPROX_KL] = s;
      s = block_sum_128(st_clip, sred); if (tid == 0) stat[AUR_STAT_CLIPFRAC] = s;
    } else {
      const float s = block_sum_128(st_vl, sred);
      if (tid == 0) stat[AUR_STAT_VALUE_LOSS] = s;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

size_t ppo_grad_tc_smem_bytes() { return TCU_SMEM; }

int launch_ppo_grad_tc(const UpdDev& d, int gx, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    AUR_CUDA_OK(cudaFuncSetAttribute(ppo_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCU_SMEM));
    attr = true;
  }
  ppo_grad_tc_kernel<<<gx, TCU_THREADS, TCU_SMEM, s>>>(d);
  AUR_LAUNCH_OK("ppo_grad_tc_kernel");
  return 0;
}

}  // namespace aur
