// Tensor-core implementation of the fused PPO minibatch update (row U, src/ppo.py:220-267).
//
// Every contraction over a 64-wide feature axis or over the samples of a tile runs on tcgen05 with fp32
// accumulation in TMEM:
//   forward        z2[s][j]  = sum_i h1[s][i] W2[j][i]
//   backward-data  dh1[s][i] = sum_j dz2[s][j] W2[j][i]
//   weight grad    dW2[j][i] = sum_s dz2[s][j] h1[s][i]
//   bias / input   db2[j] = sum_s dz2[s][j],  db1[j] = sum_s dz1[s][j],  dW1[j][c] = sum_s dz1[s][j] x[s][c]
// Operands are bf16 two-term splits (v = hi + mid, residual <= 2^-17 |v|) and three products are issued per
// contraction (hi*hi + hi*mid + mid*hi), so a product is exact to ~1e-5 relative: the reference's 1e-4 bar on
// losses and gradients holds (tests/test_update_gpu.py).
//
// One persistent CTA of 512 threads per SM walks tiles of 128 samples and trains BOTH nets.  Thread =
// (sample s, net, feature half): 4 threads share a sample, each owning 32 of the 64 hidden features of one
// net, so 16 warps keep the FMA/MUFU pipes busy while the per-sample state stays in registers.  TMEM lane s
// is sample s: every accumulator row comes back to the threads that own the sample (tcgen05.ld 32x32b, 32
// columns per thread).  The operand rows are written by their owners straight into the 128-B-swizzled UMMA
// layout; the SAME tiles serve as K-major A operands (forward / backward-data) and as MN-major A/B operands of
// the contractions over samples.  [dz2_actor | dz2_critic]^T [h1_actor | h1_critic] is one 128x128 accumulator
// that lives in TMEM for the whole kernel (its two diagonal 64x64 blocks are the two dW2); the bias and
// first-layer gradients come from the same A tiles against a 16-column K-major "aux" tile [1, x0..x3].
// Only dW3 / db3 (out_dim <= 4 columns) stay SIMT, staged through shared memory in two 32-feature rounds.
//
// Pipeline per tile t (MMA batches are issued by two elected threads, one per batch, and tracked with mbarriers):
//   dz1(t-1) <- bwd(t-1) | h1(t) | wait wgrad(t-1) | store dz1(t-1), h1(t), aux(t) | issue fwd(t), aux_w1(t-1)
//   prefetch gather(t+1) | wait fwd(t) | h2, head, loss | wait aux_w1(t-1) | store dz2(t)
//   issue bwd(t), wgrad(t), aux_b2(t) | dW3 rounds (overlap the MMAs)
#include "tc.cuh"
#include "update.cuh"

namespace aur {

constexpr int T2_S = 128;                   // samples per tile
constexpr int T2_THREADS = 512;
constexpr int T2_LD = T2_S + 4;             // fp32 staging row stride
constexpr int T2_TILE = T2_S * 128;         // operand tile: 128 rows x 128 B
constexpr int T2_WTILE = 64 * 128;          // W2 tile: 64 rows x 128 B
constexpr int T2_AUXT = 2 * 16 * 128;       // aux tile: [sample block 2][n 16][64 samples] bf16, K-major
constexpr int T2_SW = 768;                  // small-weight floats per net
// shared memory map (bytes, from a 1024-aligned base)
constexpr int O2_H1 = 0;                                // [hi: actor, critic][mid: actor, critic]
constexpr int O2_DZ = O2_H1 + 4 * T2_TILE;              // same order
constexpr int O2_W2 = O2_DZ + 4 * T2_TILE;              // [actor hi, actor mid, critic hi, critic mid]
constexpr int O2_AUX = O2_W2 + 4 * T2_WTILE;            // [buffer 2][hi, mid]
constexpr int O2_SMALL = O2_AUX + 4 * T2_AUXT;          // fp32 small weights, 2 nets
constexpr int O2_STAGE = O2_SMALL + 2 * T2_SW * 4;      // fp32 [net][32][T2_LD]; also head exchange / epilogue scratch
constexpr int O2_DOUT = O2_STAGE + 2 * 32 * T2_LD * 4;  // fp32 [net][4][T2_LD]
constexpr int O2_BAR = O2_DOUT + 2 * 4 * T2_LD * 4;     // mbarriers + tmem slot
constexpr int T2_SMEM_USED = O2_BAR + 128;
constexpr size_t T2_SMEM = T2_SMEM_USED + 1024;
static_assert(T2_SMEM <= 232448, "shared memory budget");

// small-weight block of one net (floats): W1^T [4][64] (zero rows >= obs_dim), b1 [64], b2 [64], W3 [4][64], b3 [4]
constexpr int S2_W1T = 0, S2_B1 = 256, S2_B2 = 320, S2_W3 = 384, S2_B3 = 640;
enum { BAR_FWD = 0, BAR_AUX = 1, BAR_BWD = 2, BAR_WG = 3, BAR_FIN = 4 };

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 8 fp32 values -> one 16-B chunk of bf16 hi and one of bf16 mid (v - hi), chunk `c` of row `r` (128-B swizzle)
__device__ __forceinline__ void store_split_chunk(unsigned char* tile_hi, unsigned char* tile_mid, int r, int c, const float* v) {
  unsigned int hi[4], mid[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = v[2 * e], b = v[2 * e + 1];
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const unsigned int hw = *reinterpret_cast<unsigned int*>(&h);
    hi[e] = hw;
    __nv_bfloat162 m = __floats2bfloat162_rn(a - __uint_as_float(hw << 16), b - __uint_as_float(hw & 0xFFFF0000u));
    mid[e] = *reinterpret_cast<unsigned int*>(&m);
  }
  const int off = r * 128 + ((c ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(tile_mid + off) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
}

struct T2Ptrs {
  unsigned char* base;
  __device__ __forceinline__ unsigned char* h1(int part, int net) const { return base + O2_H1 + (part * 2 + net) * T2_TILE; }
  __device__ __forceinline__ unsigned char* dz(int part, int net) const { return base + O2_DZ + (part * 2 + net) * T2_TILE; }
  __device__ __forceinline__ unsigned char* w2(int net, int part) const { return base + O2_W2 + (net * 2 + part) * T2_WTILE; }
  __device__ __forceinline__ unsigned char* aux(int buf, int part) const { return base + O2_AUX + (buf * 2 + part) * T2_AUXT; }
  __device__ __forceinline__ float* small_w(int net) const { return reinterpret_cast<float*>(base + O2_SMALL) + net * T2_SW; }
  __device__ __forceinline__ float* stage(int net) const { return reinterpret_cast<float*>(base + O2_STAGE) + net * 32 * T2_LD; }
  __device__ __forceinline__ float* sdout(int net) const { return reinterpret_cast<float*>(base + O2_DOUT) + net * 4 * T2_LD; }
  __device__ __forceinline__ uint64_t* bar(int i) const { return reinterpret_cast<uint64_t*>(base + O2_BAR) + i; }
  __device__ __forceinline__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(base + O2_BAR + 96); }
};

// byte offset of element (n, s) of an aux tile (K-major, 128-B swizzle): row n holds 64 samples of one block
__device__ __forceinline__ int aux_off(int n, int s) {
  return (s >> 6) * 2048 + n * 128 + (((((s & 63) >> 3) ^ (n & 7))) << 4) + (s & 7) * 2;
}

// three-product split MMA: D (+)= A_hi B_hi + A_hi B_mid + A_mid B_hi over `ksteps` steps of K = 16
__device__ __forceinline__ void mma_split(uint32_t d, uint64_t a_hi, uint64_t a_mid, uint64_t b_hi, uint64_t b_mid, uint32_t idesc,
                                          int ksteps, uint32_t a_step, uint32_t b_step, bool accumulate) {
  for (int k = 0; k < ksteps; ++k)
    tc::mma_f16(d, a_hi + (uint64_t)(a_step * k), b_hi + (uint64_t)(b_step * k), idesc, (accumulate || k > 0) ? 1u : 0u);
  for (int k = 0; k < ksteps; ++k) tc::mma_f16(d, a_hi + (uint64_t)(a_step * k), b_mid + (uint64_t)(b_step * k), idesc, 1u);
  for (int k = 0; k < ksteps; ++k) tc::mma_f16(d, a_mid + (uint64_t)(a_step * k), b_hi + (uint64_t)(b_step * k), idesc, 1u);
}
// D[128][16] (+)= [A_actor | A_critic]^T (MN-major, K = 128 samples) x aux (K-major, two 64-sample blocks)
__device__ __forceinline__ void mma_aux(uint32_t d, uint64_t a_hi, uint64_t a_mid, uint64_t b_hi, uint64_t b_mid, uint32_t idesc,
                                        bool accumulate) {
#pragma unroll 1
  for (int p = 0; p < 3; ++p) {
    const uint64_t ad = p == 2 ? a_mid : a_hi, bd = p == 1 ? b_mid : b_hi;
    for (int k = 0; k < 8; ++k)
      tc::mma_f16(d, ad + (uint64_t)(128 * k), bd + (uint64_t)((k & 3) * 2 + (k >> 2) * 128), idesc,
                  (accumulate || p > 0 || k > 0) ? 1u : 0u);
  }
}

__global__ void __launch_bounds__(T2_THREADS, 1) ppo_grad_tc_kernel(UpdDev a) {
  extern __shared__ unsigned char smem_raw[];
  T2Ptrs P;
  P.base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int s = tid & 127, q = tid >> 7, net = q >> 1, half = q & 1, f0 = 32 * half;
  const int obs_dim = a.obs_dim, A = a.act_dim;
  const int OUT = net == 0 ? A : 1;
  const int64_t gA = net_param_count(obs_dim, UPD_H, 2, A), gC = net_param_count(obs_dim, UPD_H, 2, 1);

  // ---- one-time setup: barriers, TMEM, weights, aux tiles
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(P.bar(i), 1);
    mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc(P.tmem_slot(), 512);
  for (int n = 0; n < 2; ++n) {
    const int O = n == 0 ? A : 1;
    const float* g = n == 0 ? a.params : a.params + gA;
    float* w = P.small_w(n);
    const float* gb1 = g + 64 * obs_dim;
    const float* gW2 = gb1 + 64;
    const float* gb2 = gW2 + 4096;
    const float* gW3 = gb2 + 64;
    const float* gb3 = gW3 + O * 64;
    for (int e = tid; e < 256; e += T2_THREADS) {
      const int c = e >> 6, j = e & 63;
      w[S2_W1T + e] = c < obs_dim ? g[j * obs_dim + c] : 0.0f;
      w[S2_W3 + e] = e < O * 64 ? gW3[e] : 0.0f;
    }
    for (int e = tid; e < 64; e += T2_THREADS) { w[S2_B1 + e] = gb1[e]; w[S2_B2 + e] = gb2[e]; }
    if (tid < 4) w[S2_B3 + tid] = tid < O ? gb3[tid] : 0.0f;
    if ((tid >> 6) == n) {                             // W2 row j -> hi / mid swizzled rows (block only 4-byte aligned)
      const int j = tid & 63;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = gW2[j * 64 + 8 * c + e];
        store_split_chunk(P.w2(n, 0), P.w2(n, 1), j, c, v);
      }
    }
  }
  {                                                    // aux tiles: zero, then the row of ones (n = 0)
    uint4* ax = reinterpret_cast<uint4*>(P.aux(0, 0));
    for (int e = tid; e < 4 * T2_AUXT / 16; e += T2_THREADS) ax[e] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  if (tid < 256) *reinterpret_cast<unsigned short*>(P.aux(tid >> 7, 0) + aux_off(0, tid & 127)) = 0x3F80;

  NormalConsts nc;
  if (a.continuous) nc = normal_consts(a.params + gA + gC, A);
  float adv_mean = 0.0f, adv_den = 1.0f;
  if (a.norm_adv) {
    const double n = a.moments[2], sm = a.moments[0], ss = a.moments[1];
    const double mean = sm / n;
    double var = (ss - sm * mean) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    adv_mean = (float)mean; adv_den = (float)sqrt(var) + 1e-8f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *P.tmem_slot();
  const uint32_t tm_z0 = tmem, tm_dh0 = tmem + 128, tm_w = tmem + 256, tm_b2 = tmem + 384, tm_w1 = tmem + 400;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;

  constexpr uint32_t ID_FWD = tc::instr_desc(tc::FMT_BF16, 128, 64, 0, 0);
  constexpr uint32_t ID_BWD = tc::instr_desc(tc::FMT_BF16, 128, 64, 0, 1);
  constexpr uint32_t ID_WG = tc::instr_desc(tc::FMT_BF16, 128, 128, 1, 1);
  constexpr uint32_t ID_AUX = tc::instr_desc(tc::FMT_BF16, 128, 16, 1, 0);

  const float* sw = P.small_w(net);
  float* stg = P.stage(net);
  float* sdo = P.sdout(net);
  const int barid = 1 + net;

  // accumulators that stay in registers for the whole kernel
  float acc_w3[2][POL_OUT_MAX];        // dW3[k][32 r + (t' & 31)] over this thread's 16-sample slice (t' = s + 128 half)
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int k = 0; k < POL_OUT_MAX; ++k) acc_w3[r][k] = 0.0f;
  float acc_b3 = 0.0f;
  float st[9];                         // actor: policy loss, entropy, old kl, kl, clipfrac, d logstd[4]; critic: st[0] = value loss
#pragma unroll
  for (int i = 0; i < 9; ++i) st[i] = 0.0f;
  const int tp = s + 128 * half, ri = tp & 31, c8 = tp >> 5;

  const long long ntiles = (a.m_local + T2_S - 1) / T2_S;
  const bool any = (long long)blockIdx.x < ntiles;
  // gather prefetch: row index and observation of the next tile
  float xn[POL_IN_PAD];
  long long rown = 0;
  bool validn = false;
  auto prefetch_tile = [&](long long tile) {
    const long long gi = tile * T2_S + s;
    validn = tile < ntiles && gi < a.m_local;
    rown = validn ? (a.idx ? (long long)a.idx[gi] : a.idx_offset + gi) : 0;
    if (validn && obs_dim == 4 && (reinterpret_cast<uintptr_t>(a.obs) & 15) == 0) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.obs) + rown);
      xn[0] = v.x; xn[1] = v.y; xn[2] = v.z; xn[3] = v.w;
    } else {
#pragma unroll
      for (int c = 0; c < POL_IN_PAD; ++c) xn[c] = (validn && c < obs_dim) ? __ldg(a.obs + rown * obs_dim + c) : 0.0f;
    }
    if (validn) {
      if (q == 0) { prefetch_l2(a.logprobs + rown); prefetch_l2(a.advantages + rown); prefetch_l2(a.actions + rown * (a.continuous ? A : 1)); }
      if (q == 2) { prefetch_l2(a.returns + rown); prefetch_l2(a.values + rown); }
    }
  };
  prefetch_tile(blockIdx.x);

  float h1[32];
  uint32_t it = 0;
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const uint32_t ph = it & 1u;
    // ---- backward of the previous tile: dz1 = dh1 * (1 - h1^2)
    float dz1[32];
    if (it > 0) {
      mbar_wait(P.bar(BAR_BWD), ph ^ 1u);
      tc::fence_after_sync();
      tc::tmem_ld32(tm_dh0 + 64 * net + lane_base + f0, dz1);
#pragma unroll
      for (int i = 0; i < 32; ++i) dz1[i] *= fmaf(-h1[i], h1[i], 1.0f);
    }
    // ---- first layer of this tile
    float x[POL_IN_PAD];
#pragma unroll
    for (int c = 0; c < POL_IN_PAD; ++c) x[c] = xn[c];
    const long long row = rown;
    const bool valid = validn;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int f = f0 + 4 * g;
      const float4 b = lds4(sw + S2_B1 + f);
      float2 a01 = make_float2(b.x, b.y), a23 = make_float2(b.z, b.w);
#pragma unroll
      for (int c = 0; c < POL_IN_PAD; ++c) {
        const float4 w = lds4(sw + S2_W1T + c * 64 + f);
        const float2 xx = make_float2(x[c], x[c]);
        a01 = __ffma2_rn(make_float2(w.x, w.y), xx, a01);
        a23 = __ffma2_rn(make_float2(w.z, w.w), xx, a23);
      }
      h1[4 * g] = tanh_fast(a01.x); h1[4 * g + 1] = tanh_fast(a01.y);
      h1[4 * g + 2] = tanh_fast(a23.x); h1[4 * g + 3] = tanh_fast(a23.y);
    }
    // ---- operand rows: dz1(t-1) once the weight-gradient MMAs of t-1 have retired, h1(t), aux(t)
    if (it > 0) {
      mbar_wait(P.bar(BAR_WG), ph ^ 1u);
#pragma unroll
      for (int c = 0; c < 4; ++c) store_split_chunk(P.dz(0, net), P.dz(1, net), s, 4 * half + c, dz1 + 8 * c);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) store_split_chunk(P.h1(0, net), P.h1(1, net), s, 4 * half + c, h1 + 8 * c);
    if (q == 0) {
      unsigned char* ah = P.aux(ph, 0);
      unsigned char* am = P.aux(ph, 1);
#pragma unroll
      for (int c = 0; c < POL_IN_PAD; ++c) {
        const __nv_bfloat16 hb = __float2bfloat16_rn(x[c]);
        const __nv_bfloat16 mb = __float2bfloat16_rn(x[c] - __bfloat162float(hb));
        const int off = aux_off(1 + c, s);
        *reinterpret_cast<unsigned short*>(ah + off) = __bfloat16_as_ushort(hb);
        *reinterpret_cast<unsigned short*>(am + off) = __bfloat16_as_ushort(mb);
      }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 128) {                                  // a thread that idles through the loss issues this batch
      tc::fence_after_sync();
      for (int n = 0; n < 2; ++n)
        mma_split(tm_z0 + 64 * n, tc::smem_desc_k_sw128(P.h1(0, n)), tc::smem_desc_k_sw128(P.h1(1, n)),
                  tc::smem_desc_k_sw128(P.w2(n, 0)), tc::smem_desc_k_sw128(P.w2(n, 1)), ID_FWD, 4, 2, 2, false);
      tc::mma_commit(P.bar(BAR_FWD));
      if (it > 0)
        mma_aux(tm_w1, tc::smem_desc_mn_sw128(P.dz(0, 0), T2_TILE, 1024), tc::smem_desc_mn_sw128(P.dz(1, 0), T2_TILE, 1024),
                tc::smem_desc_k_sw128(P.aux(ph ^ 1u, 0)), tc::smem_desc_k_sw128(P.aux(ph ^ 1u, 1)), ID_AUX, it > 1);
      tc::mma_commit(P.bar(BAR_AUX));
    }
    // ---- gather of the next tile (consumed one iteration later)
    prefetch_tile(tile + gridDim.x);
    // ---- second layer, head, loss
    mbar_wait(P.bar(BAR_FWD), ph);
    tc::fence_after_sync();
    float h2[32];
    tc::tmem_ld32(tm_z0 + 64 * net + lane_base + f0, h2);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 b = lds4(sw + S2_B2 + f0 + 4 * g);
      h2[4 * g] = tanh_fast(h2[4 * g] + b.x); h2[4 * g + 1] = tanh_fast(h2[4 * g + 1] + b.y);
      h2[4 * g + 2] = tanh_fast(h2[4 * g + 2] + b.z); h2[4 * g + 3] = tanh_fast(h2[4 * g + 3] + b.w);
    }
#pragma unroll
    for (int k = 0; k < POL_OUT_MAX; ++k) {
      if (k < OUT) {
        float p0 = 0.0f, p1 = 0.0f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 w = lds4(sw + S2_W3 + k * 64 + f0 + 4 * g);
          p0 = fmaf(w.x, h2[4 * g], p0); p1 = fmaf(w.y, h2[4 * g + 1], p1);
          p0 = fmaf(w.z, h2[4 * g + 2], p0); p1 = fmaf(w.w, h2[4 * g + 3], p1);
        }
        stg[(half * 4 + k) * T2_S + s] = p0 + p1;       // head exchange: [half][k][sample]
      }
    }
    bar_sync_named(barid, 256);
    if (half == 0) {
      float out[POL_OUT_MAX], dout[POL_OUT_MAX];
#pragma unroll
      for (int k = 0; k < POL_OUT_MAX; ++k) {
        out[k] = k < OUT ? (stg[k * T2_S + s] + stg[(4 + k) * T2_S + s]) + sw[S2_B3 + k] : 0.0f;
        dout[k] = 0.0f;
      }
      if (valid) {
        if (net == 0) {
          const float oldlp = __ldg(a.logprobs + row), adv = __ldg(a.advantages + row);
          float newlogp, entropy, dlp[POL_OUT_MAX], dH[POL_OUT_MAX];
          float g_ls[POL_OUT_MAX] = {0.f, 0.f, 0.f, 0.f};
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
              g_ls[k] = k < OUT ? d * d / var - 1.0f : 0.0f;
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
            for (int k = 0; k < POL_OUT_MAX; ++k) if (k < OUT) st[5 + k] += g_logp * g_ls[k] + g_H;
          }
          st[0] += fmaxf(l1, l2); st[1] += entropy; st[2] += -logr; st[3] += (ratio - 1.0f) - logr;
          st[4] += fabsf(ratio - 1.0f) > a.clip ? 1.0f : 0.0f;
        } else {
          const float R = __ldg(a.returns + row), vold = __ldg(a.values + row), v = out[0];
          if (a.clip_vloss) {
            const float du = v - R, vu = du * du;
            const float d = v - vold, vc = vold + fminf(fmaxf(d, -a.clip), a.clip);
            const float dc = vc - R, lc = dc * dc;
            const float w1 = vu > lc ? 1.0f : (vu == lc ? 0.5f : 0.0f);
            const float inr = (d >= -a.clip && d <= a.clip) ? 1.0f : 0.0f;
            dout[0] = (w1 * du + (1.0f - w1) * dc * inr) * a.vf_c * a.inv_m;
            st[0] += 0.5f * fmaxf(vu, lc);
          } else {
            const float d = v - vold;
            dout[0] = d * a.vf_c * a.inv_m;
            st[0] += 0.5f * d * d;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < POL_OUT_MAX; ++k) sdo[k * T2_LD + s] = dout[k];
    }
    bar_sync_named(barid, 256);
    float dout[POL_OUT_MAX];
#pragma unroll
    for (int k = 0; k < POL_OUT_MAX; ++k) dout[k] = sdo[k * T2_LD + s];
    // ---- dz2 = (W3^T dout) * (1 - h2^2): operand rows of the backward MMAs (the dz tiles are free once aux_w1(t-1) retired)
    mbar_wait(P.bar(BAR_AUX), ph);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float dz2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) dz2[e] = 0.0f;
#pragma unroll
      for (int k = 0; k < POL_OUT_MAX; ++k) {
        if (k < OUT) {
          const float4 wa = lds4(sw + S2_W3 + k * 64 + f0 + 8 * c), wb = lds4(sw + S2_W3 + k * 64 + f0 + 8 * c + 4);
          dz2[0] = fmaf(wa.x, dout[k], dz2[0]); dz2[1] = fmaf(wa.y, dout[k], dz2[1]);
          dz2[2] = fmaf(wa.z, dout[k], dz2[2]); dz2[3] = fmaf(wa.w, dout[k], dz2[3]);
          dz2[4] = fmaf(wb.x, dout[k], dz2[4]); dz2[5] = fmaf(wb.y, dout[k], dz2[5]);
          dz2[6] = fmaf(wb.z, dout[k], dz2[6]); dz2[7] = fmaf(wb.w, dout[k], dz2[7]);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) dz2[e] *= fmaf(-h2[8 * c + e], h2[8 * c + e], 1.0f);
      store_split_chunk(P.dz(0, net), P.dz(1, net), s, 4 * half + c, dz2);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 384) {
      tc::fence_after_sync();
      for (int n = 0; n < 2; ++n)
        mma_split(tm_dh0 + 64 * n, tc::smem_desc_k_sw128(P.dz(0, n)), tc::smem_desc_k_sw128(P.dz(1, n)),
                  tc::smem_desc_mn_sw128(P.w2(n, 0), 8192, 1024), tc::smem_desc_mn_sw128(P.w2(n, 1), 8192, 1024), ID_BWD, 4, 2, 128,
                  false);
      tc::mma_commit(P.bar(BAR_BWD));
      // D_w[m][n] (+)= sum_s [dz2_a | dz2_c][s][m] * [h1_a | h1_c][s][n], K = 128 samples = 8 steps of 16 rows
      mma_split(tm_w, tc::smem_desc_mn_sw128(P.dz(0, 0), T2_TILE, 1024), tc::smem_desc_mn_sw128(P.dz(1, 0), T2_TILE, 1024),
                tc::smem_desc_mn_sw128(P.h1(0, 0), T2_TILE, 1024), tc::smem_desc_mn_sw128(P.h1(1, 0), T2_TILE, 1024), ID_WG, 8, 128, 128,
                it > 0);
      mma_aux(tm_b2, tc::smem_desc_mn_sw128(P.dz(0, 0), T2_TILE, 1024), tc::smem_desc_mn_sw128(P.dz(1, 0), T2_TILE, 1024),
              tc::smem_desc_k_sw128(P.aux(ph, 0)), tc::smem_desc_k_sw128(P.aux(ph, 1)), ID_AUX, it > 0);
      tc::mma_commit(P.bar(BAR_WG));
    }
    // ---- dW3[k][j] += sum_s dout[s][k] h2[s][j], db3[k] += sum_s dout[s][k]: two rounds of 32 features through smem
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (half == r) {
#pragma unroll
        for (int i = 0; i < 32; ++i) stg[i * T2_LD + s] = h2[i];
      }
      bar_sync_named(barid, 256);
      const float* hp = stg + ri * T2_LD + 16 * c8;
      const float4 hv[4] = {lds4(hp), lds4(hp + 4), lds4(hp + 8), lds4(hp + 12)};
#pragma unroll
      for (int k = 0; k < POL_OUT_MAX; ++k) {
        if (k < OUT) {
          const float* dp = sdo + k * T2_LD + 16 * c8;
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 dv = lds4(dp + 4 * g);
            s0 = fmaf(hv[g].x, dv.x, s0); s1 = fmaf(hv[g].y, dv.y, s1);
            s0 = fmaf(hv[g].z, dv.z, s0); s1 = fmaf(hv[g].w, dv.w, s1);
          }
          acc_w3[r][k] += s0 + s1;
        }
      }
      if (r == 0) {
        if (ri < OUT) {
          const float* dp = sdo + ri * T2_LD + 16 * c8;
          float sb = 0.0f;
#pragma unroll
          for (int g = 0; g < 4; ++g) { const float4 dv = lds4(dp + 4 * g); sb += (dv.x + dv.y) + (dv.z + dv.w); }
          acc_b3 += sb;
        }
        bar_sync_named(barid, 256);
      }
    }
  }

  // ---- tail: first-layer gradients of the last tile
  if (any) {
    const uint32_t ph = (it - 1u) & 1u;
    float dz1[32];
    mbar_wait(P.bar(BAR_BWD), ph);
    tc::fence_after_sync();
    tc::tmem_ld32(tm_dh0 + 64 * net + lane_base + f0, dz1);
#pragma unroll
    for (int i = 0; i < 32; ++i) dz1[i] *= fmaf(-h1[i], h1[i], 1.0f);
    mbar_wait(P.bar(BAR_WG), ph);
#pragma unroll
    for (int c = 0; c < 4; ++c) store_split_chunk(P.dz(0, net), P.dz(1, net), s, 4 * half + c, dz1 + 8 * c);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 128) {
      tc::fence_after_sync();
      mma_aux(tm_w1, tc::smem_desc_mn_sw128(P.dz(0, 0), T2_TILE, 1024), tc::smem_desc_mn_sw128(P.dz(1, 0), T2_TILE, 1024),
              tc::smem_desc_k_sw128(P.aux(ph, 0)), tc::smem_desc_k_sw128(P.aux(ph, 1)), ID_AUX, it > 1);
      tc::mma_commit(P.bar(BAR_FIN));
    }
    mbar_wait(P.bar(BAR_FIN), 0);
    tc::fence_after_sync();
  }
  __syncthreads();

  // ---- epilogue: this CTA's partial sums in the nets' flat parameter order
  {
    // TMEM accumulators: thread (row m = s, q): net_m = m / 64, feature j = m % 64
    const int net_m = s >> 6, j = s & 63, OUTm = net_m == 0 ? A : 1;
    float* part = a.partials + ((size_t)net_m * gridDim.x + blockIdx.x) * UPD_PSTRIDE;
    const int oB1 = 64 * obs_dim, oW2 = oB1 + 64, oB2 = oW2 + 4096;
    (void)OUTm;
    {
      uint32_t v[16];
      if (any) tc::tmem_ld16(tm_w + lane_base + (uint32_t)(net_m * 64 + 16 * q), v);
#pragma unroll
      for (int i = 0; i < 16; ++i) part[oW2 + j * 64 + 16 * q + i] = any ? __uint_as_float(v[i]) : 0.0f;
    }
    if (q == 0) {
      uint32_t v[16];
      if (any) tc::tmem_ld16(tm_b2 + lane_base, v);
      part[oB2 + j] = any ? __uint_as_float(v[0]) : 0.0f;
    }
    if (q == 1) {
      uint32_t v[16];
      if (any) tc::tmem_ld16(tm_w1 + lane_base, v);
      part[oB1 + j] = any ? __uint_as_float(v[0]) : 0.0f;
#pragma unroll
      for (int c = 0; c < POL_IN_PAD; ++c)
        if (c < obs_dim) part[j * obs_dim + c] = any ? __uint_as_float(v[1 + c]) : 0.0f;
    }
  }
  {
    // SIMT accumulators of this thread's net: dW3 / db3 summed over the 8 sample slices, statistics over the warps
    float* part = a.partials + ((size_t)net * gridDim.x + blockIdx.x) * UPD_PSTRIDE;
    const int oW3 = 64 * obs_dim + 64 + 4096 + 64, oB3 = oW3 + OUT * 64, oLS = oB3 + OUT;
    float* red = stg;                                   // [(r*4 + k)][c8][32]  then db3 [4][8], then stats [9][4]
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int k = 0; k < POL_OUT_MAX; ++k) red[((r * 4 + k) * 8 + c8) * 32 + ri] = acc_w3[r][k];
    if (ri < POL_OUT_MAX) red[2048 + ri * 8 + c8] = acc_b3;
    if (half == 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float v = warp_sum(st[i]);
        if ((tid & 31) == 0) red[2048 + 32 + i * 4 + (warp & 3)] = v;
      }
    }
    bar_sync_named(barid, 256);
    {
      const int f = tp & 63, k = tp >> 6;
      if (k < OUT) {
        const int r = f >> 5, i = f & 31;
        float sum = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) sum += red[((r * 4 + k) * 8 + c) * 32 + i];
        part[oW3 + k * 64 + f] = sum;
      }
      if (tp < OUT) {
        float sum = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) sum += red[2048 + tp * 8 + c];
        part[oB3 + tp] = sum;
      }
      if (tp < 9) {
        const float* p = red + 2048 + 32 + tp * 4;
        const float sum = (p[0] + p[1]) + (p[2] + p[3]);
        float* stat = part + UPD_STAT_OFF;
        if (net == 0) {
          if (tp == 0) stat[AUR_STAT_POLICY_LOSS] = sum;
          if (tp == 1) stat[AUR_STAT_ENTROPY] = sum;
          if (tp == 2) stat[AUR_STAT_OLD_APPROX_KL] = sum;
          if (tp == 3) stat[AUR_STAT_APPROX_KL] = sum;
          if (tp == 4) stat[AUR_STAT_CLIPFRAC] = sum;
          if (tp >= 5 && a.continuous && tp - 5 < OUT) part[oLS + tp - 5] = sum;
        } else if (tp == 0) {
          stat[AUR_STAT_VALUE_LOSS] = sum;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

size_t ppo_grad_tc_smem_bytes() { return T2_SMEM; }

int launch_ppo_grad_tc(const UpdDev& d, int gx, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    AUR_CUDA_OK(cudaFuncSetAttribute(ppo_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM));
    attr = true;
  }
  ppo_grad_tc_kernel<<<gx, T2_THREADS, T2_SMEM, s>>>(d);
  AUR_LAUNCH_OK("ppo_grad_tc_kernel");
  return 0;
}

}  // namespace aur
