// Batched critic forward on tensor cores: values[i] = critic(obs[i]) for a whole [T*N] observation buffer.
// In the reference the critic is evaluated inside the rollout loop (actor_critic.evaluate, ppo.py:105, and
// policy.value, ppo.py:161), but a value never feeds back into the trajectory (only the actor's action does),
// so the fused rollout kernel keeps just the actor + physics on its sequential path and the T*N values are one
// embarrassingly parallel pass over the observation rows it has just written (contiguous, no gather).
//
// Same scheme as update_tc.cu's forward: CTA = 128 samples x 2 feature halves, first layer in registers, h1 rows
// written as a THREE-term bf16 split (hi + mid + lo, fp32-equivalent: values are compared with the reference at
// 2e-6) into 128-B-swizzled tiles, z2 = h1 W2^T on tcgen05 as six products with fp32 accumulation in TMEM, tanh +
// 64->1 head from TMEM lanes.
// Three CTAs per SM (75 KB of shared memory each) cover each other's MMA round trip.  HBM: 16 + 4 B per sample.
#include "tc_split.cuh"
#include "policy.cuh"

namespace aur {

constexpr int VT_S = 128, VT_THREADS = 256;
constexpr int VT_TILE = VT_S * 128, VT_WTILE = 64 * 128;
constexpr int VO_H1 = 0;                         // [hi][mid][lo]
constexpr int VO_W2 = VO_H1 + 3 * VT_TILE;       // [hi][mid][lo]
constexpr int VO_SMALL = VO_W2 + 3 * VT_WTILE;   // W1^T [4][64], b1 [64], b2 [64], W3 [64], b3 (452 floats)
constexpr int VO_BAR = VO_SMALL + 452 * 4;
constexpr size_t VT_SMEM = VO_BAR + 16 + 1024;
static_assert(3 * (VT_SMEM + 1024) <= 233472, "three CTAs per SM");
constexpr int VS_W1T = 0, VS_B1 = 256, VS_B2 = 320, VS_W3 = 384, VS_B3 = 448;

__global__ void __launch_bounds__(VT_THREADS, 3) critic_values_tc_kernel(const float* __restrict__ critic, int obs_dim,
                                                                         const float* __restrict__ obs, long long M,
                                                                         float* __restrict__ out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* h1t[3] = {base + VO_H1, base + VO_H1 + VT_TILE, base + VO_H1 + 2 * VT_TILE};
  unsigned char* w2t[3] = {base + VO_W2, base + VO_W2 + VT_WTILE, base + VO_W2 + 2 * VT_WTILE};
  float* sw = reinterpret_cast<float*>(base + VO_SMALL);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + VO_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + VO_BAR + 8);
  const int tid = threadIdx.x, warp = tid >> 5, s = tid & 127, half = tid >> 7, f0 = 32 * half;

  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc(tslot, 64);
  {
    const float* gb1 = critic + 64 * obs_dim;
    const float* gW2 = gb1 + 64;
    const float* gb2 = gW2 + 4096;
    const float* gW3 = gb2 + 64;
    const int c = tid >> 6, j = tid & 63;
    sw[VS_W1T + tid] = c < obs_dim ? TANH_PRESCALE * critic[j * obs_dim + c] : 0.0f;      // tanh argument scale folded in
    if (tid < 64) { sw[VS_B1 + tid] = TANH_PRESCALE * gb1[tid]; sw[VS_B2 + tid] = TANH_PRESCALE * gb2[tid]; sw[VS_W3 + tid] = gW3[tid]; }
    if (tid == 0) sw[VS_B3] = gW3[64];
#pragma unroll 1
    for (int ch = 2 * c; ch < 2 * c + 2; ++ch) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = gW2[j * 64 + 8 * ch + e];
      store_split3_chunk(w2t[0], w2t[1], w2t[2], j, ch, v);
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tm_z = *tslot;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  constexpr uint32_t ID_FWD = tc::instr_desc(tc::FMT_BF16, 128, 64, 0, 0);
  const bool vec = obs_dim == 4 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0;

  const long long ntiles = (M + VT_S - 1) / VT_S;
  auto load_obs = [&](long long tile, float (&x)[POL_IN_PAD]) {
    const long long row = tile * VT_S + s;
    const bool ok = tile < ntiles && row < M;
    if (ok && vec) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(obs) + row);
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
      for (int c = 0; c < POL_IN_PAD; ++c) x[c] = (ok && c < obs_dim) ? __ldg(obs + row * obs_dim + c) : 0.0f;
    }
  };
  float xn[POL_IN_PAD];
  load_obs(blockIdx.x, xn);
  uint32_t it = 0;
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    // ---- first layer -> operand rows (8-feature chunks)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float hv[8];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int f = f0 + 8 * c + 4 * g;
        const float4 b = lds4(sw + VS_B1 + f);
        float2 a01 = make_float2(b.x, b.y), a23 = make_float2(b.z, b.w);
#pragma unroll
        for (int cc = 0; cc < POL_IN_PAD; ++cc) {
          const float4 w = lds4(sw + VS_W1T + cc * 64 + f);
          const float2 xx = make_float2(xn[cc], xn[cc]);
          a01 = __ffma2_rn(make_float2(w.x, w.y), xx, a01);
          a23 = __ffma2_rn(make_float2(w.z, w.w), xx, a23);
        }
        hv[4 * g] = tanh_prescaled(a01.x); hv[4 * g + 1] = tanh_prescaled(a01.y);
        hv[4 * g + 2] = tanh_prescaled(a23.x); hv[4 * g + 3] = tanh_prescaled(a23.y);
      }
      store_split3_chunk(h1t[0], h1t[1], h1t[2], s, 4 * half + c, hv);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint64_t da[3] = {tc::smem_desc_k_sw128(h1t[0]), tc::smem_desc_k_sw128(h1t[1]), tc::smem_desc_k_sw128(h1t[2])};
      const uint64_t db[3] = {tc::smem_desc_k_sw128(w2t[0]), tc::smem_desc_k_sw128(w2t[1]), tc::smem_desc_k_sw128(w2t[2])};
      mma_split6(tm_z, da, db, ID_FWD, 4, 2, 2);
      tc::mma_commit(bar);
    }
    load_obs(tile + gridDim.x, xn);                  // next tile's rows, in flight across the MMA round trip
    mbar_wait(bar, it & 1u);
    tc::fence_after_sync();
    float z[32];
    tc::tmem_ld32(tm_z + lane_base + f0, z);
    float p0 = 0.0f, p1 = 0.0f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 b = lds4(sw + VS_B2 + f0 + 4 * g), w = lds4(sw + VS_W3 + f0 + 4 * g);
      p0 = fmaf(w.x, tanh_prescaled(fmaf(z[4 * g], TANH_PRESCALE, b.x)), p0);
      p1 = fmaf(w.y, tanh_prescaled(fmaf(z[4 * g + 1], TANH_PRESCALE, b.y)), p1);
      p0 = fmaf(w.z, tanh_prescaled(fmaf(z[4 * g + 2], TANH_PRESCALE, b.z)), p0);
      p1 = fmaf(w.w, tanh_prescaled(fmaf(z[4 * g + 3], TANH_PRESCALE, b.w)), p1);
    }
    // the upper half hands its head partial over through a TMEM column of its own (already consumed) z range
    if (half == 1) { tc::tmem_st1(tm_z + lane_base + 32, p0 + p1); tc::tmem_wait_st(); }
    tc::fence_before_sync();
    __syncthreads();                                  // head partials visible; the h1 tile is free again
    if (half == 0) {
      tc::fence_after_sync();
      const float other = tc::tmem_ld1(tm_z + lane_base + 32);
      const long long row = tile * VT_S + s;
      if (row < M) out[row] = ((p0 + p1) + other) + sw[VS_B3];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm_z, 64);
}

// values of M observation rows with the critic of a (hidden 64, 2 layers) policy; `critic` points at the critic's
// first parameter inside the flat buffer
int launch_critic_values_tc(const float* critic, int obs_dim, const float* obs, long long M, float* out, cudaStream_t s) {
  if (M <= 0) return 0;
  static DeviceOnce attr;
  if (attr.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(critic_values_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VT_SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(critic_values_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr.done();
  }
  const long long ntiles = (M + VT_S - 1) / VT_S;
  long long grid = 3LL * sm_count();
  if (grid > ntiles) grid = ntiles;
  critic_values_tc_kernel<<<(unsigned)grid, VT_THREADS, VT_SMEM, s>>>(critic, obs_dim, obs, M, out);
  AUR_LAUNCH_OK("critic_values_tc_kernel");
  return 0;
}

}  // namespace aur
