// GAE / returns reverse scan over the [T,N] rollout buffers (row G of the scope table).
// Replaces ppo.run_gae (src/ppo.py:125-142) and ppo.normal_advantage (src/ppo.py:145-157).
//
// HBM-bound: 3 reads + 2 writes of fp32 per (t, env) = 20 B, a T-long dependent chain per
// column.  N columns alone cannot hide DRAM latency (65536 columns = 14 warps/SM), so the
// bandwidth comes from depth, not occupancy: a persistent CTA owns a tile of GAE_W adjacent
// columns and a producer warp streams [GAE_TT rows x GAE_W cols] chunks of rewards / values /
// terminals, newest rows first, through a GAE_STAGES-deep shared-memory ring with 1-D bulk
// async copies (cp.async.bulk -> UBLKCP, completion on mbarriers).  Consumer threads own one
// column each, read the staged rows conflict-free, and store advantages / returns straight
// from registers (one full 128-B line per warp per row).
//
// Arithmetic: fp32, one rounding per reference operation, explicit __fmul_rn/__fadd_rn so
// nvcc cannot contract: bit-identical to the reference's torch CPU loop.
#include <stdlib.h>

#include "common.cuh"

namespace aur {

// Tile shape: GAE_W columns per tile, GAE_TT rows per stage, GAE_STAGES ring depth (template parameters;
// the default is picked in aur_gae_f32, AUR_GAE_VARIANT overrides it for tuning sweeps).

struct GaeStep {
  float g32, gl32;
  int use_gae;
  // state carried down the column
  float v_next, nnt_next, last;
  __device__ __forceinline__ void init(float next_value, float next_done) {
    v_next = next_value;
    nnt_next = __fsub_rn(1.0f, next_done);
    last = use_gae ? 0.0f : next_value;   // lastgaelam = 0 (ppo.py:127) / next_return = next_value (ppo.py:151)
  }
  // one reference loop iteration (ppo.py:129-140 / ppo.py:149-155)
  __device__ __forceinline__ void step(float r, float v, float term, float& adv, float& ret) {
    if (use_gae) {
      float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g32, v_next), nnt_next)), v);
      adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl32, nnt_next), last));
      ret = __fadd_rn(adv, v);
      last = adv;
    } else {
      ret = __fadd_rn(r, __fmul_rn(__fmul_rn(g32, nnt_next), last));
      adv = __fsub_rn(ret, v);
      last = ret;
    }
    v_next = v;
    nnt_next = __fsub_rn(1.0f, term);
  }
};

template <int GAE_W, int GAE_TT, int GAE_STAGES>
__global__ void __launch_bounds__(GAE_W + 32)
gae_bulk_kernel(int T, long long N, const float* __restrict__ rew, const float* __restrict__ val,
                const float* __restrict__ term, const float* __restrict__ next_value,
                const float* __restrict__ next_done, float g32, float gl32, int use_gae,
                float* __restrict__ adv_out, float* __restrict__ ret_out) {
  constexpr int GAE_CONSUMER_WARPS = GAE_W / 32;
  constexpr int GAE_STAGE_FLOATS = 3 * GAE_TT * GAE_W;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(float) * GAE_STAGES * GAE_STAGE_FLOATS);
  uint64_t* empty = full + GAE_STAGES;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (N + GAE_W - 1) / GAE_W;
  const int nchunks = (T + GAE_TT - 1) / GAE_TT;

  if (tid == 0) {
    for (int s = 0; s < GAE_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GAE_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == GAE_CONSUMER_WARPS) {
    // ===== producer warp: stream chunks, newest rows first =====
    unsigned it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long c0 = tile * GAE_W;
      const int wcols = (int)min((long long)GAE_W, N - c0);
      const uint32_t row_bytes = (uint32_t)wcols * 4u;
      for (int k = 0; k < nchunks; ++k, ++it) {
        const int s = it % GAE_STAGES;
        const uint32_t ph = (it / GAE_STAGES) & 1u;
        const int t_hi = T - k * GAE_TT;
        const int t_lo = max(0, t_hi - GAE_TT);
        const int nrows = t_hi - t_lo;
        mbar_wait(&empty[s], ph ^ 1u);
        if (lane == 0) mbar_arrive_expect_tx(&full[s], row_bytes * 3u * (uint32_t)nrows);
        __syncwarp();
        float* st = ring + (size_t)s * GAE_STAGE_FLOATS;
        for (int q = lane; q < 3 * nrows; q += 32) {
          const int a = q / nrows, j = q - a * nrows;
          const float* src = (a == 0 ? rew : (a == 1 ? val : term)) + (size_t)(t_lo + j) * (size_t)N + c0;
          bulk_g2s(st + (a * GAE_TT + j) * GAE_W, src, row_bytes, &full[s]);
        }
      }
    }
  } else {
    // ===== consumers: one column per thread =====
    GaeStep gs;
    gs.g32 = g32; gs.gl32 = gl32; gs.use_gae = use_gae;
    unsigned it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long col = tile * GAE_W + tid;
      const bool active = col < N;
      gs.init(active ? __ldg(next_value + col) : 0.0f, active ? __ldg(next_done + col) : 0.0f);
      for (int k = 0; k < nchunks; ++k, ++it) {
        const int s = it % GAE_STAGES;
        const uint32_t ph = (it / GAE_STAGES) & 1u;
        const int t_hi = T - k * GAE_TT;
        const int t_lo = max(0, t_hi - GAE_TT);
        const int nrows = t_hi - t_lo;
        mbar_wait(&full[s], ph);
        const float* st = ring + (size_t)s * GAE_STAGE_FLOATS;
        if (active) {
          if (nrows == GAE_TT) {
            float r[GAE_TT], v[GAE_TT], d[GAE_TT];
#pragma unroll
            for (int j = 0; j < GAE_TT; ++j) {
              r[j] = st[(0 * GAE_TT + j) * GAE_W + tid];
              v[j] = st[(1 * GAE_TT + j) * GAE_W + tid];
              d[j] = st[(2 * GAE_TT + j) * GAE_W + tid];
            }
#pragma unroll
            for (int j = GAE_TT - 1; j >= 0; --j) {
              float adv, ret;
              gs.step(r[j], v[j], d[j], adv, ret);
              const size_t o = (size_t)(t_lo + j) * (size_t)N + (size_t)col;
              __stcs(adv_out + o, adv);
              __stcs(ret_out + o, ret);
            }
          } else {
            for (int j = nrows - 1; j >= 0; --j) {
              float adv, ret;
              gs.step(st[(0 * GAE_TT + j) * GAE_W + tid], st[(1 * GAE_TT + j) * GAE_W + tid],
                      st[(2 * GAE_TT + j) * GAE_W + tid], adv, ret);
              const size_t o = (size_t)(t_lo + j) * (size_t)N + (size_t)col;
              __stcs(adv_out + o, adv);
              __stcs(ret_out + o, ret);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
    }
  }
}

// Plain column kernel for shapes the bulk path cannot take (N % 4 != 0, unaligned
// pointers): one column per thread, rows prefetched 8 deep into registers.
__global__ void __launch_bounds__(128)
gae_column_kernel(int T, long long N, const float* __restrict__ rew, const float* __restrict__ val,
                  const float* __restrict__ term, const float* __restrict__ next_value,
                  const float* __restrict__ next_done, float g32, float gl32, int use_gae,
                  float* __restrict__ adv_out, float* __restrict__ ret_out) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  GaeStep gs;
  gs.g32 = g32; gs.gl32 = gl32; gs.use_gae = use_gae;
  gs.init(next_value[col], next_done[col]);
  constexpr int U = 8;
  int t = T;
  while (t > 0) {
    const int nrows = min(U, t);
    float r[U], v[U], d[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (j < nrows) {
        const size_t o = (size_t)(t - 1 - j) * (size_t)N + (size_t)col;
        r[j] = __ldg(rew + o); v[j] = __ldg(val + o); d[j] = __ldg(term + o);
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (j < nrows) {
        float adv, ret;
        gs.step(r[j], v[j], d[j], adv, ret);
        const size_t o = (size_t)(t - 1 - j) * (size_t)N + (size_t)col;
        adv_out[o] = adv; ret_out[o] = ret;
      }
    }
    t -= nrows;
  }
}

template <int W, int TT, int STAGES>
static int launch_gae_bulk(int32_t T, int64_t N, const float* rewards, const float* values, const float* terminals,
                           const float* next_value, const float* next_done, float g32, float gl32, int use_gae,
                           float* adv_out, float* ret_out, cudaStream_t st) {
  constexpr size_t smem = sizeof(float) * STAGES * 3 * TT * W + 2 * STAGES * sizeof(uint64_t);
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(gae_bulk_kernel<W, TT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.done();
  }
  const long long ntiles = (N + W - 1) / W;
  int per_sm = (int)(220 * 1024 / (smem + 1024));
  const int by_threads = 2048 / (W + 32);
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm < 1) per_sm = 1;
  const long long cap = (long long)sm_count() * per_sm;
  const long long grid = ntiles < cap ? ntiles : cap;
  gae_bulk_kernel<W, TT, STAGES><<<(unsigned)grid, W + 32, smem, st>>>(T, (long long)N, rewards, values, terminals, next_value,
                                                                     next_done, g32, gl32, use_gae, adv_out, ret_out);
  AUR_LAUNCH_OK("gae_bulk_kernel");
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int gae_kind(int32_t T, int64_t N, const float* rew, const float* val, const float* term,
                    const float* adv, const float* ret) {
  (void)T; (void)adv; (void)ret;
  return (N % 4 == 0 && N >= 64 && aligned16(rew) && aligned16(val) && aligned16(term)) ? 1 : 0;
}

}  // namespace aur

extern "C" int aur_gae_kernel_kind(int32_t T, int64_t N, const float* rewards, const float* values,
                                   const float* terminals, const float* adv_out, const float* ret_out) {
  return aur::gae_kind(T, N, rewards, values, terminals, adv_out, ret_out);
}

extern "C" int aur_gae_f32(int32_t T, int64_t N, const float* rewards, const float* values, const float* terminals,
                           const float* next_value, const float* next_done, double gamma, double gae_lambda,
                           int32_t use_gae, float* adv_out, float* ret_out, void* stream) {
  using namespace aur;
  if (T < 0 || N < 0) { set_error("aur_gae_f32: negative T or N"); return AUR_ERR_ARG; }
  if (T == 0 || N == 0) return 0;
  if (!rewards || !values || !terminals || !next_value || !next_done || !adv_out || !ret_out) {
    set_error("aur_gae_f32: null pointer"); return AUR_ERR_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float g32 = (float)gamma;                 // torch: tensor * python-float -> scalar cast to fp32
  const float gl32 = (float)(gamma * gae_lambda); // python computes gamma*gae_lambda in double first
  if (gae_kind(T, N, rewards, values, terminals, adv_out, ret_out)) {
    static int variant = -1;
    if (variant < 0) {
      const char* e = getenv("AUR_GAE_VARIANT");
      variant = e ? atoi(e) : 1;    // <64,8,4>: best of the round-1 sweep (tools/sweep_gae.sh) at [128,65536]..[2048,131072]
    }
    int rc = 0;
    switch (variant) {
      case 7: rc = launch_gae_bulk<128, 8, 4>(T, N, rewards, values, terminals, next_value, next_done, g32, gl32, use_gae, adv_out, ret_out, st); break;
      case 2: rc = launch_gae_bulk<64, 16, 3>(T, N, rewards, values, terminals, next_value, next_done, g32, gl32, use_gae, adv_out, ret_out, st); break;
      case 3: rc = launch_gae_bulk<128, 16, 3>(T, N, rewards, values, terminals, next_value, next_done, g32, gl32, use_gae, adv_out, ret_out, st); break;
      case 4: rc = launch_gae_bulk<128, 4, 6>(T, N, rewards, values, terminals, next_value, next_done, g32, gl32, use_gae, adv_out, ret_out, st); break;
      case 5: rc = launch_gae_bulk<64, 4, 8>(T, N, rewards, values, terminals, next_value, next_done, g32, gl32, use_gae, adv_out, ret_out, st); break;
      case 6: rc = launch_gae_bulk<256, 4, 4>(T, N, rewards, values, terminals, next_value, next_done, g32, gl32, use_gae, adv_out, ret_out, st); break;
      default: rc = launch_gae_bulk<64, 8, 4>(T, N, rewards, values, terminals, next_value, next_done, g32, gl32, use_gae, adv_out, ret_out, st); break;
    }
    if (rc) return rc;
  } else {
    const unsigned grid = (unsigned)((N + 127) / 128);
    gae_column_kernel<<<grid, 128, 0, st>>>(T, (long long)N, rewards, values, terminals, next_value, next_done, g32,
                                            gl32, use_gae, adv_out, ret_out);
    AUR_LAUNCH_OK("gae_column_kernel");
  }
  return 0;
}
