// Backward pass and small layers of the equivariant actor-critic update (row X):
//   wgrad3x3_kernel      weight gradient of a 3x3 layer as a split-K tcgen05 GEMM over the FLAT haloed
//                        pixel index: dW[co][tap][ci] = sum_q dY_cm[co][q] * X_cm[ci][q + off(tap)]
//   unpool_relu_bwd      max-pool + ReLU backward (routes to the stored arg-max, masks by output > 0)
//   transpose_bf16       NHWC [Q][C] -> channel-major [C][Q] copies feeding the wgrad GEMM
//   project_regular      dWmat -> dpsi (adjoint of the p4 filter expansion) and per-field bias gradients
//   conv0_wgrad          layer-0 weight gradient fused with its un-pooling (direct, fp32)
//   heads / loss         actor head decode + Normal log-prob / entropy + PPO loss seeds (robot_ppo.py:345-398),
//                        critic head ReLU + GroupPooling + value and its backward
//   adam_flat / sumsq    torch Adam math on large flat buffers, global-norm clipping
#include <stdlib.h>

#include "tc.cuh"
#include "policy.cuh"

namespace aur {
namespace tc {

int launch_conv0_wgrad_tc(const float* obs, const float* state, const void* da1, const void* a1, const uint8_t* arg, int B,
                          float* dw0, float* dbias_ch, cudaStream_t s, bool plain = false, int parts = 2);

constexpr int WG_BM = 128, WG_BN = 128, WG_BK = 64, WG_STAGES = 3, WG_TAPS = 3;
constexpr int WG_A_BYTES = WG_BM * WG_BK * 2, WG_B_BYTES = WG_BN * WG_BK * 2;
constexpr size_t WG_SMEM = (size_t)WG_STAGES * (WG_A_BYTES + WG_TAPS * WG_B_BYTES) + 1024 + 256;

// grid: x = (Cout tiles) * (Cin tiles), y = tap ROW dy, z = split-K slice.  A CTA computes the three taps dx = 0, 1, 2 of
// its row from ONE load of the output-gradient tile per stage (three shifted input tiles, three TMEM accumulators): the
// gradient operand is fetched 3x instead of 9x, 85 B of operands per MMA cycle instead of 128.
// Operands come STRAIGHT from the NHWC buffers (pixel rows, channels contiguous) as MN-major UMMA
// operands: a stage holds, per operand, two TMA boxes of [64 pixel rows][64 channels] (128-B rows,
// 128-B swizzle); 8 pixel rows form one 1024-B swizzle atom, the two channel halves are 8192 B apart
// (LBO), successive groups of 8 pixel rows 1024 B apart (SBO).  The tap shift is a shift of the PIXEL
// (row) coordinate of the input map, which TMA takes at element granularity.
__global__ void __launch_bounds__(256, 1)
wgrad3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int Cout, int Cin,
                long long Q, int base_off, int Wb, int n_tiles, float* __restrict__ dwmat, int nterm) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
  unsigned char* sA = smem;
  unsigned char* sB = smem + WG_STAGES * WG_A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + WG_STAGES * WG_TAPS * WG_B_BYTES);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* tmem_full = empty + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x / n_tiles, nt = blockIdx.x - mt * n_tiles;
  const int m0 = mt * WG_BM, n0 = nt * WG_BN;
  const int dy = blockIdx.y;
  const int off = base_off + dy * Wb;                  // + dx for the three taps of this row
  const long long nkb_total = (Q + WG_BK - 1) / WG_BK;
  const long long per = (nkb_total + gridDim.z - 1) / gridDim.z;
  const long long kb_lo = (long long)blockIdx.z * per;
  long long kb_hi = kb_lo + per;
  if (kb_hi > nkb_total) kb_hi = nkb_total;
  // split operand planes: the (hi,hi), (hi,mid), (mid,hi) products are extra K steps over the same pixel block
  const int nkb = (kb_hi > kb_lo ? (int)(kb_hi - kb_lo) : 0) * nterm;
  const bool c64 = Cin <= 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);          // 3 x 128 accumulator columns (power of two)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  if (nkb > 0) {
    if (warp == 0 && lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % WG_STAGES;
        const uint32_t ph = (i / WG_STAGES) & 1u;
        const int term = i % nterm, pa = term_plane_a(term), pb = term_plane_b(term);
        const long long k0 = (kb_lo + i / nterm) * WG_BK;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], WG_A_BYTES + WG_TAPS * (c64 ? WG_B_BYTES / 2 : WG_B_BYTES));
        unsigned char* a = sA + s * WG_A_BYTES;
        tma_load_3d(a, &tmA, m0, (int)k0, pa, &full[s]);
        tma_load_3d(a + 8192, &tmA, m0 + 64, (int)k0, pa, &full[s]);
        for (int dx = 0; dx < WG_TAPS; ++dx) {
          if (c64) {            // one 64-channel box per tap, the three taps 8192 B apart: ONE N = 192 operand
            tma_load_3d(sB + s * WG_TAPS * WG_B_BYTES + dx * 8192, &tmB, 0, (int)(k0 + off + dx), pb, &full[s]);
          } else {
            unsigned char* b = sB + (s * WG_TAPS + dx) * WG_B_BYTES;
            tma_load_3d(b, &tmB, n0, (int)(k0 + off + dx), pb, &full[s]);
            tma_load_3d(b + 8192, &tmB, n0 + 64, (int)(k0 + off + dx), pb, &full[s]);
          }
        }
      }
    } else if (warp == 1 && lane == 0) {
      constexpr uint32_t idesc = instr_desc(FMT_BF16, WG_BM, WG_BN, 1, 1);     // both operands MN-major
      for (int i = 0; i < nkb; ++i) {
        const int s = i % WG_STAGES;
        const uint32_t ph = (i / WG_STAGES) & 1u;
        mbar_wait(&full[s], ph);
        fence_after_sync();
        const uint64_t ad = smem_desc_mn_sw128(sA + s * WG_A_BYTES, 8192, 1024);
        if (c64) {
          // Cin <= 64: the three taps' input tiles are three 64-channel groups of one MN-major operand (LBO = 8192 B), so
          // one N = 192 MMA per 16 pixel rows replaces three N = 128 ones (half of whose columns were zero padding) and
          // the gradient tile is read from shared memory once instead of three times
          constexpr uint32_t idesc192 = instr_desc(FMT_BF16, WG_BM, 192, 1, 1);
          const uint64_t bd = smem_desc_mn_sw128(sB + s * WG_TAPS * WG_B_BYTES, 8192, 1024);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k)
            mma_f16(tmem_d, ad + (uint64_t)(128 * k), bd + (uint64_t)(128 * k), idesc192, (i | k) != 0);
          mma_commit(&empty[s]);
          continue;
        }
        {
          // taps dx = 0, 1 are adjacent in the stage (four 64-channel groups, 8192 B apart): one N = 256 MMA feeds both
          // accumulators from a single read of the gradient tile; tap 2 follows as N = 128
          constexpr uint32_t idesc256 = instr_desc(FMT_BF16, WG_BM, 256, 1, 1);
          const uint64_t bd01 = smem_desc_mn_sw128(sB + (s * WG_TAPS) * WG_B_BYTES, 8192, 1024);
          const uint64_t bd2 = smem_desc_mn_sw128(sB + (s * WG_TAPS + 2) * WG_B_BYTES, 8192, 1024);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k) {    // 16 pixel rows per MMA = 2048 B = 128 x 16 B
            mma_f16(tmem_d, ad + (uint64_t)(128 * k), bd01 + (uint64_t)(128 * k), idesc256, (i | k) != 0);
            mma_f16(tmem_d + 2 * WG_BN, ad + (uint64_t)(128 * k), bd2 + (uint64_t)(128 * k), idesc, (i | k) != 0);
          }
        }
        mma_commit(&empty[s]);
      }
      mma_commit(tmem_full);
    } else if (warp >= 4) {
      const int q = warp - 4;
      mbar_wait(tmem_full, 0);
      fence_after_sync();
      const int co = m0 + 32 * q + lane;
      const int ncol = c64 ? 64 : WG_BN;                // accumulator columns per tap
#pragma unroll 1
      for (int cc = 0; cc < WG_TAPS * ncol; cc += 32) {
        const int dx = cc / ncol, c = cc - dx * ncol, tap = dy * 3 + dx;
        float v[32];
        tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)cc, v);
        if (co < Cout) {
          float* dst = dwmat + ((size_t)co * 9 + tap) * Cin + n0 + c;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + c + i < Cin) atomicAdd(dst + i, v[i]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, 512);
}

// ---- max-pool(2) + ReLU backward: one thread per pooled pixel x 8 channels ------------------
// colsum (nullable): the bias gradient of the layer, out[c / group] += sum over all written positions of dy[.., c] - what
// colsum_bf16_kernel would compute from dy, taken here from registers instead of re-reading the buffer.  The grid is a
// multiple of C / 8 threads (C / 8 divides 256), so a thread keeps one channel group for its whole grid-stride loop.
// planes: every plane of the gradient stack (dpool -> dy, plane strides dp_plane / dy_plane) is routed exactly alike; the ReLU
// mask reads the hi plane of `act` (hi > 0 <=> value > 0).
__global__ void unpool_relu_bwd_kernel(int B, int Hp, int Wp, int C, const __nv_bfloat16* __restrict__ dpool,
                                       const __nv_bfloat16* __restrict__ act, int aHb, int aWb, int aoff,
                                       const unsigned char* __restrict__ arg, __nv_bfloat16* __restrict__ dy, int dHb, int dWb,
                                       int doff, int group, float* __restrict__ colsum, int planes, size_t dp_plane, size_t dy_plane) {
  __shared__ float sacc[1024];
  float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int cg = C >> 3;
  const long long total = (long long)B * Hp * Wp * cg;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long r = e;
    const int c8 = (int)(r % cg); r /= cg;
    const int px = (int)(r % Wp); r /= Wp;
    const int py = (int)(r % Hp); r /= Hp;
    const int b = (int)r;
    const size_t pp = (((size_t)b * Hp + py) * Wp + px) * C + c8 * 8;
    const uint4 av = *reinterpret_cast<const uint4*>(act + (((size_t)b * aHb + py + aoff) * aWb + px + aoff) * C + c8 * 8);
    const uint2 ar = *reinterpret_cast<const uint2*>(arg + pp);
    const unsigned short* as = reinterpret_cast<const unsigned short*>(&av);
    const unsigned char* ab = reinterpret_cast<const unsigned char*>(&ar);
    for (int pl = 0; pl < planes; ++pl) {
      const uint4 g = *reinterpret_cast<const uint4*>(dpool + pl * dp_plane + pp);
      const unsigned short* gs = reinterpret_cast<const unsigned short*>(&g);
      unsigned short o[4][8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool pos = (as[i] & 0x7FFFu) != 0 && !(as[i] & 0x8000u);
#pragma unroll
        for (int w = 0; w < 4; ++w) o[w][i] = (pos && ab[i] == w) ? gs[i] : (unsigned short)0;
        if (pos && ab[i] < 4) bsum[i] += __uint_as_float((unsigned int)gs[i] << 16);
      }
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int y = 2 * py + (w >> 1) + doff, x = 2 * px + (w & 1) + doff;
        *reinterpret_cast<uint4*>(dy + pl * dy_plane + (((size_t)b * dHb + y) * dWb + x) * C + c8 * 8) = *reinterpret_cast<const uint4*>(o[w]);
      }
    }
  }
  if (colsum) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.0f;
    __syncthreads();
    const int c8 = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % cg);
#pragma unroll
    for (int i = 0; i < 8; ++i) if (bsum[i] != 0.0f) atomicAdd(&sacc[c8 * 8 + i], bsum[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x)
      if (sacc[i] != 0.0f) atomicAdd(colsum + i / group, sacc[i]);
  }
}

// ---- [R][C] bf16 -> [C][R] bf16 -----------------------------------------------------------------
__global__ void transpose_bf16_kernel(long long R, int C, const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out) {
  __shared__ unsigned short tile[64][66];
  const long long r0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const unsigned short* src = reinterpret_cast<const unsigned short*>(in);
  unsigned short* dst = reinterpret_cast<unsigned short*>(out);
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const long long r = r0 + i;
    for (int j = threadIdx.x; j < 64; j += blockDim.x) {
      const int c = c0 + j;
      tile[i][j] = (r < R && c < C) ? src[r * C + c] : (unsigned short)0;
    }
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 64; j += blockDim.y) {
    const int c = c0 + j;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
      const long long r = r0 + i;
      if (r < R && c < C) dst[(size_t)c * R + r] = tile[i][j];
    }
  }
}

// ---- adjoint of expand_reg_reg: dpsi[o,i,t,ys,xs] += dWmat[(o,r)][tap][(i,s)] -----------------------
__device__ __forceinline__ void rot_src_b(int r, int y, int x, int& ys, int& xs) {
  switch (r & 3) {
    case 0: ys = y; xs = x; break;
    case 1: ys = x; xs = 2 - y; break;
    case 2: ys = 2 - y; xs = 2 - x; break;
    default: ys = 2 - x; xs = y; break;
  }
}
__global__ void project_reg_reg_kernel(const float* __restrict__ dwmat, int Fo, int Fi, float* __restrict__ dpsi) {
  const int Cin = Fi * 4;
  const long long total = (long long)Fo * 4 * 9 * Cin;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long rr = e;
    const int ci = (int)(rr % Cin); rr /= Cin;
    const int tap = (int)(rr % 9); rr /= 9;
    const int co = (int)rr;
    const int o = co >> 2, r = co & 3, i = ci >> 2, s = ci & 3;
    const int y = tap / 3, x = tap - 3 * y;
    int ys, xs;
    rot_src_b(r, y, x, ys, xs);
    atomicAdd(dpsi + ((((size_t)o * Fi + i) * 4 + ((s - r) & 3)) * 3 + ys) * 3 + xs, dwmat[e]);
  }
}
// per-field bias gradient from an NHWC gradient buffer [Q][C]: out[c / group] += sum_q in[q][c].
// HBM-bound column reduction: a thread owns 8 adjacent channels (one 16-B load per row), the CTA's 256 threads
// are (C/8) channel lanes x row lanes, rows are strided over the whole grid; partial sums meet in shared memory
// before one atomicAdd per channel and CTA.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(long long Q, int C, const __nv_bfloat16* __restrict__ in, int group,
                                                        float* __restrict__ out) {
  __shared__ float sacc[1024];
  const int c8n = C >> 3;                              // 8-channel groups per row (<= 128)
  const int rl_n = 256 / c8n;                          // row lanes per CTA
  const int cl = threadIdx.x % c8n, rl = threadIdx.x / c8n;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < rl_n) {
    const long long step = (long long)gridDim.x * rl_n;
    long long r = (long long)blockIdx.x * rl_n + rl;
    // four independent 16-B loads in flight per thread (the accumulation order per thread stays row order)
    for (; r + 3 * step < Q; r += 4 * step) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const uint4*>(in + (r + u * step) * C + cl * 8));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned int w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(w[j] << 16);
          acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
        }
      }
    }
    for (; r < Q; r += step) {
      const uint4 v = *reinterpret_cast<const uint4*>(in + r * C + cl * 8);
      const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += __uint_as_float(w[j] << 16);
        acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
      }
    }
  }
  for (int i = threadIdx.x; i < C; i += 256) sacc[i] = 0.0f;
  __syncthreads();
  if (rl < rl_n) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sacc[cl * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256)
    if (sacc[i] != 0.0f) atomicAdd(out + i / group, sacc[i]);
}

// ---- layer-0 weight gradient fused with its un-pooling ---------------------------------------------
// dpsi0[o][ci][ys][xs] += sum over pooled pixels of g * in[ci][2py+wy+dy-1][2px+wx+dx-1] routed through the
// rotation; g = da1[b,py,px,co] where a1 > 0, (wy,wx) = arg.  Thread = output channel, block = pixel strip.
constexpr int C0W_PIX = 32;               // pooled pixels staged per barrier pair
__global__ void __launch_bounds__(256)
conv0_wgrad_kernel(const float* __restrict__ obs, const float* __restrict__ state, const __nv_bfloat16* __restrict__ da1,
                   int planes /*of da1, stacked B * 64 * 64 * 64 elements apart*/,
                   const __nv_bfloat16* __restrict__ a1 /*[B,66,66,64]*/, const unsigned char* __restrict__ arg, int B,
                   float* __restrict__ dw0 /*[64][2][9] expanded-channel gradient*/, float* __restrict__ dbias_ch /*[64]*/) {
  __shared__ __align__(16) float patch[C0W_PIX][2][4][4];
  const int co = threadIdx.x & 63, sub = threadIdx.x >> 6;
  float acc[2][9];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[c][t] = 0.0f;
  float bacc = 0.0f;
  const long long npix = (long long)B * 64 * 64;
  for (long long base = (long long)blockIdx.x * C0W_PIX; base < npix; base += (long long)gridDim.x * C0W_PIX) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < C0W_PIX * 32 / 256; ++q) {
      const int idx = q * 256 + threadIdx.x;
      const int pi = idx >> 5, e = idx & 31, ci = e >> 4, i = (e >> 2) & 3, j = e & 3;
      const long long pix = base + pi;
      float v = 0.0f;
      if (pix < npix) {
        const int px = (int)(pix & 63), py = (int)((pix >> 6) & 63), b = (int)(pix >> 12);
        const int yy = 2 * py - 1 + i, xx = 2 * px - 1 + j;
        if (yy >= 0 && yy < 128 && xx >= 0 && xx < 128) v = ci == 0 ? __ldg(obs + ((size_t)b * 128 + yy) * 128 + xx) : __ldg(state + b);
      }
      patch[pi][ci][i][j] = v;
    }
    __syncthreads();
    // this thread's channel over 8 of the staged pixels; the three per-pixel loads of all 8 are issued up front
    // raw loads first (nothing consumes them inside this loop, so all 24 are in flight together), conversions after
    unsigned short ar[C0W_PIX / 4];
    float gsum[C0W_PIX / 4];
    unsigned char wr[C0W_PIX / 4];
    const unsigned short* a1u = reinterpret_cast<const unsigned short*>(a1);
    const unsigned short* da1u = reinterpret_cast<const unsigned short*>(da1);
    const size_t gplane = (size_t)B * 64 * 64 * 64;
#pragma unroll
    for (int k = 0; k < C0W_PIX / 4; ++k) {
      long long pix = base + sub + 4 * k;
      const bool ok = pix < npix;
      if (!ok) pix = npix - 1;                         // clamp instead of branching; masked below
      const int px = (int)(pix & 63), py = (int)((pix >> 6) & 63), b = (int)(pix >> 12);
      ar[k] = __ldg(a1u + (((size_t)b * 66 + py + 1) * 66 + px + 1) * 64 + co);
      gsum[k] = 0.0f;
      for (int pl = planes - 1; pl >= 0; --pl) gsum[k] += __uint_as_float((unsigned int)__ldg(da1u + pl * gplane + (size_t)pix * 64 + co) << 16);
      wr[k] = __ldg(arg + (size_t)pix * 64 + co);
      if (!ok) ar[k] = 0;
    }
    float av[C0W_PIX / 4], gv[C0W_PIX / 4];
    int wv[C0W_PIX / 4];
#pragma unroll
    for (int k = 0; k < C0W_PIX / 4; ++k) {
      av[k] = __uint_as_float((unsigned int)ar[k] << 16);
      gv[k] = gsum[k];
      wv[k] = wr[k];
    }
#pragma unroll
    for (int k = 0; k < C0W_PIX / 4; ++k) {
      // the 4x4 patch of both channels is warp-uniform: broadcast loads into registers, then the pool window
      // (wy, wx) of THIS channel picks its 3x3 view with selects (no lane-divergent shared-memory addresses)
      const float g = av[k] > 0.0f ? gv[k] : 0.0f;
      const bool wy = (wv[k] >> 1) != 0, wx = (wv[k] & 1) != 0;
      bacc += g;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float p[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 r = *reinterpret_cast<const float4*>(&patch[sub + 4 * k][c][i][0]);
          p[i][0] = r.x; p[i][1] = r.y; p[i][2] = r.z; p[i][3] = r.w;
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          float q[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) q[j] = wy ? p[dy + 1][j] : p[dy][j];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) acc[c][dy * 3 + dx] = fmaf(g, wx ? q[dx + 1] : q[dx], acc[c][dy * 3 + dx]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(dw0 + (co * 2 + c) * 9 + t, acc[c][t]);
  atomicAdd(dbias_ch + co, bacc);
}
// dpsi0[o][ci][ys][xs] = sum_r dw0[(o,r)][ci][tap(r)] ; dbias_f[o] = sum_r dbias_ch[(o,r)]
__global__ void project_conv0_kernel(const float* __restrict__ dw0, const float* __restrict__ dbias_ch, float* __restrict__ dpsi,
                                     float* __restrict__ dbias_f) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 64 * 18) {
    const int co = e / 18, rem = e - co * 18, ci = rem / 9, tap = rem - ci * 9;
    const int o = co >> 2, r = co & 3, y = tap / 3, x = tap - 3 * y;
    int ys, xs;
    rot_src_b(r, y, x, ys, xs);
    atomicAdd(dpsi + ((o * 2 + ci) * 3 + ys) * 3 + xs, dw0[e]);
  }
  if (e < 64) atomicAdd(dbias_f + (e >> 2), dbias_ch[e]);
}

// ---- elementwise helpers ---------------------------------------------------------------------------
// out_bf16[r][c] = relu(in_f32[r][c] + bias[c])
// (every helper: `planes` output planes, stacked n elements apart)
__global__ void bias_relu_kernel(long long n, int C, const float* __restrict__ in, const float* __restrict__ bias,
                                 __nv_bfloat16* __restrict__ out, int planes) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    store_planes(out + i, (size_t)n, planes, fmaxf(in[i] + bias[(int)(i % C)], 0.0f));
}
// out_bf16 = g_f32 * (ref_bf16 > 0)
__global__ void relu_mask_kernel(long long n, const float* __restrict__ g, const __nv_bfloat16* __restrict__ ref,
                                 __nv_bfloat16* __restrict__ out, int planes) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    store_planes(out + i, (size_t)n, planes, __bfloat162float(ref[i]) > 0.0f ? g[i] : 0.0f);
}
__global__ void f32_to_bf16_kernel(long long n, const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int planes) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    store_planes(out + i, (size_t)n, planes, in[i]);
}

// ---- heads + loss: one warp per sample --------------------------------------------------------------
struct HeadLossDev {
  int B;
  const float* a_out;      // [B,16] actor head pre-bias output (cols 0..9 used): irrep(1) dx,dy | 8 trivial
  const float* a_bias;     // [10] (first two are zero: irrep fields carry no bias)
  const float* c_pre;      // [B,512] critic head-1 pre-activation (no bias)
  const float* c_bias1;    // [512] per-channel
  const float* c_w2;       // [128]
  const float* c_b2;       // [1]
  const float *action, *oldlp, *adv, *ret, *vold;   // action [B,5]
  const double* moments;   // [3] sum, sumsq, n of adv (NULL: no normalisation)
  float clip, clip_lo, clip_hi, ent_c, vf_c, inv_m;
  int clip_vloss;
  __nv_bfloat16* d_a_out;  // [B,16] gradient wrt actor head output (bf16, zero padded)
  __nv_bfloat16* d_c_h;    // [B,512] gradient wrt critic head-1 pre-activation
  int planes;              // planes of d_a_out / d_c_h (stacked B * 16 / B * 512 elements apart)
  float* d_head;           // [10 + 128 + 1 + 512]: d a_bias | d c_w2 | d c_b2 | d c_bias1(per channel)
  float* stats;            // [8] sums: policy loss, value loss (x vf_c as the reference), entropy, -logr, r-1-logr, clip
  float* value_out;        // [B]
  float* logp_out;         // [B]
};
__global__ void __launch_bounds__(256) head_loss_kernel(HeadLossDev a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float adv_mean = 0.0f, adv_den = 1.0f;
  if (a.moments) {
    const double n = a.moments[2], s = a.moments[0], ss = a.moments[1];
    const double mean = s / n;
    double var = (ss - s * mean) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    adv_mean = (float)mean; adv_den = (float)sqrt(var) + 1e-8f;
  }
  float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dbias_a[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) dbias_a[k] = 0.0f;
  float db2 = 0.0f;
  for (int b = warp; b < a.B; b += nwarps) {
    // ---- critic head: relu(pre + bias) -> max over the 4 group channels -> dot w2
    float hv[16];
    int field_arg[4];
    float pooled[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {          // lane owns fields lane, lane+32, lane+64, lane+96
      const int fld = lane + 32 * f;
      float best = 0.0f; int bi = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float h = fmaxf(a.c_pre[(size_t)b * 512 + fld * 4 + r] + a.c_bias1[fld * 4 + r], 0.0f);
        hv[f * 4 + r] = h;
        if (r == 0 || h > best) { best = h; bi = r; }
      }
      pooled[f] = best; field_arg[f] = bi;
    }
    float vpart = 0.0f;
#pragma unroll
    for (int f = 0; f < 4; ++f) vpart = fmaf(pooled[f], a.c_w2[lane + 32 * f], vpart);
    const float value = warp_sum(vpart) + a.c_b2[0];
    // ---- actor head decode (equiv.py:86-90) + Normal (robot_actor_critic.py:115-130)
    float o10[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) o10[k] = a.a_out[(size_t)b * 16 + k] + a.a_bias[k];
    // mean = [inv0, dx, dy, inv1, inv2] = out[2], out[0], out[1], out[3], out[4]; log_std = clamp(out[5:10], -20, 2)
    const int mean_src[5] = {2, 0, 1, 3, 4};
    float logp = 0.0f, ent = 0.0f, dmean[5], dls[5];
    const float LOG_SQRT_2PI = 0.91893853320467267f;
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const float mu = o10[mean_src[d]];
      const float lraw = o10[5 + d];
      const float ls = fminf(fmaxf(lraw, -20.0f), 2.0f);
      const float sd = expf(ls), var = sd * sd;
      const float diff = a.action[(size_t)b * 5 + d] - mu;
      const float lsc = logf(sd);
      logp += -(diff * diff) / (2.0f * var) - lsc - LOG_SQRT_2PI;
      ent += 0.5f + LOG_SQRT_2PI + lsc;
      dmean[d] = diff / var;                                   // d logp / d mu
      const float inr = (lraw >= -20.0f && lraw <= 2.0f) ? 1.0f : 0.0f;
      dls[d] = (diff * diff / var - 1.0f) * inr;               // d logp / d log_std (through the clamp)
      // d entropy / d log_std = inr
      o10[5 + d] = inr;
    }
    // ---- PPO loss (robot_ppo.py:345-398)
    const float logr = logp - a.oldlp[b];
    const float ratio = expf(logr);
    const float advn = a.moments ? (a.adv[b] - adv_mean) / adv_den : a.adv[b];
    const float l1 = -advn * ratio, l2 = -advn * fminf(fmaxf(ratio, a.clip_lo), a.clip_hi);
    const float w1 = l1 > l2 ? 1.0f : (l1 == l2 ? 0.5f : 0.0f);
    const float inr = (ratio >= a.clip_lo && ratio <= a.clip_hi) ? 1.0f : 0.0f;
    const float g_logp = -advn * (w1 + (1.0f - w1) * inr) * ratio * a.inv_m;
    const float g_H = -a.ent_c * a.inv_m;
    const float R = a.ret[b], vold = a.vold[b];
    float dv, vl;
    if (a.clip_vloss) {
      const float du = value - R, vu = du * du;
      const float dd = value - vold, vc = vold + fminf(fmaxf(dd, -a.clip), a.clip);
      const float dc = vc - R, lc = dc * dc;
      const float ww = vu > lc ? 1.0f : (vu == lc ? 0.5f : 0.0f);
      const float ir = (dd >= -a.clip && dd <= a.clip) ? 1.0f : 0.0f;
      dv = (ww * du + (1.0f - ww) * dc * ir) * a.vf_c * a.inv_m;
      vl = 0.5f * fmaxf(vu, lc);
    } else {
      const float du = value - R;
      dv = du * a.vf_c * a.inv_m;
      vl = 0.5f * du * du;
    }
    if (lane == 0) {
      st[0] += fmaxf(l1, l2); st[1] += vl * a.vf_c; st[2] += ent; st[3] += -logr; st[4] += (ratio - 1.0f) - logr;
      st[5] += fabsf(ratio - 1.0f) > a.clip ? 1.0f : 0.0f;
      if (a.value_out) a.value_out[b] = value;
      if (a.logp_out) a.logp_out[b] = logp;
      // gradient wrt the 10 head outputs
      float d10[10];
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        d10[mean_src[d]] = g_logp * dmean[d];
        d10[5 + d] = g_logp * dls[d] + g_H * o10[5 + d];
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) store_planes(a.d_a_out + (size_t)b * 16 + k, (size_t)a.B * 16, a.planes, k < 10 ? d10[k] : 0.0f);
#pragma unroll
      for (int k = 0; k < 10; ++k) dbias_a[k] += d10[k];
      db2 += dv;
    }
    // ---- critic head backward: dv -> pooled -> arg-max channel with relu mask
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int fld = lane + 32 * f;
      const float gp = dv * a.c_w2[fld];
      atomicAdd(a.d_head + 10 + fld, dv * pooled[f]);            // d c_w2
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float g = (r == field_arg[f] && hv[f * 4 + r] > 0.0f) ? gp : 0.0f;
        store_planes(a.d_c_h + (size_t)b * 512 + fld * 4 + r, (size_t)a.B * 512, a.planes, g);
        if (g != 0.0f) atomicAdd(a.d_head + 10 + 128 + 1 + fld * 4 + r, g);   // d c_bias1 (per channel)
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd(a.stats + k, st[k]);
#pragma unroll
    for (int k = 0; k < 10; ++k) atomicAdd(a.d_head + k, dbias_a[k]);
    atomicAdd(a.d_head + 10 + 128, db2);
  }
}

// ---- heads, inference: robot_actor_critic.evaluate / value (src/models/robot_actor_critic.py:57-60,104-131) -------------
// One warp per sample (the critic head needs 512 channels), lane 0 decodes the actor head: Normal(mean, exp(log_std)),
// action = given or mean + std * N(0,1) (Philox keyed like squash.cu), summed log-prob / entropy, decodeActions scaling
// (robot_actor_critic.py:63-82) with the reference's operation order (no FMA contraction: bit-exact with torch).
struct HeadEvalDev {
  int B;
  const float *a_out, *a_bias, *c_pre, *c_bias1, *c_w2, *c_b2, *action_in;
  unsigned long long seed, stream_id;
  float lo[5], hi[5];      // action ranges in the order p, dx, dy, dz, dtheta
  float *unscaled_out, *scaled_out, *logp_out, *ent_out, *value_out, *mean_out, *logstd_out;
  // plain CNN heads (src/nets/base_cnns.py:57-84): a_out cols 0..4 = mean_linear output, log_std = actor_logstd parameter,
  // critic = Linear(128,128)-ReLU-Linear(128,1) on c_pre [B,128] (no group pooling)
  const float* plain_logstd;   // [5] or NULL (equivariant heads)
};
__global__ void __launch_bounds__(256) head_eval_kernel(HeadEvalDev a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < a.B; b += nwarps) {
    if (a.c_pre) {
      float vpart = 0.0f;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int fld = lane + 32 * f;
        float best = 0.0f;
        if (a.plain_logstd) {
          best = fmaxf(a.c_pre[(size_t)b * 128 + fld] + a.c_bias1[fld], 0.0f);
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r) best = fmaxf(best, fmaxf(a.c_pre[(size_t)b * 512 + fld * 4 + r] + a.c_bias1[fld * 4 + r], 0.0f));
        }
        vpart = fmaf(best, a.c_w2[fld], vpart);
      }
      const float value = warp_sum(vpart) + a.c_b2[0];
      if (lane == 0) a.value_out[b] = value;
    }
    if (a.a_out && lane == 0) {
      float o10[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) o10[k] = a.a_out[(size_t)b * 16 + k] + (a.plain_logstd && k >= 5 ? 0.0f : a.a_bias[k]);
      const int mean_src_e[5] = {2, 0, 1, 3, 4}, mean_src_p[5] = {0, 1, 2, 3, 4};
      const int* mean_src = a.plain_logstd ? mean_src_p : mean_src_e;
      const float LOG_SQRT_2PI = 0.91893853320467267f;
      float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (!a.action_in) {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const Philox r = philox4x32_10((uint32_t)b, 0u, (uint32_t)a.stream_id, (uint32_t)(a.stream_id >> 32) ^ ((uint32_t)blk << 24),
                                         (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
          float zz[POL_OUT_MAX];
          normal4(r, zz);
#pragma unroll
          for (int j = 0; j < 4; ++j) z[blk * 4 + j] = zz[j];
        }
      }
      float logp = 0.0f, ent = 0.0f;
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        const float mu = o10[mean_src[d]];
        const float ls = a.plain_logstd ? a.plain_logstd[d] : fminf(fmaxf(o10[5 + d], -20.0f), 2.0f);
        const float sd = expf(ls), var = sd * sd;
        const float x = a.action_in ? a.action_in[(size_t)b * 5 + d] : __fadd_rn(mu, __fmul_rn(sd, z[d]));
        const float diff = x - mu, lsc = logf(sd);
        logp += -(diff * diff) / (2.0f * var) - lsc - LOG_SQRT_2PI;
        ent += 0.5f + LOG_SQRT_2PI + lsc;
        a.unscaled_out[(size_t)b * 5 + d] = x;
        // 0.5 * (u + 1) * (hi - lo) + lo, evaluated left to right as torch does
        const float t = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(x, 1.0f)), __fsub_rn(a.hi[d], a.lo[d]));
        a.scaled_out[(size_t)b * 5 + d] = __fadd_rn(t, a.lo[d]);
        if (a.mean_out) a.mean_out[(size_t)b * 5 + d] = mu;
        if (a.logstd_out) a.logstd_out[(size_t)b * 5 + d] = ls;
      }
      a.logp_out[b] = logp;
      a.ent_out[b] = ent;
    }
  }
}

// ---- plain CNN heads + loss (robot_actor_critic, equivariant = False): base_actor.mean_linear + actor_logstd,
// base_critic.critic = Linear(128,128)-ReLU-Linear(128,1) (src/nets/base_cnns.py:57-84, src/models/robot_actor_critic.py:44-51),
// same PPO loss seeds as head_loss_kernel (src/robot_ppo.py:345-398).  One warp per sample.
struct PlainHeadDev {
  int B;
  const float* a_out;      // [B,16] mean_linear output before bias (cols 0..4 used)
  const float* a_bias;     // [5]
  const float* logstd;     // [5] actor_logstd
  const float* c_pre;      // [B,128] critic.0 output before bias
  const float* c_bias1;    // [128]
  const float* c_w2;       // [128]
  const float* c_b2;       // [1]
  const float *action, *oldlp, *adv, *ret, *vold;
  const double* moments;
  float clip, clip_lo, clip_hi, ent_c, vf_c, inv_m;
  int clip_vloss;
  __nv_bfloat16* d_a_out;  // [B,16]
  __nv_bfloat16* d_c_h;    // [B,128]
  int planes;              // planes of d_a_out / d_c_h (stacked B * 16 / B * 128 elements apart)
  float* d_head;           // [5 + 5 + 128 + 1 + 128]: d a_bias | d logstd | d c_w2 | d c_b2 | d c_bias1
  float* stats;
  float* value_out;
  float* logp_out;
};
__global__ void __launch_bounds__(256) plain_head_loss_kernel(PlainHeadDev a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float adv_mean = 0.0f, adv_den = 1.0f;
  if (a.moments) {
    const double n = a.moments[2], s = a.moments[0], ss = a.moments[1];
    const double mean = s / n;
    double var = (ss - s * mean) / (n - 1.0);
    if (var < 0.0) var = 0.0;
    adv_mean = (float)mean; adv_den = (float)sqrt(var) + 1e-8f;
  }
  float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dmu_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, dls_acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float db2 = 0.0f, dw2[4] = {0.f, 0.f, 0.f, 0.f}, db1[4] = {0.f, 0.f, 0.f, 0.f};
  const float LOG_SQRT_2PI = 0.91893853320467267f;
  for (int b = warp; b < a.B; b += nwarps) {
    float h[4], vpart = 0.0f;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int c = lane + 32 * f;
      h[f] = fmaxf(a.c_pre[(size_t)b * 128 + c] + a.c_bias1[c], 0.0f);
      vpart = fmaf(h[f], a.c_w2[c], vpart);
    }
    const float value = warp_sum(vpart) + a.c_b2[0];
    float logp = 0.0f, ent = 0.0f, dmean[5], dls[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const float mu = a.a_out[(size_t)b * 16 + d] + a.a_bias[d];
      const float sd = expf(a.logstd[d]), var = sd * sd, lsc = logf(sd);
      const float diff = a.action[(size_t)b * 5 + d] - mu;
      logp += -(diff * diff) / (2.0f * var) - lsc - LOG_SQRT_2PI;
      ent += 0.5f + LOG_SQRT_2PI + lsc;
      dmean[d] = diff / var;
      dls[d] = diff * diff / var - 1.0f;
    }
    const float logr = logp - a.oldlp[b];
    const float ratio = expf(logr);
    const float advn = a.moments ? (a.adv[b] - adv_mean) / adv_den : a.adv[b];
    const float l1 = -advn * ratio, l2 = -advn * fminf(fmaxf(ratio, a.clip_lo), a.clip_hi);
    const float w1 = l1 > l2 ? 1.0f : (l1 == l2 ? 0.5f : 0.0f);
    const float inr = (ratio >= a.clip_lo && ratio <= a.clip_hi) ? 1.0f : 0.0f;
    const float g_logp = -advn * (w1 + (1.0f - w1) * inr) * ratio * a.inv_m;
    const float g_H = -a.ent_c * a.inv_m;
    const float R = a.ret[b], vold = a.vold[b];
    float dv, vl;
    if (a.clip_vloss) {
      const float du = value - R, vu = du * du;
      const float dd = value - vold, vc = vold + fminf(fmaxf(dd, -a.clip), a.clip);
      const float dc = vc - R, lc = dc * dc;
      const float ww = vu > lc ? 1.0f : (vu == lc ? 0.5f : 0.0f);
      const float ir = (dd >= -a.clip && dd <= a.clip) ? 1.0f : 0.0f;
      dv = (ww * du + (1.0f - ww) * dc * ir) * a.vf_c * a.inv_m;
      vl = 0.5f * fmaxf(vu, lc);
    } else {
      const float du = value - R;
      dv = du * a.vf_c * a.inv_m;
      vl = 0.5f * du * du;
    }
    if (lane == 0) {
      st[0] += fmaxf(l1, l2); st[1] += vl * a.vf_c; st[2] += ent; st[3] += -logr; st[4] += (ratio - 1.0f) - logr;
      st[5] += fabsf(ratio - 1.0f) > a.clip ? 1.0f : 0.0f;
      if (a.value_out) a.value_out[b] = value;
      if (a.logp_out) a.logp_out[b] = logp;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        store_planes(a.d_a_out + (size_t)b * 16 + k, (size_t)a.B * 16, a.planes, k < 5 ? g_logp * dmean[k < 5 ? k : 0] : 0.0f);
#pragma unroll
      for (int d = 0; d < 5; ++d) { dmu_acc[d] += g_logp * dmean[d]; dls_acc[d] += g_logp * dls[d] + g_H; }
      db2 += dv;
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int c = lane + 32 * f;
      const float g = h[f] > 0.0f ? dv * a.c_w2[c] : 0.0f;
      store_planes(a.d_c_h + (size_t)b * 128 + c, (size_t)a.B * 128, a.planes, g);
      dw2[f] += dv * h[f];
      db1[f] += g;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd(a.stats + k, st[k]);
#pragma unroll
    for (int d = 0; d < 5; ++d) { atomicAdd(a.d_head + d, dmu_acc[d]); atomicAdd(a.d_head + 5 + d, dls_acc[d]); }
    atomicAdd(a.d_head + 10 + 128, db2);
  }
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    atomicAdd(a.d_head + 10 + lane + 32 * f, dw2[f]);
    atomicAdd(a.d_head + 10 + 128 + 1 + lane + 32 * f, db1[f]);
  }
}

// ---- Adam on large flat buffers + global-norm pieces ---------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(long long n, const float* __restrict__ g, double* __restrict__ out) {
  __shared__ double sh[8];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = g[i];
    s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(out, t);
  }
}
// clip coefficient from *sumsq (NULL -> 1): min(1, max_norm / (sqrt(sumsq) + 1e-6))
__global__ void adam_flat_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m1,
                                 float* __restrict__ m2, float lr_over_bc1, float beta1, float beta2, float eps, float sqrt_bc2,
                                 const double* __restrict__ sumsq, float max_norm) {
  float coef = 1.0f;
  if (sumsq) {
    const float total = (float)sqrt(*sumsq);
    coef = fminf(max_norm / (total + 1e-6f), 1.0f);
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float m = m1[i], v = m2[i];
    m = m + (gi - m) * (1.0f - beta1);
    v = v * beta2 + (1.0f - beta2) * gi * gi;
    p[i] = p[i] - lr_over_bc1 * (m / (sqrtf(v) / sqrt_bc2 + eps));
    m1[i] = m; m2[i] = v;
  }
}

static unsigned grid_for(long long n, int block = 256, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace tc
}  // namespace aur

using namespace aur;
using namespace aur::tc;

extern "C" int aur_wgrad3x3_bf16(int32_t Cout, int32_t Cin, int64_t Q, const void* dy, const void* x, int32_t base_off,
                                 int32_t Wb, float* dwmat, int32_t split_k, void* stream) {
  if (Cout <= 0 || Cin <= 0 || Q <= 0 || !dy || !x || !dwmat || Cout % 8 != 0 || Cin % 8 != 0) {
    set_error("aur_wgrad3x3_bf16: bad arguments (channel counts must be multiples of 8)"); return AUR_ERR_ARG;
  }
  CUtensorMap tmA, tmB;
  const int P = tc_planes();
  const uint64_t dA[3] = {(uint64_t)Cout, (uint64_t)Q, (uint64_t)P}, dB[3] = {(uint64_t)Cin, (uint64_t)Q, (uint64_t)P};
  const uint64_t stA[2] = {(uint64_t)Cout * 2, (uint64_t)Cout * Q * 2}, stB[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * Q * 2};
  const uint32_t bx[3] = {64, WG_BK, 1};
  int rc;
  if ((rc = make_tensor_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dy, dA, stA, bx))) return rc;
  if ((rc = make_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dB, stB, bx))) return rc;
  static DeviceOnce attr;
  if (attr.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(wgrad3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
    attr.done();
  }
  const int mt = (Cout + WG_BM - 1) / WG_BM, nt = (Cin + WG_BN - 1) / WG_BN;
  if (split_k < 1) {
    const int tiles = mt * nt * 3;
    split_k = (2 * sm_count() + tiles - 1) / tiles;
    const long long nkb = (Q + WG_BK - 1) / WG_BK;
    if (split_k > nkb / 8) split_k = (int)(nkb / 8 > 0 ? nkb / 8 : 1);
    if (split_k < 1) split_k = 1;
  }
  if (P > 1) {
    // multi-plane precisions: bound the MMA steps that accumulate in TMEM before the fp32 (round-to-nearest) adds of the
    // epilogue - the tensor core's accumulator adds truncate, ~2e-8 per step (64 K-blocks x 3 products x 4 steps ~ 2e-5)
    const long long nkb = (Q + WG_BK - 1) / WG_BK;
    long long need = (nkb + 63) / 64;
    if (need > 65535) need = 65535;
    if (split_k < need) split_k = (int)need;
  }
  dim3 grid((unsigned)(mt * nt), 3, (unsigned)split_k);
  wgrad3x3_kernel<<<grid, 256, WG_SMEM, (cudaStream_t)stream>>>(tmA, tmB, Cout, Cin, (long long)Q, base_off, Wb, nt, dwmat, tc_terms(P));
  AUR_LAUNCH_OK("wgrad3x3_kernel");
  return 0;
}

extern "C" int aur_unpool_relu_bwd_colsum(int32_t B, int32_t Hp, int32_t Wp, int32_t C, const void* dpool, const void* act,
                                          int32_t aHb, int32_t aWb, int32_t aoff, const uint8_t* arg, void* dy, int32_t dHb,
                                          int32_t dWb, int32_t doff, int32_t group, float* colsum_out, void* stream) {
  if (B <= 0 || C % 8 != 0 || !dpool || !act || !arg || !dy) { set_error("aur_unpool_relu_bwd: bad arguments"); return AUR_ERR_ARG; }
  if (colsum_out && (group <= 0 || C > 1024 || 256 % (C / 8) != 0)) {
    set_error("aur_unpool_relu_bwd_colsum: needs C / 8 dividing 256, C <= 1024 and group >= 1"); return AUR_ERR_ARG;
  }
  const long long total = (long long)B * Hp * Wp * (C / 8);
  unpool_relu_bwd_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
      B, Hp, Wp, C, (const __nv_bfloat16*)dpool, (const __nv_bfloat16*)act, aHb, aWb, aoff, arg, (__nv_bfloat16*)dy, dHb, dWb, doff,
      group, colsum_out, tc_planes(), (size_t)B * Hp * Wp * C, (size_t)B * dHb * dWb * C);
  AUR_LAUNCH_OK("unpool_relu_bwd_kernel");
  return 0;
}
extern "C" int aur_unpool_relu_bwd(int32_t B, int32_t Hp, int32_t Wp, int32_t C, const void* dpool, const void* act,
                                   int32_t aHb, int32_t aWb, int32_t aoff, const uint8_t* arg, void* dy, int32_t dHb,
                                   int32_t dWb, int32_t doff, void* stream) {
  return aur_unpool_relu_bwd_colsum(B, Hp, Wp, C, dpool, act, aHb, aWb, aoff, arg, dy, dHb, dWb, doff, 1, nullptr, stream);
}

extern "C" int aur_transpose_bf16(int64_t R, int32_t C, const void* in, void* out, void* stream) {
  if (R <= 0 || C <= 0 || !in || !out) { set_error("aur_transpose_bf16: bad arguments"); return AUR_ERR_ARG; }
  dim3 grid((unsigned)((R + 63) / 64), (unsigned)((C + 63) / 64));
  transpose_bf16_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((long long)R, C, (const __nv_bfloat16*)in, (__nv_bfloat16*)out);
  AUR_LAUNCH_OK("transpose_bf16_kernel");
  return 0;
}

extern "C" int aur_equiv_project_regular(const float* dwmat, int32_t Fo, int32_t Fi, float* dpsi, void* stream) {
  if (!dwmat || !dpsi || Fo <= 0 || Fi <= 0) { set_error("aur_equiv_project_regular: bad arguments"); return AUR_ERR_ARG; }
  const long long total = (long long)Fo * 4 * 9 * Fi * 4;
  project_reg_reg_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(dwmat, Fo, Fi, dpsi);
  AUR_LAUNCH_OK("project_reg_reg_kernel");
  return 0;
}

extern "C" int aur_colsum_bf16(int64_t Q, int32_t C, const void* in, int32_t group, float* out, void* stream) {
  if (C <= 0 || C > 1024 || C % 8 != 0 || Q <= 0 || !in || !out || group <= 0) { set_error("aur_colsum_bf16: bad arguments (C % 8 == 0, C <= 1024)"); return AUR_ERR_ARG; }
  const int rl_n = 256 / (C / 8) > 0 ? 256 / (C / 8) : 1;
  long long grid = (Q + rl_n - 1) / rl_n;
  if (grid > 148 * 8) grid = 148 * 8;
  colsum_bf16_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((long long)Q, C, (const __nv_bfloat16*)in, group, out);
  AUR_LAUNCH_OK("colsum_bf16_kernel");
  return 0;
}

// layer-0 weight gradient of the plain CNN (channels 0..15 of the 64-channel buffers): rows 0..15 of the tensor-core result
// ARE the filter gradients [16][2][3][3], entries 0..15 of the column of ones the bias gradients
extern "C" int aur_plain_conv0_wgrad(const float* obs, const float* state, const void* da1, const void* a1, const uint8_t* arg,
                                     int32_t B, float* scratch /*[64*18 + 64], zeroed by this call*/, float* dweight, float* dbias,
                                     void* stream) {
  if (!obs || !state || !da1 || !a1 || !arg || !scratch || !dweight || !dbias || B <= 0) {
    set_error("aur_plain_conv0_wgrad: bad arguments"); return AUR_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  AUR_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float) * (64 * 18 + 64), s));
  int rc = aur::tc::launch_conv0_wgrad_tc(obs, state, da1, a1, arg, B, scratch, scratch + 64 * 18, s, true);
  if (rc) return rc;
  // further planes of the gradient stack: mid against both parts of the input views (2 planes: hi part only, mid x mid is
  // dropped there), lo against the hi part
  for (int pl = 1; pl < tc_planes(); ++pl) {
    rc = aur::tc::launch_conv0_wgrad_tc(obs, state, (const __nv_bfloat16*)da1 + (size_t)pl * B * 64 * 64 * 64, a1, arg, B, scratch,
                                         scratch + 64 * 18, s, true, (tc_planes() == 3 && pl == 1) ? 2 : 1);
    if (rc) return rc;
  }
  AUR_CUDA_OK(cudaMemcpyAsync(dweight, scratch, sizeof(float) * 16 * 18, cudaMemcpyDeviceToDevice, s));
  AUR_CUDA_OK(cudaMemcpyAsync(dbias, scratch + 64 * 18, sizeof(float) * 16, cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" int aur_equiv_conv0_wgrad(const float* obs, const float* state, const void* da1, const void* a1, const uint8_t* arg,
                                     int32_t B, float* scratch /*[64*18 + 64], zeroed by this call*/, float* dpsi, float* dbias_f,
                                     void* stream) {
  if (!obs || !state || !da1 || !a1 || !arg || !scratch || !dpsi || !dbias_f || B <= 0) {
    set_error("aur_equiv_conv0_wgrad: bad arguments"); return AUR_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  AUR_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float) * (64 * 18 + 64), s));
  static int simt = -1;                              // AUR_CONV0_WGRAD=simt keeps the fp32 SIMT kernel (cross-check)
  if (simt < 0) { const char* e = getenv("AUR_CONV0_WGRAD"); simt = (e && e[0] == 's') ? 1 : 0; }
  if (simt) {
    conv0_wgrad_kernel<<<148 * 8, 256, 0, s>>>(obs, state, (const __nv_bfloat16*)da1,
                                              tc_planes(), (const __nv_bfloat16*)a1, arg, B, scratch, scratch + 64 * 18);
    AUR_LAUNCH_OK("conv0_wgrad_kernel");
  } else {
    int rc = aur::tc::launch_conv0_wgrad_tc(obs, state, da1, a1, arg, B, scratch, scratch + 64 * 18, s);
    if (rc) return rc;
    for (int pl = 1; pl < tc_planes(); ++pl) {
      rc = aur::tc::launch_conv0_wgrad_tc(obs, state, (const __nv_bfloat16*)da1 + (size_t)pl * B * 64 * 64 * 64, a1, arg, B, scratch,
                                           scratch + 64 * 18, s, false, (tc_planes() == 3 && pl == 1) ? 2 : 1);
      if (rc) return rc;
    }
  }
  project_conv0_kernel<<<(64 * 18 + 255) / 256, 256, 0, s>>>(scratch, scratch + 64 * 18, dpsi, dbias_f);
  AUR_LAUNCH_OK("project_conv0_kernel");
  return 0;
}

extern "C" int aur_bias_relu_bf16(int64_t rows, int32_t C, const float* in, const float* bias, void* out, void* stream) {
  if (rows <= 0 || C <= 0 || !in || !bias || !out) { set_error("aur_bias_relu_bf16: bad arguments"); return AUR_ERR_ARG; }
  bias_relu_kernel<<<grid_for(rows * C), 256, 0, (cudaStream_t)stream>>>(rows * C, C, in, bias, (__nv_bfloat16*)out, tc_planes());
  AUR_LAUNCH_OK("bias_relu_kernel");
  return 0;
}
extern "C" int aur_relu_mask_bf16(int64_t n, const float* g, const void* ref, void* out, void* stream) {
  if (n <= 0 || !g || !out) { set_error("aur_relu_mask_bf16: bad arguments"); return AUR_ERR_ARG; }
  if (ref) relu_mask_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, g, (const __nv_bfloat16*)ref, (__nv_bfloat16*)out, tc_planes());
  else f32_to_bf16_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, g, (__nv_bfloat16*)out, tc_planes());
  AUR_LAUNCH_OK("relu_mask_kernel");
  return 0;
}

extern "C" int aur_equiv_head_loss(const aur_equiv_head_args* h, void* stream) {
  if (!h || h->B <= 0 || !h->a_out || !h->a_bias || !h->c_pre || !h->c_bias1 || !h->c_w2 || !h->c_b2 || !h->action || !h->oldlp ||
      !h->adv || !h->ret || !h->vold || !h->d_a_out || !h->d_c_h || !h->d_head || !h->stats || h->m_total <= 0) {
    set_error("aur_equiv_head_loss: bad arguments"); return AUR_ERR_ARG;
  }
  HeadLossDev d;
  d.B = h->B; d.a_out = h->a_out; d.a_bias = h->a_bias; d.c_pre = h->c_pre; d.c_bias1 = h->c_bias1; d.c_w2 = h->c_w2; d.c_b2 = h->c_b2;
  d.action = h->action; d.oldlp = h->oldlp; d.adv = h->adv; d.ret = h->ret; d.vold = h->vold; d.moments = h->adv_moments;
  d.clip = h->clip_coeff; d.clip_lo = (float)(1.0 - (double)h->clip_coeff); d.clip_hi = (float)(1.0 + (double)h->clip_coeff);
  d.ent_c = h->entropy_coeff; d.vf_c = h->value_coeff; d.inv_m = (float)(1.0 / (double)h->m_total); d.clip_vloss = h->clip_vloss;
  d.d_a_out = (__nv_bfloat16*)h->d_a_out; d.d_c_h = (__nv_bfloat16*)h->d_c_h; d.d_head = h->d_head; d.stats = h->stats;
  d.planes = tc_planes();
  d.value_out = h->value_out; d.logp_out = h->logp_out;
  const unsigned grid = grid_for((long long)h->B * 32, 256, 148 * 4);
  head_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d);
  AUR_LAUNCH_OK("head_loss_kernel");
  return 0;
}

static int head_eval_impl(int32_t B, const float* a_out, const float* a_bias, const float* c_pre, const float* c_bias1,
                          const float* c_w2, const float* c_b2, const float* action_in, uint64_t seed, uint64_t stream_id,
                          const float* ranges_lo_hi, float* unscaled_out, float* scaled_out, float* logp_out,
                          float* entropy_out, float* value_out, float* mean_out, float* logstd_out, const float* plain_logstd,
                          void* stream);
extern "C" int aur_equiv_head_eval(int32_t B, const float* a_out, const float* a_bias, const float* c_pre, const float* c_bias1,
                                   const float* c_w2, const float* c_b2, const float* action_in, uint64_t seed, uint64_t stream_id,
                                   const float* ranges_lo_hi, float* unscaled_out, float* scaled_out, float* logp_out,
                                   float* entropy_out, float* value_out, float* mean_out, float* logstd_out, void* stream) {
  return head_eval_impl(B, a_out, a_bias, c_pre, c_bias1, c_w2, c_b2, action_in, seed, stream_id, ranges_lo_hi, unscaled_out,
                        scaled_out, logp_out, entropy_out, value_out, mean_out, logstd_out, nullptr, stream);
}
extern "C" int aur_plain_head_eval(int32_t B, const float* a_out, const float* a_bias, const float* actor_logstd, const float* c_pre,
                                   const float* c_bias1, const float* c_w2, const float* c_b2, const float* action_in, uint64_t seed,
                                   uint64_t stream_id, const float* ranges_lo_hi, float* unscaled_out, float* scaled_out,
                                   float* logp_out, float* entropy_out, float* value_out, float* mean_out, float* logstd_out,
                                   void* stream) {
  if (!actor_logstd) { set_error("aur_plain_head_eval: actor_logstd is required"); return AUR_ERR_ARG; }
  return head_eval_impl(B, a_out, a_bias, c_pre, c_bias1, c_w2, c_b2, action_in, seed, stream_id, ranges_lo_hi, unscaled_out,
                        scaled_out, logp_out, entropy_out, value_out, mean_out, logstd_out, actor_logstd, stream);
}
static int head_eval_impl(int32_t B, const float* a_out, const float* a_bias, const float* c_pre, const float* c_bias1,
                          const float* c_w2, const float* c_b2, const float* action_in, uint64_t seed, uint64_t stream_id,
                          const float* ranges_lo_hi, float* unscaled_out, float* scaled_out, float* logp_out,
                          float* entropy_out, float* value_out, float* mean_out, float* logstd_out, const float* plain_logstd,
                          void* stream) {
  if (B <= 0 || (!a_out && !c_pre)) { set_error("aur_equiv_head_eval: bad arguments"); return AUR_ERR_ARG; }
  if (a_out && (!a_bias || !ranges_lo_hi || !unscaled_out || !scaled_out || !logp_out || !entropy_out)) {
    set_error("aur_equiv_head_eval: actor head needs a_bias, ranges and the four outputs"); return AUR_ERR_ARG;
  }
  if (c_pre && (!c_bias1 || !c_w2 || !c_b2 || !value_out)) {
    set_error("aur_equiv_head_eval: critic head needs c_bias1, c_w2, c_b2, value_out"); return AUR_ERR_ARG;
  }
  HeadEvalDev d;
  d.B = B; d.a_out = a_out; d.a_bias = a_bias; d.c_pre = c_pre; d.c_bias1 = c_bias1; d.c_w2 = c_w2; d.c_b2 = c_b2;
  d.action_in = action_in; d.seed = seed; d.stream_id = stream_id;
  for (int k = 0; k < 5; ++k) { d.lo[k] = a_out ? ranges_lo_hi[2 * k] : 0.f; d.hi[k] = a_out ? ranges_lo_hi[2 * k + 1] : 0.f; }
  d.unscaled_out = unscaled_out; d.scaled_out = scaled_out; d.logp_out = logp_out; d.ent_out = entropy_out; d.value_out = value_out;
  d.mean_out = mean_out; d.logstd_out = logstd_out; d.plain_logstd = plain_logstd;
  head_eval_kernel<<<grid_for((long long)B * 32, 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(d);
  AUR_LAUNCH_OK("head_eval_kernel");
  return 0;
}

extern "C" int aur_plain_head_loss(const aur_plain_head_args* h, void* stream) {
  if (!h || h->B <= 0 || !h->a_out || !h->a_bias || !h->actor_logstd || !h->c_pre || !h->c_bias1 || !h->c_w2 || !h->c_b2 ||
      !h->action || !h->oldlp || !h->adv || !h->ret || !h->vold || !h->d_a_out || !h->d_c_h || !h->d_head || !h->stats ||
      h->m_total <= 0) {
    set_error("aur_plain_head_loss: bad arguments"); return AUR_ERR_ARG;
  }
  PlainHeadDev d;
  d.B = h->B; d.a_out = h->a_out; d.a_bias = h->a_bias; d.logstd = h->actor_logstd; d.c_pre = h->c_pre; d.c_bias1 = h->c_bias1;
  d.c_w2 = h->c_w2; d.c_b2 = h->c_b2;
  d.action = h->action; d.oldlp = h->oldlp; d.adv = h->adv; d.ret = h->ret; d.vold = h->vold; d.moments = h->adv_moments;
  d.clip = h->clip_coeff; d.clip_lo = (float)(1.0 - (double)h->clip_coeff); d.clip_hi = (float)(1.0 + (double)h->clip_coeff);
  d.ent_c = h->entropy_coeff; d.vf_c = h->value_coeff; d.inv_m = (float)(1.0 / (double)h->m_total); d.clip_vloss = h->clip_vloss;
  d.d_a_out = (__nv_bfloat16*)h->d_a_out; d.d_c_h = (__nv_bfloat16*)h->d_c_h; d.d_head = h->d_head; d.stats = h->stats;
  d.planes = tc_planes();
  d.value_out = h->value_out; d.logp_out = h->logp_out;
  plain_head_loss_kernel<<<grid_for((long long)h->B * 32, 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(d);
  AUR_LAUNCH_OK("plain_head_loss_kernel");
  return 0;
}

extern "C" int aur_sumsq_f32(int64_t n, const float* g, double* out_accum, void* stream) {
  if (n <= 0 || !g || !out_accum) { set_error("aur_sumsq_f32: bad arguments"); return AUR_ERR_ARG; }
  sumsq_kernel<<<grid_for(n, 256, 148 * 2), 256, 0, (cudaStream_t)stream>>>(n, g, out_accum);
  AUR_LAUNCH_OK("sumsq_kernel");
  return 0;
}

extern "C" int aur_adam_flat(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, double lr,
                             double beta1, double beta2, double eps, int64_t step, const double* clip_sumsq,
                             double max_grad_norm, void* stream) {
  if (n <= 0 || !params || !grads || !exp_avg || !exp_avg_sq || step < 1) { set_error("aur_adam_flat: bad arguments"); return AUR_ERR_ARG; }
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_flat_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, params, grads, exp_avg, exp_avg_sq, (float)(lr / bc1),
                                                                 (float)beta1, (float)beta2, (float)eps, (float)sqrt(bc2),
                                                                 clip_sumsq, (float)max_grad_norm);
  AUR_LAUNCH_OK("adam_flat_kernel");
  return 0;
}
