// Layer-0 weight gradient of the equivariant encoder on tensor cores (row X): replaces conv0_wgrad_kernel's ~80 SIMT
// operations per (pooled pixel, channel) with a contraction over pixels.
//   dW0[co][ci][tap] = sum_pix g[pix][co] * in[ci][2py + wy + dy - 1][2px + wx + dx - 1],  (wy, wx) = pool arg-max of (pix, co)
// The pool window differs per (pixel, channel), so the sum is split by window w: G_w[pix][co] = g * 1[a1 > 0] * 1[arg = w]
// (exact in bf16: g is bf16) against P_w[pix][n] = the 3x3 view of the pixel's 4x4 input patch for window w (n = ci*9 + tap,
// n = 18 is a column of ones that yields the bias gradient), written as a bf16 hi/mid split (obs are fp32).
//   A_w: [64 pixels][64 co] MN-major, 128-B rows, 128-B swizzle (as the dz tiles of update_tc.cu; M = 128 reads a second,
//        ignored atom)      B_w: [32 n][64 pixels] K-major hi and mid (as update_tc.cu's aux tile)
//   D  : [128 lanes (co, rows 64.. ignored)][32 n] fp32 in TMEM, accumulated over every block of the CTA and both B parts.
// A CTA of 256 threads prepares one 64-pixel block (64 KB of operands), issues 32 MMAs, and moves on when they retire;
// three CTAs per SM overlap each other.  DRAM: 5 B per (pixel, channel) + the input patches.
#include "tc.cuh"

namespace aur {
namespace tc {

constexpr int C0_PIX = 64;                              // pooled pixels per block = K of one MMA batch
constexpr int C0_A_BYTES = C0_PIX * 128;                // one window's A tile
constexpr int C0_B_BYTES = 32 * 128;                    // one window's B tile (hi or mid)
constexpr int C0O_A = 0;                                // [w 4]
constexpr int C0O_B = C0O_A + 4 * C0_A_BYTES;           // [w 4][hi, mid]
constexpr int C0O_BAR = C0O_B + 8 * C0_B_BYTES;
constexpr size_t C0_SMEM = C0O_BAR + 64 + 1024;

// PLAIN: only channels 0..15 of the 64-channel layer-0 buffers carry data (aur_plain_conv0), so only the first two channel
// octets are read; rows 0..15 of the result are the plain filter gradients.
template <bool PLAIN>
__global__ void __launch_bounds__(256, 3)
conv0_wgrad_tc_kernel(const float* __restrict__ obs, const float* __restrict__ state, const __nv_bfloat16* __restrict__ da1,
                      const __nv_bfloat16* __restrict__ a1 /*[B,66,66,64]*/, const unsigned char* __restrict__ arg, int B,
                      float* __restrict__ dw0 /*[64][2][9]*/, float* __restrict__ dbias_ch /*[64]*/, int parts, int flush) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = base + C0O_A;
  unsigned char* sB = base + C0O_B;
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + C0O_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(base + C0O_BAR + 32);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(tslot, 32);
  // B tiles: rows 19..31 stay zero for the whole kernel, row 18 of the hi part is the ones column
  for (int e = tid; e < 8 * C0_B_BYTES / 16; e += 256) reinterpret_cast<uint4*>(sB)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  {
    const int w = tid >> 6, p = tid & 63;
    *reinterpret_cast<unsigned short*>(sB + (2 * w) * C0_B_BYTES + 18 * 128 + ((((p >> 3) ^ (18 & 7))) << 4) + (p & 7) * 2) = 0x3F80;
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm_d = *tslot;
  constexpr uint32_t IDESC = instr_desc(FMT_BF16, 128, 32, 1, 0);

  const long long npix = (long long)B * 64 * 64;
  const long long nblk = (npix + C0_PIX - 1) / C0_PIX;
  // flush > 0 (multi-plane precisions): the TMEM accumulator is drained into the global fp32 sums every `flush` blocks - the
  // tensor core's accumulator adds truncate (~2e-8 per MMA step), and a CTA walks hundreds of blocks at minibatch 4096
  auto drain = [&]() {
    fence_after_sync();
    if (tid < 64) {                                     // accumulator row = output channel: warps 0, 1 hold rows 0..63
      float v[32];
      tmem_ld32(tm_d + ((uint32_t)(32 * warp) << 16), v);
#pragma unroll
      for (int n = 0; n < 18; ++n) atomicAdd(dw0 + tid * 18 + n, v[n]);
      atomicAdd(dbias_ch + tid, v[18]);
    }
  };
  uint32_t it = 0, since = 0;                           // since: blocks accumulated since the last drain
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x, ++it, ++since) {
    if (it > 0) mbar_wait(bar, (it - 1u) & 1u);          // the previous block's MMAs have read the tiles
    if (flush > 0 && since == (uint32_t)flush) { drain(); since = 0; }
    const long long pix0 = blk * C0_PIX;
    // ---- B: thread (window w, pixel p) writes the 18 taps of its window's view as hi / mid
    {
      const int w = tid >> 6, p = tid & 63, wy = w >> 1, wx = w & 1;
      const long long pix = pix0 + p;
      float v[2][3][3];
#pragma unroll
      for (int i = 0; i < 18; ++i) (&v[0][0][0])[i] = 0.0f;
      if (pix < npix) {
        const int px = (int)(pix & 63), py = (int)((pix >> 6) & 63), b = (int)(pix >> 12);
        const float st = __ldg(state + b);
        const float* img = obs + (size_t)b * 128 * 128;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int yy = 2 * py - 1 + wy + dy, xx = 2 * px - 1 + wx + dx;
            const bool in = yy >= 0 && yy < 128 && xx >= 0 && xx < 128;
            v[0][dy][dx] = in ? __ldg(img + yy * 128 + xx) : 0.0f;
            v[1][dy][dx] = in ? st : 0.0f;
          }
      }
      unsigned char* bh = sB + (2 * w) * C0_B_BYTES;
      unsigned char* bm = bh + C0_B_BYTES;
#pragma unroll
      for (int n = 0; n < 18; ++n) {
        const float x = (&v[0][0][0])[n];
        const __nv_bfloat16 hb = __float2bfloat16_rn(x);
        const __nv_bfloat16 mb = __float2bfloat16_rn(x - __bfloat162float(hb));
        const int off = n * 128 + ((((p >> 3) ^ (n & 7))) << 4) + (p & 7) * 2;
        *reinterpret_cast<unsigned short*>(bh + off) = __bfloat16_as_ushort(hb);
        *reinterpret_cast<unsigned short*>(bm + off) = __bfloat16_as_ushort(mb);
      }
    }
    // ---- A: task = (pixel, channel octet): g masked by ReLU and split over the four pool windows
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int task = rep * 256 + tid, p = task >> 3, c = task & 7;
      const long long pix = pix0 + p;
      uint4 gq = make_uint4(0u, 0u, 0u, 0u), aq = gq;
      uint2 wq = make_uint2(0u, 0u);
      if (pix < npix && (!PLAIN || c < 2)) {
        const int px = (int)(pix & 63), py = (int)((pix >> 6) & 63), b = (int)(pix >> 12);
        gq = __ldcs(reinterpret_cast<const uint4*>(da1 + (size_t)pix * 64 + c * 8));
        aq = __ldcs(reinterpret_cast<const uint4*>(a1 + (((size_t)b * 66 + py + 1) * 66 + px + 1) * 64 + c * 8));
        wq = __ldcs(reinterpret_cast<const uint2*>(arg + (size_t)pix * 64 + c * 8));
      }
      const unsigned int gw[4] = {gq.x, gq.y, gq.z, gq.w}, aw[4] = {aq.x, aq.y, aq.z, aq.w};
      unsigned int outw[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // two channels per 32-bit word; positive bf16 <=> sign clear and magnitude non-zero
        const unsigned int al = aw[j] & 0xFFFFu, ah = aw[j] >> 16;
        const bool pl = (al & 0x7FFFu) != 0 && !(al & 0x8000u), phh = (ah & 0x7FFFu) != 0 && !(ah & 0x8000u);
        const unsigned int wl = ((j < 2 ? wq.x : wq.y) >> (16 * (j & 1))) & 0xFFu;
        const unsigned int wh = ((j < 2 ? wq.x : wq.y) >> (16 * (j & 1) + 8)) & 0xFFu;
        const unsigned int gl = pl ? (gw[j] & 0xFFFFu) : 0u, gh = phh ? (gw[j] & 0xFFFF0000u) : 0u;
#pragma unroll
        for (int w = 0; w < 4; ++w) outw[w][j] = (wl == (unsigned)w ? gl : 0u) | (wh == (unsigned)w ? gh : 0u);
      }
      const int off = p * 128 + ((c ^ (p & 7)) << 4);
#pragma unroll
      for (int w = 0; w < 4; ++w)
        *reinterpret_cast<uint4*>(sA + w * C0_A_BYTES + off) = make_uint4(outw[w][0], outw[w][1], outw[w][2], outw[w][3]);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      for (int w = 0; w < 4; ++w) {
        const uint64_t ad = smem_desc_mn_sw128(sA + w * C0_A_BYTES, C0_A_BYTES, 1024);
        for (int part = 0; part < parts; ++part) {         // parts = 1: the hi views only (second pass over the gradient's mid plane)
          const uint64_t bd = smem_desc_k_sw128(sB + (2 * w + part) * C0_B_BYTES);
          for (int k = 0; k < C0_PIX / 16; ++k)
            mma_f16(tm_d, ad + (uint64_t)(128 * k), bd + (uint64_t)(2 * k), IDESC, (since | (uint32_t)(w | part | k)) != 0u);
        }
      }
      mma_commit(bar);
    }
  }
  if (it > 0) {
    mbar_wait(bar, (it - 1u) & 1u);
    drain();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm_d, 32);
}

int launch_conv0_wgrad_tc(const float* obs, const float* state, const void* da1, const void* a1, const uint8_t* arg, int B,
                          float* dw0, float* dbias_ch, cudaStream_t s, bool plain, int parts) {
  const int flush = tc_planes() > 1 ? 8 : 0;
  static DeviceOnce attr;
  if (attr.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(conv0_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C0_SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv0_wgrad_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv0_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C0_SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv0_wgrad_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr.done();
  }
  const long long nblk = ((long long)B * 4096 + C0_PIX - 1) / C0_PIX;
  long long grid = 3LL * sm_count();
  if (grid > nblk) grid = nblk;
  if (plain)
    conv0_wgrad_tc_kernel<true><<<(unsigned)grid, 256, C0_SMEM, s>>>(obs, state, (const __nv_bfloat16*)da1, (const __nv_bfloat16*)a1, arg,
                                                                     B, dw0, dbias_ch, parts, flush);
  else
    conv0_wgrad_tc_kernel<false><<<(unsigned)grid, 256, C0_SMEM, s>>>(obs, state, (const __nv_bfloat16*)da1, (const __nv_bfloat16*)a1, arg,
                                                                      B, dw0, dbias_ch, parts, flush);
  AUR_LAUNCH_OK("conv0_wgrad_tc_kernel");
  return 0;
}

}  // namespace tc
}  // namespace aur
