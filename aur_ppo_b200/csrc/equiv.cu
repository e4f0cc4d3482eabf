// Equivariant (C4) encoder convolutions on tensor cores (row X of the scope table).
// Replaces the e2cnn R2Conv -> cuDNN conv2d calls of src/nets/equiv.py:17-59 for the
// close_loop_block_picking-shaped observations ([B,2,128,128]).
//
// A C4-steerable 3x3 filter bank is a dense filter W = expand(psi) (p4 group convolution: rotate
// the 3x3 taps by r and cyclically shift the input rotation index), so every layer is a dense
// convolution = real contraction over (tap, Cin): implicit GEMM on tcgen05.
//   A  = activations, NHWC bf16 with an explicit zero halo, loaded by 4-D TMA boxes
//        {64 channels, TW, TH, images} -> 128 pixel rows x 64 channels, K-major, 128-B swizzle;
//        the 9 taps are 9 shifted boxes of the SAME tensor map (no im2col buffer in HBM)
//   B  = expanded weights [Cout][tap][Cin] bf16, 2-D TMA
//   D  = fp32 accumulator in TMEM (128 lanes = pixels x 128 columns = output channels)
//   epilogue (tcgen05.ld): + bias, ReLU, optional 2x2 max-pool through warp shuffles (a 2x2 window
//        lives inside one warp by construction of the pixel tile), bf16 store into the NEXT
//        layer's haloed buffer, 2-bit arg-max per pooled element for the backward pass.
// The same kernel computes backward-data (input = haloed output gradient, weights = transposed,
// tap-flipped filter, linear epilogue).
#include <stdlib.h>

#include "tc.cuh"

namespace aur {
namespace tc {

// Output-channel tile BN = 128 (6 stages) or 256 (4 stages; one pixel tile then feeds twice the MMA work: 94 B of
// operands per MMA cycle instead of 128).  Both use 192 KB of operand stages and 2 x BN TMEM columns.
constexpr int CV_BM = 128, CV_BK = 64;
constexpr int CV_A_BYTES = CV_BM * CV_BK * 2;
template <int BN> struct CvCfg {
  static constexpr int STAGES = BN == 64 ? 8 : (BN == 128 ? 6 : 4);      // 192 KB of operand stages in every shape
  static constexpr int B_BYTES = BN * CV_BK * 2;
  static constexpr size_t SMEM = (size_t)STAGES * (CV_A_BYTES + B_BYTES) + 1024 + 256;
};

struct ConvDev {
  int B, Hb, Wb, Ho, Wo, Cin, Cout;
  int TH, TW, NIMG, tiles_x, tiles_y;      // pixel tile = NIMG images x TH x TW = 128
  int pix_tiles, n_tiles;                  // work items: pix_tiles x n_tiles (output-channel blocks), walked by a persistent grid
  int epi;                                 // 0 linear, 1 bias+ReLU, 2 bias+ReLU+maxpool2
  int oHb, oWb, ooff;                      // output buffer geometry
  int nterm;                               // K-step terms per (tap, channel chunk): 1 / 3 / 6 for 1 / 2 / 3 operand planes (tc.cuh)
  int planes;                              // planes of the output stack
  int chunk;                               // CHUNKED kernels: MMA groups (4 steps each) per promoted accumulator chunk
  size_t out_plane;                        // elements between the output planes
  const float* bias;
  __nv_bfloat16* out;
  unsigned char* pool_arg;
  const __nv_bfloat16* relu_ref;           // epi 3: out = acc * (relu_ref > 0), relu_ref buffer [B,rHb,rWb,Cout] at +roff
  int rHb, rWb, roff;
};

// Persistent: one CTA per SM walks the (pixel tile, channel block) items with a grid stride; the smem ring runs
// across items and the accumulator is double-buffered in TMEM (2 x 128 columns), so the epilogue of item i (TMEM ->
// bias / ReLU / pool -> bf16 stores) overlaps the TMA + MMA main loop of item i+1.
constexpr int CV_THREADS = 384;   // warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..11 epilogue (two per TMEM lane quarter)
// SWAP (pooled layers with shallow K, BN = 128): the operand roles are exchanged - D[m = output channel][n = pixel] - so a
// TMEM lane is a channel and a thread's 32 accumulator columns are 32 pixels = eight complete 2x2 pool windows.  Bias is
// one register, ReLU + max-pool + arg-max are register compares (no shuffles, no per-element bias loads): ~3 instructions
// per element instead of ~12, which is what bounded the pooled layers 1-2 (20 / 39 % tensor-pipe activity).
// RESB (Cin = 64, Cout <= 128: the pooled layer 1): the whole weight matrix (9 taps x 128 x 64 bf16 = 144 KB) is loaded
// into shared memory once per CTA instead of once per pixel tile - at 9 K-steps per tile half of the L2 -> SM traffic
// was the same weights over and over (9.2 GB of L2 reads per launch, 11.6 TB/s: the layer was L2-bandwidth bound).
constexpr int CV_RESB_STAGES = 4;
constexpr size_t CV_RESB_SMEM = (size_t)CV_RESB_STAGES * CV_A_BYTES + 9 * (128 * CV_BK * 2) + 1024 + 256;
// CHUNKED (the multi-plane precisions): the tensor core adds every K = 16 step into its fp32 accumulator with TRUNCATION, so a
// contraction over thousands of steps drifts by ~3e-8 per step (measured: 1.5e-5 at K = 4608 x 3 products, 5e-5 for layer 5,
// more than all the operand rounding of the two-plane split).  The MMA warp therefore switches between the two TMEM
// accumulators every CV_CHUNK stages instead of every work item, and the epilogue warps PROMOTE each finished chunk into fp32
// registers (round-to-nearest adds, 64 registers per thread at BN = 128); the epilogue proper then runs from the registers.
// A chunk is 32 MMA steps (~1e-6); draining it (two tcgen05.ld + 64 FADD per thread) hides behind the next chunk's MMAs.
constexpr int CV_CHUNK = 8;       // stages per promoted chunk
template <int CV_BN, bool SWAP = false, bool RESB = false, bool CHUNKED = false>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvDev a) {
  static_assert(!SWAP || CV_BN == 128, "the swapped epilogue is built for 128 x 128 accumulators");
  static_assert(!RESB || SWAP, "resident weights come with the channel-major kernel");
  static_assert(!CHUNKED || (!SWAP && CV_BN <= 128), "chunk promotion keeps BN / 2 accumulator registers per epilogue thread");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
  constexpr int CV_STAGES = RESB ? CV_RESB_STAGES : CvCfg<CV_BN>::STAGES, CV_B_BYTES = CvCfg<CV_BN>::B_BYTES;
  unsigned char* sA = smem;
  unsigned char* sB = smem + CV_STAGES * CV_A_BYTES;       // RESB: nine resident tap blocks instead of a ring
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + (RESB ? 9 : CV_STAGES) * CV_B_BYTES);
  uint64_t* empty = full + CV_STAGES;
  uint64_t* tmem_full = empty + CV_STAGES;       // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* bres = tmem_empty + 2;               // RESB: the resident weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cchunks = a.Cin / CV_BK;
  const int nkb = 9 * cchunks * a.nterm;
  const long long items = (long long)a.pix_tiles * a.n_tiles;
  // item -> (pixel tile, channel block): channel blocks of one pixel tile are adjacent, so the CTAs of a wave share A
  auto decode = [&](long long item, int& b0, int& y0, int& x0, int& n0) {
    n0 = (int)(item % a.n_tiles) * CV_BN;
    int t = (int)(item / a.n_tiles);
    const int tx = t % a.tiles_x; t /= a.tiles_x;
    const int ty = t % a.tiles_y; t /= a.tiles_y;
    b0 = t * a.NIMG; y0 = ty * a.TH; x0 = tx * a.TW;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < CV_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    mbar_init(bres, 1);
    mbar_fence_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * CV_BN);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: per item, K loop over (tap, channel chunk); the ring index runs across items =====
    uint32_t kg = 0;
    if constexpr (RESB) {
      if (blockIdx.x < items) {
        mbar_arrive_expect_tx(bres, 9 * CV_B_BYTES);
        for (int tap = 0; tap < 9; ++tap) tma_load_3d(sB + tap * CV_B_BYTES, &tmB, tap * a.Cin, 0, 0, bres);
      }
    }
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      int b0, y0, x0, n0;
      decode(item, b0, y0, x0, n0);
      if constexpr (CHUNKED) {
        // multi-plane: ONE super-stage per (tap, channel chunk) holds every plane of both operands (P x (A + B) tiles in P
        // consecutive ring slots), and all the plane products are issued from it - each tile is fetched once instead of once
        // per product (12 -> 6 tile loads per six-product step)
        const int P = a.planes, SUPER = CV_STAGES / P;
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - 3 * dy;
          for (int cc = 0; cc < cchunks; ++cc, ++kg) {
            const int u = kg % SUPER;
            const uint32_t ph = (kg / SUPER) & 1u;
            mbar_wait(&empty[u], ph ^ 1u);
            mbar_arrive_expect_tx(&full[u], P * (CV_A_BYTES + CV_B_BYTES));
            for (int pl = 0; pl < P; ++pl) {
              tma_load_5d(sA + (u * P + pl) * CV_A_BYTES, &tmA, cc * CV_BK, x0 + dx, y0 + dy, b0, pl, &full[u]);
              tma_load_3d(sB + (u * P + pl) * CV_B_BYTES, &tmB, tap * a.Cin + cc * CV_BK, n0, pl, &full[u]);
            }
          }
        }
        continue;
      }
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap - 3 * dy;
        for (int cc = 0; cc < cchunks; ++cc) {
          for (int term = 0; term < a.nterm; ++term, ++kg) {
            const int s = kg % CV_STAGES;
            const uint32_t ph = (kg / CV_STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full[s], RESB ? CV_A_BYTES : CV_A_BYTES + CV_B_BYTES);
            tma_load_5d(sA + s * CV_A_BYTES, &tmA, cc * CV_BK, x0 + dx, y0 + dy, b0, term_plane_a(term), &full[s]);
            if constexpr (!RESB) tma_load_3d(sB + s * CV_B_BYTES, &tmB, tap * a.Cin + cc * CV_BK, n0, term_plane_b(term), &full[s]);
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer: accumulator (it & 1) once the epilogue has drained its previous use =====
    constexpr uint32_t idesc = instr_desc(FMT_BF16, CV_BM, CV_BN, 0, 0);
    uint32_t kg = 0, it = 0;
    if constexpr (RESB) {
      if (blockIdx.x < items) { mbar_wait(bres, 0u); fence_after_sync(); }
    }
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      if constexpr (CHUNKED) {
        // super-stages (see the producer); an accumulator hand-over every `span` of them (~32-48 MMA steps)
        const int P = a.planes, SUPER = CV_STAGES / P, nks = 9 * cchunks, span = (a.chunk + a.nterm - 1) / a.nterm;
        for (int ks0 = 0; ks0 < nks; ks0 += span, ++it) {
          const uint32_t acc = it & 1u;
          mbar_wait(&tmem_empty[acc], ((it >> 1) & 1u) ^ 1u);
          fence_after_sync();
          const int ks1 = ks0 + span < nks ? ks0 + span : nks;
          for (int ks = ks0; ks < ks1; ++ks, ++kg) {
            const int u = kg % SUPER;
            const uint32_t ph = (kg / SUPER) & 1u;
            mbar_wait(&full[u], ph);
            fence_after_sync();
            for (int term = 0; term < a.nterm; ++term) {
              const uint64_t ad = smem_desc_k_sw128(sA + (u * P + term_plane_a(term)) * CV_A_BYTES);
              const uint64_t bd = smem_desc_k_sw128(sB + (u * P + term_plane_b(term)) * CV_B_BYTES);
#pragma unroll
              for (int k = 0; k < CV_BK / 16; ++k)
                mma_f16(tmem_d + acc * CV_BN, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, ((ks - ks0) | term | k) != 0);
            }
            mma_commit(&empty[u]);
          }
          mma_commit(&tmem_full[acc]);
        }
        continue;
      }
      // one accumulator hand-over per work item; `it` counts hand-overs across items
      const int span = nkb;
      for (int kb0 = 0; kb0 < nkb; kb0 += span, ++it) {
        const uint32_t acc = it & 1u;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1u) ^ 1u);
        fence_after_sync();
        const int kb1 = kb0 + span < nkb ? kb0 + span : nkb;
        for (int kb = kb0; kb < kb1; ++kb, ++kg) {
          const int s = kg % CV_STAGES;
          const uint32_t ph = (kg / CV_STAGES) & 1u;
          mbar_wait(&full[s], ph);
          fence_after_sync();
          const uint64_t ad = smem_desc_k_sw128(sA + s * CV_A_BYTES), bd = smem_desc_k_sw128(sB + (RESB ? kb : s) * CV_B_BYTES);
#pragma unroll
          for (int k = 0; k < CV_BK / 16; ++k) {
            if constexpr (SWAP) mma_f16(tmem_d + acc * CV_BN, bd + (uint64_t)(2 * k), ad + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0);
            else mma_f16(tmem_d + acc * CV_BN, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0);
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&tmem_full[acc]);
      }
    }
  } else if (SWAP && warp >= 4) {
    // ===== swapped epilogue: lane = output channel, columns = the tile's 128 pixels; bias + ReLU + 2x2 max-pool in registers =====
    const int q = (warp - 4) & 3, phalf = (warp - 4) >> 2;     // TMEM lane quarter (32 channels), half of the pixel columns
    uint32_t it = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      int b0, y0, x0, n0;
      decode(item, b0, y0, x0, n0);
      const uint32_t acc = it & 1u;
      const int ch = n0 + 32 * q + lane;
      const bool ch_ok = ch < a.Cout;
      const float bias = (a.bias && ch_ok) ? __ldg(a.bias + ch) : 0.0f;
      const int Hp = a.Ho >> 1, Wp = a.Wo >> 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1u);
      fence_after_sync();
#pragma unroll 1
      for (int c = 64 * phalf; c < 64 * (phalf + 1); c += 32) {
        float v[32];
        tmem_ld32(tmem_d + acc * CV_BN + ((uint32_t)(32 * q) << 16) + (uint32_t)c, v);
        if (c + 32 == 64 * (phalf + 1)) {                 // last read of this accumulator: hand it back to the MMA warp
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (!ch_ok) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bias, 0.0f);
        // pixel column p = (img * TH + yy) * TW + xx; 32 columns are 2 rows x 16 (TW = 16) or 4 rows x 8 (TW = 8)
        auto emit = [&](int img, int yy, int xx, float v0, float v1, float v2, float v3) {
          const int b = b0 + img, y = y0 + yy, x = x0 + xx;
          if (b >= a.B || y >= a.Ho || x >= a.Wo) return;
          float m = v0; unsigned int w = 0u;               // first maximum in (y, x) scan order wins ties (torch)
          if (v1 > m) { m = v1; w = 1u; }
          if (v2 > m) { m = v2; w = 2u; }
          if (v3 > m) { m = v3; w = 3u; }
          const size_t pix = ((size_t)b * a.oHb + (y >> 1) + a.ooff) * a.oWb + (x >> 1) + a.ooff;
          a.out[pix * a.Cout + ch] = __float2bfloat16_rn(m);
          if (a.pool_arg) a.pool_arg[(((size_t)b * Hp + (y >> 1)) * Wp + (x >> 1)) * a.Cout + ch] = (unsigned char)w;
        };
        if (a.TW == 16) {
          const int yy = c >> 4;                          // rows yy, yy + 1 of image 0
#pragma unroll
          for (int k = 0; k < 8; ++k) emit(0, yy, 2 * k, v[2 * k], v[2 * k + 1], v[16 + 2 * k], v[16 + 2 * k + 1]);
        } else {
          const int img = c >> 6, yy = (c & 63) >> 3;     // rows yy .. yy + 3 of image img
#pragma unroll
          for (int rp = 0; rp < 2; ++rp)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              emit(img, yy + 2 * rp, 2 * k, v[16 * rp + 2 * k], v[16 * rp + 2 * k + 1], v[16 * rp + 8 + 2 * k], v[16 * rp + 8 + 2 * k + 1]);
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> (bias, ReLU, pool) -> bf16 NHWC =====
    const int q = (warp - 4) & 3, chalf = (warp - 4) >> 2;     // TMEM lane quarter, half of the BN accumulator columns
    const int m = 32 * q + lane;
    int r = m;
    const int xx = r % a.TW; r /= a.TW;
    const int yy = r % a.TH; r /= a.TH;
    uint32_t it = 0;
    constexpr int NG = CHUNKED ? (CV_BN / 2 + 31) / 32 : 1;      // 32-column groups this warp owns (CHUNKED: kept in registers)
    float racc[NG][32];
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    int b0, y0, x0, n0;
    decode(item, b0, y0, x0, n0);
    const int b = b0 + r, y = y0 + yy, x = x0 + xx;
    const bool valid = b < a.B && y < a.Ho && x < a.Wo;
    uint32_t acc = it & 1u;
    if constexpr (CHUNKED) {
      // promote every finished chunk of this item into the register accumulators (fp32 round-to-nearest adds)
      const int nks = 9 * cchunks, span = (a.chunk + a.nterm - 1) / a.nterm;     // as the MMA warp counts them
      for (int kb0 = 0; kb0 < nks; kb0 += span, ++it) {
        acc = it & 1u;
        mbar_wait(&tmem_full[acc], (it >> 1) & 1u);
        fence_after_sync();
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int c = (CV_BN / 2) * chalf + 32 * g;
          if (c < (CV_BN / 2) * (chalf + 1)) {
            float v[32];
            tmem_ld32(tmem_d + acc * CV_BN + ((uint32_t)(32 * q) << 16) + (uint32_t)c, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) racc[g][i] = kb0 == 0 ? v[i] : racc[g][i] + v[i];
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      }
    } else {
      mbar_wait(&tmem_full[acc], (it >> 1) & 1u);
      fence_after_sync();
      ++it;
    }
#pragma unroll
    for (int g = 0; g < (CV_BN / 2 + 31) / 32; ++g) {
      const int c = (CV_BN / 2) * chalf + 32 * g;
      if (c >= (CV_BN / 2) * (chalf + 1)) continue;
      float vloc[CHUNKED ? 1 : 32];
      float (&v)[32] = *reinterpret_cast<float (*)[32]>(CHUNKED ? &racc[g][0] : &vloc[0]);     // CHUNKED: work in place on the promoted sums
      if constexpr (!CHUNKED) {
        tmem_ld32(tmem_d + acc * CV_BN + ((uint32_t)(32 * q) << 16) + (uint32_t)c, v);
        if (c + 32 >= (CV_BN / 2) * (chalf + 1)) {        // last read of this accumulator: hand it back to the MMA warp
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
      }
      const int ch = n0 + c;
      if (ch >= a.Cout) continue;
      if (a.epi == 1 || a.epi == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + (a.bias ? __ldg(a.bias + ch + i) : 0.0f), 0.0f);
      }
      if (a.epi == 2) {
        // 2x2 max pool: window partners are lane^1 (x) and lane^TW (y); first maximum in (y,x) scan order wins
        // ties, like torch's max_pool2d backward.  Post-ReLU values are >= 0, so their bit patterns order like
        // unsigned integers: the window index rides in the two lowest mantissa bits (as 3 - w, so that the earlier
        // position wins a tie) and ONE shuffle + max per partner does value and arg-max together; the 2^-22
        // relative truncation disappears in the bf16 rounding of the stored activation.
        const unsigned int w = (unsigned int)(((yy & 1) << 1) | (xx & 1));
        unsigned int arg_pack[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) arg_pack[i] = 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          unsigned int key = (__float_as_uint(v[i]) & 0x7FFFFFFCu) | (3u - w);     // (sign bit cleared: fmaxf may return -0)
          key = max(key, __shfl_xor_sync(0xffffffffu, key, 1));
          key = max(key, __shfl_xor_sync(0xffffffffu, key, a.TW));
          v[i] = __uint_as_float(key & ~3u);
          arg_pack[i >> 2] |= (3u - (key & 3u)) << (8 * (i & 3));
        }
        if (valid && w == 0) {
          const size_t pix = ((size_t)b * a.oHb + (y >> 1) + a.ooff) * a.oWb + (x >> 1) + a.ooff;
          __nv_bfloat16* dst = a.out + pix * a.Cout + ch;
#pragma unroll
          for (int i = 0; i < 32; i += 8) store8_planes(dst + i, a.out_plane, a.planes, v + i);
          if (a.pool_arg) {
            const size_t ppix = ((size_t)b * (a.Ho >> 1) + (y >> 1)) * (a.Wo >> 1) + (x >> 1);
            uint4* ad = reinterpret_cast<uint4*>(a.pool_arg + ppix * a.Cout + ch);
            ad[0] = make_uint4(arg_pack[0], arg_pack[1], arg_pack[2], arg_pack[3]);
            ad[1] = make_uint4(arg_pack[4], arg_pack[5], arg_pack[6], arg_pack[7]);
          }
        }
      } else if (valid) {
        const size_t pix = ((size_t)b * a.oHb + y + a.ooff) * a.oWb + x + a.ooff;
        __nv_bfloat16* dst = a.out + pix * a.Cout + ch;
        if (a.epi == 3) {
          const __nv_bfloat16* ref = a.relu_ref + (((size_t)b * a.rHb + y + a.roff) * a.rWb + x + a.roff) * a.Cout + ch;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const uint4 rv = *reinterpret_cast<const uint4*>(ref + i);
            const unsigned int rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // positive bf16 <=> sign bit clear and magnitude non-zero
              if (!((rw[j] & 0x7FFFu) != 0 && !(rw[j] & 0x8000u))) v[i + 2 * j] = 0.0f;
              if (!(((rw[j] >> 16) & 0x7FFFu) != 0 && !((rw[j] >> 16) & 0x8000u))) v[i + 2 * j + 1] = 0.0f;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) store8_planes(dst + i, a.out_plane, a.planes, v + i);
      }
    }
    }   // items
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, 2 * CV_BN);
}

// psi [Fo,Fi,4,3,3] fp32 -> Wmat [(o,r)][tap][(i,s)] bf16 (forward) and, if wt != NULL, the
// backward-data matrix Wt [(i,s)][tap'][(o,r)] with tap' = 8 - tap (flipped filter).
__device__ __forceinline__ void rot_src(int r, int y, int x, int& ys, int& xs) {
  // torch.rot90(w, r, dims=(-2,-1)): r quarter turns counter-clockwise
  switch (r & 3) {
    case 0: ys = y; xs = x; break;
    case 1: ys = x; xs = 2 - y; break;
    case 2: ys = 2 - y; xs = 2 - x; break;
    default: ys = 2 - x; xs = y; break;
  }
}
__global__ void expand_reg_reg_kernel(const float* __restrict__ psi, int Fo, int Fi, __nv_bfloat16* __restrict__ wmat,
                                      __nv_bfloat16* __restrict__ wt, int planes) {
  const int Cout = Fo * 4, Cin = Fi * 4;
  const long long total = (long long)Cout * 9 * Cin;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long rr = e;
    const int ci = (int)(rr % Cin); rr /= Cin;
    const int tap = (int)(rr % 9); rr /= 9;
    const int co = (int)rr;
    const int o = co >> 2, r = co & 3, i = ci >> 2, s = ci & 3;
    const int y = tap / 3, x = tap - 3 * y;
    int ys, xs;
    rot_src(r, y, x, ys, xs);
    const float w = psi[((((size_t)o * Fi + i) * 4 + ((s - r) & 3)) * 3 + ys) * 3 + xs];
    const size_t te = ((size_t)ci * 9 + (8 - tap)) * Cout + co;
    store_planes(wmat + e, (size_t)total, planes, w);          // planes stacked behind each other
    if (wt) store_planes(wt + te, (size_t)total, planes, w);
  }
}
__global__ void expand_bias_kernel(const float* __restrict__ bias_f, int F, float* __restrict__ bias_ch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < F * 4) bias_ch[i] = bias_f[i >> 2];
}

// Layer 0 (trivial -> regular, Cin = 2: heightmap + tiled gripper state), fused with ReLU and the
// 2x2 max pool.  K = 18 is far too small for the tensor pipe (and the layer is 1.3 % of the FLOPs):
// direct convolution, one thread per pooled pixel x 16 output channels, weights in shared memory.
// obs [B,1,128,128] fp32, state [B] fp32, psi [16,2,3,3], bias [16] -> out [B,66,66,64] bf16 interior.
// PLAIN (base_encoder layer 0, src/nets/base_cnns.py:25): the 16 filters are used as they are and land in channels 0..15 of
// the 64-channel buffer (the other 48 stay zero from allocation), one 16-channel group per pooled pixel instead of four.
template <bool PLAIN>
__global__ void __launch_bounds__(256)
conv0_direct_kernel(const float* __restrict__ obs, const float* __restrict__ state, const float* __restrict__ psi,
                    const float* __restrict__ bias_f, int B, __nv_bfloat16* __restrict__ out, int planes,
                    unsigned char* __restrict__ pool_arg) {
  // weights of a PAIR of output channels interleaved: [pair][18 taps + bias + pad][2], read as ten warp-uniform
  // LDS.128 and fed to FFMA2 (the two halves are the two channels)
  __shared__ __align__(16) float sW[32][20][2];
  for (int e = threadIdx.x; e < 64 * 20; e += blockDim.x) {
    const int co = e / 20, rem = e - co * 20;
    float v = 0.0f;
    if (rem < 18) {
      const int ci = rem / 9, tap = rem - ci * 9;
      if (PLAIN) {
        v = co < 16 ? psi[(co * 2 + ci) * 9 + tap] : 0.0f;
      } else {
        const int o = co >> 2, r = co & 3, y = tap / 3, x = tap - 3 * y;
        int ys, xs;
        rot_src(r, y, x, ys, xs);
        v = psi[((o * 2 + ci) * 3 + ys) * 3 + xs];
      }
    } else if (rem == 18) {
      v = PLAIN ? (co < 16 ? bias_f[co] : 0.0f) : bias_f[co >> 2];
    }
    sW[co >> 1][rem][co & 1] = v;
  }
  __syncthreads();
  const long long total = (long long)B * 64 * 64 * (PLAIN ? 1 : 4);       // pooled pixels x channel groups of 16
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    // px fastest, then the channel group: a warp shares its weights (broadcast loads) and reads adjacent pixels
    long long rr = e;
    const int px = (int)(rr & 63); rr >>= 6;
    int cg = 0;
    if (!PLAIN) { cg = (int)(rr & 3); rr >>= 2; }
    const int py = (int)(rr & 63); rr >>= 6;
    const int b = (int)rr;
    const float st = state[b];
    const float* img = obs + (size_t)b * 128 * 128;
    // 4x4 input patch (rows 2py-1 .. 2py+2) of both channels, zero padded
    float p0[4][4], p1[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int yy = 2 * py - 1 + i, xx = 2 * px - 1 + j;
        const bool in = yy >= 0 && yy < 128 && xx >= 0 && xx < 128;
        p0[i][j] = in ? __ldg(img + yy * 128 + xx) : 0.0f;
        p1[i][j] = in ? st : 0.0f;
      }
    float bestv[16];
    unsigned int argp[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int cp = 0; cp < 8; ++cp) {
      float2 wv[20];
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const float4 w4 = *reinterpret_cast<const float4*>(&sW[cg * 8 + cp][2 * q][0]);
        wv[2 * q] = make_float2(w4.x, w4.y); wv[2 * q + 1] = make_float2(w4.z, w4.w);
      }
      float2 best = make_float2(0.0f, 0.0f);
      int bw0 = 0, bw1 = 0;
#pragma unroll
      for (int wy = 0; wy < 2; ++wy)
#pragma unroll
        for (int wx = 0; wx < 2; ++wx) {
          float2 acc = wv[18];
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t - 3 * dy;
            const float a0 = p0[wy + dy][wx + dx], a1 = p1[wy + dy][wx + dx];
            acc = __ffma2_rn(wv[t], make_float2(a0, a0), acc);
            acc = __ffma2_rn(wv[9 + t], make_float2(a1, a1), acc);
          }
          acc.x = fmaxf(acc.x, 0.0f); acc.y = fmaxf(acc.y, 0.0f);
          const int w = wy * 2 + wx;
          if (w == 0 || acc.x > best.x) { best.x = acc.x; bw0 = w; }
          if (w == 0 || acc.y > best.y) { best.y = acc.y; bw1 = w; }
        }
      bestv[2 * cp] = best.x; bestv[2 * cp + 1] = best.y;
      const int c = 2 * cp;
      if ((c & 3) == 0) argp[c >> 2] = 0u;
      argp[c >> 2] |= ((unsigned int)bw0 << (8 * (c & 3))) | ((unsigned int)bw1 << (8 * ((c + 1) & 3)));
    }
    const size_t pix = ((size_t)b * 66 + py + 1) * 66 + px + 1;
    const size_t plane = (size_t)B * 66 * 66 * 64;
    store8_planes(out + pix * 64 + cg * 16, plane, planes, bestv);
    store8_planes(out + pix * 64 + cg * 16 + 8, plane, planes, bestv + 8);
    if (pool_arg) {
      const size_t ppix = ((size_t)b * 64 + py) * 64 + px;
      *reinterpret_cast<uint4*>(pool_arg + ppix * 64 + cg * 16) = make_uint4(argp[0], argp[1], argp[2], argp[3]);
    }
  }
}

}  // namespace tc
}  // namespace aur

extern "C" int aur_equiv_expand_regular(const float* psi, int32_t Fo, int32_t Fi, const float* bias_f, void* wmat, void* wt,
                                        float* bias_ch, void* stream) {
  using namespace aur;
  using namespace aur::tc;
  if (!psi || !wmat || Fo <= 0 || Fi <= 0) { set_error("aur_equiv_expand_regular: bad arguments"); return AUR_ERR_ARG; }
  const long long total = (long long)Fo * 4 * 9 * Fi * 4;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  expand_reg_reg_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(psi, Fo, Fi, (__nv_bfloat16*)wmat, (__nv_bfloat16*)wt, tc_planes());
  AUR_LAUNCH_OK("expand_reg_reg_kernel");
  if (bias_f && bias_ch) {
    expand_bias_kernel<<<(Fo * 4 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bias_f, Fo, bias_ch);
    AUR_LAUNCH_OK("expand_bias_kernel");
  }
  return 0;
}

extern "C" int aur_equiv_conv0(const float* obs, const float* state, const float* psi, const float* bias_f, int32_t B,
                               void* out, uint8_t* pool_arg, void* stream) {
  using namespace aur;
  using namespace aur::tc;
  if (!obs || !state || !psi || !bias_f || !out || B <= 0) { set_error("aur_equiv_conv0: bad arguments"); return AUR_ERR_ARG; }
  const long long total = (long long)B * 64 * 64 * 4;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  conv0_direct_kernel<false><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(obs, state, psi, bias_f, B, (__nv_bfloat16*)out, tc_planes(), pool_arg);
  AUR_LAUNCH_OK("conv0_direct_kernel");
  return 0;
}

extern "C" int aur_plain_conv0(const float* obs, const float* state, const float* weight, const float* bias, int32_t B,
                               void* out, uint8_t* pool_arg, void* stream) {
  using namespace aur;
  using namespace aur::tc;
  if (!obs || !state || !weight || !bias || !out || B <= 0) { set_error("aur_plain_conv0: bad arguments"); return AUR_ERR_ARG; }
  const long long total = (long long)B * 64 * 64;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  conv0_direct_kernel<true><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(obs, state, weight, bias, B, (__nv_bfloat16*)out, tc_planes(), pool_arg);
  AUR_LAUNCH_OK("conv0_direct_kernel<plain>");
  return 0;
}

// AUR_CONV_SWAP=0 keeps the pixel-major epilogue everywhere (A/B comparison)
static bool conv_swap_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AUR_CONV_SWAP"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

extern "C" int aur_conv3x3_bf16(const aur_conv_args* args, void* stream) {
  using namespace aur;
  using namespace aur::tc;
  if (!args || !args->in || !args->wmat || !args->out) { set_error("aur_conv3x3_bf16: null argument"); return AUR_ERR_ARG; }
  const aur_conv_args& c = *args;
  if (c.B <= 0 || c.Hb < 3 || c.Wb < 3 || c.Cin % 64 != 0 || c.Cout % 8 != 0 || c.Cin <= 0 || c.Cout <= 0) {
    set_error("aur_conv3x3_bf16: needs Cin %% 64 == 0, Cout %% 8 == 0 (got Cin %d Cout %d)", c.Cin, c.Cout);
    return AUR_ERR_UNSUPPORTED;
  }
  ConvDev d;
  d.B = c.B; d.Hb = c.Hb; d.Wb = c.Wb; d.Ho = c.Hb - 2; d.Wo = c.Wb - 2; d.Cin = c.Cin; d.Cout = c.Cout;
  d.epi = c.epilogue; d.oHb = c.out_Hb; d.oWb = c.out_Wb; d.ooff = c.out_off; d.bias = c.bias;
  const int P = tc_planes();
  d.nterm = tc_terms(P);
  d.planes = P;
  {
    static int chunk = -1;                 // AUR_CONV_CHUNK: MMA groups per promoted chunk (default CV_CHUNK = 8 -> 32..48 steps)
    if (chunk < 0) { const char* e = getenv("AUR_CONV_CHUNK"); chunk = e ? atoi(e) : CV_CHUNK; if (chunk < 1) chunk = CV_CHUNK; }
    d.chunk = chunk;
  }
  d.out_plane = (size_t)c.B * c.out_Hb * c.out_Wb * c.Cout;
  d.out = (__nv_bfloat16*)c.out; d.pool_arg = c.pool_arg; d.relu_ref = (const __nv_bfloat16*)c.relu_ref; d.rHb = c.ref_Hb; d.rWb = c.ref_Wb; d.roff = c.ref_off;
  if (c.epilogue == 3 && !c.relu_ref) { set_error("aur_conv3x3_bf16: epilogue 3 needs relu_ref"); return AUR_ERR_ARG; }
  if (c.epilogue == 2 && ((d.Ho | d.Wo) & 1)) { set_error("aur_conv3x3_bf16: pooling needs even output size"); return AUR_ERR_ARG; }
  // pixel tile: 128 = NIMG x TH x TW with TW in {8,16} so a 2x2 pool window stays inside a warp
  d.TW = d.Wo > 8 ? 16 : 8;
  d.TH = 8;
  d.NIMG = 128 / (d.TW * d.TH);
  d.tiles_x = (d.Wo + d.TW - 1) / d.TW;
  d.tiles_y = (d.Ho + d.TH - 1) / d.TH;
  const long long img_groups = (c.B + d.NIMG - 1) / d.NIMG;
  CUtensorMap tmA, tmB;
  // planes (hi / mid stacks) are the outermost tensor-map dimension of both operands
  const uint64_t dA[5] = {(uint64_t)c.Cin, (uint64_t)c.Wb, (uint64_t)c.Hb, (uint64_t)c.B, (uint64_t)P};
  const uint64_t sA[4] = {(uint64_t)c.Cin * 2, (uint64_t)c.Wb * c.Cin * 2, (uint64_t)c.Hb * c.Wb * c.Cin * 2,
                          (uint64_t)c.B * c.Hb * c.Wb * c.Cin * 2};
  const uint32_t bA[5] = {CV_BK, (uint32_t)d.TW, (uint32_t)d.TH, (uint32_t)d.NIMG, 1};
  const uint64_t dB[3] = {(uint64_t)9 * c.Cin, (uint64_t)c.Cout, (uint64_t)P};
  const uint64_t sB[2] = {(uint64_t)9 * c.Cin * 2, (uint64_t)9 * c.Cin * c.Cout * 2};
  // pooled layers with a shallow contraction (K = 9 x 64 / 9 x 128) are epilogue-bound: channel-major accumulator there
  // (split planes: the pixel-major kernels only - the resident-weight variant has no room for two weight planes)
  const bool swap = c.epilogue == 2 && c.Cin == 64 && P == 1 && conv_swap_enabled();       // Cin 128 (layer 2) is faster on 256-wide tiles
  const bool resb = swap && c.Cout <= 128;
  const bool chunked = P > 1;                     // multi-plane precisions promote the accumulator chunk by chunk (BN <= 128)
  const bool wide = !swap && !chunked && c.Cout % 256 == 0;
  const bool narrow = !swap && c.Cout <= 64;     // e.g. the backward-data convolution into the 64-channel layer-0 output: N = 64 MMAs
  const int BN = wide ? 256 : (narrow ? 64 : 128);
  const uint32_t bB[3] = {CV_BK, (uint32_t)BN, 1};
  int rc;
  if ((rc = make_tensor_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, c.in, dA, sA, bA))) return rc;
  if ((rc = make_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, c.wmat, dB, sB, bB))) return rc;
  static DeviceOnce attr;
  if (attr.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CvCfg<128>::SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CvCfg<256>::SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CvCfg<64>::SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CvCfg<128>::SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<128, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CV_RESB_SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<128, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CvCfg<128>::SMEM));
    AUR_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<64, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CvCfg<64>::SMEM));
    attr.done();
  }
  d.pix_tiles = (int)(img_groups * d.tiles_y * d.tiles_x);
  d.n_tiles = (c.Cout + BN - 1) / BN;
  const long long items = (long long)d.pix_tiles * d.n_tiles;
  const unsigned grid = (unsigned)(items < sm_count() ? items : sm_count());
  if (chunked && narrow) conv_igemm_kernel<64, false, false, true><<<grid, CV_THREADS, CvCfg<64>::SMEM, (cudaStream_t)stream>>>(tmA, tmB, d);
  else if (chunked) conv_igemm_kernel<128, false, false, true><<<grid, CV_THREADS, CvCfg<128>::SMEM, (cudaStream_t)stream>>>(tmA, tmB, d);
  else if (resb) conv_igemm_kernel<128, true, true><<<grid, CV_THREADS, CV_RESB_SMEM, (cudaStream_t)stream>>>(tmA, tmB, d);
  else if (swap) conv_igemm_kernel<128, true><<<grid, CV_THREADS, CvCfg<128>::SMEM, (cudaStream_t)stream>>>(tmA, tmB, d);
  else if (wide) conv_igemm_kernel<256><<<grid, CV_THREADS, CvCfg<256>::SMEM, (cudaStream_t)stream>>>(tmA, tmB, d);
  else if (narrow) conv_igemm_kernel<64><<<grid, CV_THREADS, CvCfg<64>::SMEM, (cudaStream_t)stream>>>(tmA, tmB, d);
  else conv_igemm_kernel<128><<<grid, CV_THREADS, CvCfg<128>::SMEM, (cudaStream_t)stream>>>(tmA, tmB, d);
  AUR_LAUNCH_OK("conv_igemm_kernel");
  return 0;
}
