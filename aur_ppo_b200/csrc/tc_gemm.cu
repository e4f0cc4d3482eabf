// tcgen05 GEMM: C[M,N] (fp32) = A[M,K] * B[N,K]^T, A and B bf16 row-major (K contiguous).
// TMA (128-B swizzle) -> 4-stage shared-memory ring -> tcgen05.mma kind::f16, accumulator in TMEM
// -> tcgen05.ld epilogue.  One CTA per 128 x BN output tile; warp 0 = TMA producer, warp 1 = MMA
// issuer (one elected thread), warp 2 = TMEM allocator, warps 4-7 = epilogue.
// This is the dense-contraction core the equivariant encoder's convolutions are built from.
#include "tc.cuh"

namespace aur {
namespace tc {

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return AUR_ERR_UNSUPPORTED; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return AUR_ERR_ARG; }
  return 0;
}

static thread_local int g_tc_planes = 1;
int tc_planes() { return g_tc_planes; }

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 64, GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2, GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;
constexpr size_t GEMM_SMEM = (size_t)GEMM_STAGES * (GEMM_A_BYTES + GEMM_B_BYTES) + 1024 + 256;

// `chunk` stages per accumulator hand-over: the tensor core adds each K = 16 step into its fp32 accumulator with truncation
// (~3e-8 per step: 1.5e-5 over K = 4608 x 3 products), so in the multi-plane precisions the MMA warp alternates between two TMEM
// accumulators every `chunk` stages and the eight epilogue warps promote every finished chunk into fp32 registers
// (round-to-nearest adds); single-plane bf16 passes chunk = all stages (one hand-over, as before).
constexpr int GEMM_THREADS = 384;     // warps: 0 TMA, 1 MMA, 2 TMEM allocator, 4..11 epilogue (lane quarter x column half)
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                    int M, int N, int K, int ldc, int nterm, int chunk) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
  unsigned char* sA = smem;
  unsigned char* sB = smem + GEMM_STAGES * GEMM_A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + GEMM_STAGES * GEMM_B_BYTES);
  uint64_t* empty = full + GEMM_STAGES;
  uint64_t* tmem_full = empty + GEMM_STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * GEMM_BN;
  const int nkb = ((K + GEMM_BK - 1) / GEMM_BK) * nterm;     // multi-plane operands: the plane products are extra K steps
  if (chunk <= 0 || chunk > nkb) chunk = nkb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    mbar_fence_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * GEMM_BN);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % GEMM_STAGES;
      const uint32_t ph = (kb / GEMM_STAGES) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      mbar_arrive_expect_tx(&full[s], GEMM_A_BYTES + GEMM_B_BYTES);
      const int kk = kb / nterm, term = kb - kk * nterm;
      tma_load_3d(sA + s * GEMM_A_BYTES, &tmA, kk * GEMM_BK, m0, term_plane_a(term), &full[s]);
      tma_load_3d(sB + s * GEMM_B_BYTES, &tmB, kk * GEMM_BK, n0, term_plane_b(term), &full[s]);
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = instr_desc(FMT_BF16, GEMM_BM, GEMM_BN, 0, 0);
    uint32_t it = 0;
    for (int kb0 = 0; kb0 < nkb; kb0 += chunk, ++it) {
      const uint32_t acc = it & 1u;
      mbar_wait(&tmem_empty[acc], ((it >> 1) & 1u) ^ 1u);
      fence_after_sync();
      const int kb1 = kb0 + chunk < nkb ? kb0 + chunk : nkb;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = kb % GEMM_STAGES;
        const uint32_t ph = (kb / GEMM_STAGES) & 1u;
        mbar_wait(&full[s], ph);
        fence_after_sync();
        const uint64_t ad = smem_desc_k_sw128(sA + s * GEMM_A_BYTES), bd = smem_desc_k_sw128(sB + s * GEMM_B_BYTES);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k)          // UMMA_K = 16 bf16 = 32 B: advance the start address by 2 x 16 B
          mma_f16(tmem_d + acc * GEMM_BN, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0);
        mma_commit(&empty[s]);
      }
      mma_commit(&tmem_full[acc]);
    }
  } else if (warp >= 4) {
    const int q = (warp - 4) & 3, chalf = (warp - 4) >> 2;
    const int row = m0 + 32 * q + lane;
    float racc[2][32];
    uint32_t it = 0;
    for (int kb0 = 0; kb0 < nkb; kb0 += chunk, ++it) {
      const uint32_t acc = it & 1u;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1u);
      fence_after_sync();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float v[32];
        tmem_ld32(tmem_d + acc * GEMM_BN + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * chalf + 32 * g), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) racc[g][i] = kb0 == 0 ? v[i] : racc[g][i] + v[i];
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (row < M) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int c = 64 * chalf + 32 * g;
        float* dst = C + (size_t)row * ldc + n0 + c;
        if (n0 + c + 32 <= N && (ldc & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(racc[g][i], racc[g][i + 1], racc[g][i + 2], racc[g][i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (n0 + c + i < N) dst[i] = racc[g][i];
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, 2 * GEMM_BN);
}

}  // namespace tc
}  // namespace aur

extern "C" int aur_tc_set_precision(int planes) {
  if (planes < 1 || planes > 3) { aur::set_error("aur_tc_set_precision: planes must be 1 (bf16), 2 (hi + mid) or 3 (hi + mid + lo)"); return AUR_ERR_ARG; }
  const int prev = aur::tc::g_tc_planes;
  aur::tc::g_tc_planes = planes;
  return prev;
}
extern "C" int aur_tc_get_precision(void) { return aur::tc::g_tc_planes; }

namespace aur {
namespace tc {
// C[M,N] (fp32, row stride ldc) = A[M,K] * B[N,K]^T for operand stacks of `planes` bf16 planes, a_plane / b_plane ELEMENTS apart
// (the internal form of aur_tc_gemm_bf16: explicit precision and plane strides; update_wide.cu runs sub-batches out of larger buffers)
int launch_tc_gemm(int64_t M, int64_t N, int64_t K, const void* A, size_t a_plane, const void* B, size_t b_plane, float* C, int ldc,
                   int planes, cudaStream_t stream) {
  const int P = planes;
  CUtensorMap tmA, tmB;
  const uint64_t dA[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)P}, dB[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)P};
  const uint64_t stA[2] = {(uint64_t)K * 2, (uint64_t)a_plane * 2}, stB[2] = {(uint64_t)K * 2, (uint64_t)b_plane * 2};
  const uint32_t boxA[3] = {GEMM_BK, GEMM_BM, 1}, boxB[3] = {GEMM_BK, GEMM_BN, 1};
  int rc;
  if ((rc = make_tensor_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, dA, stA, boxA))) return rc;
  if ((rc = make_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, B, dB, stB, boxB))) return rc;
  static DeviceOnce attr;
  if (attr.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(tc_gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    attr.done();
  }
  dim3 grid((unsigned)((M + GEMM_BM - 1) / GEMM_BM), (unsigned)((N + GEMM_BN - 1) / GEMM_BN));
  tc_gemm_bf16_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(tmA, tmB, C, (int)M, (int)N, (int)K, ldc, tc_terms(P), P > 1 ? 8 : 0);
  AUR_LAUNCH_OK("tc_gemm_bf16_kernel");
  return 0;
}
}  // namespace tc
}  // namespace aur

namespace aur {
namespace tc {
// ---- persistent GEMM for the wide-policy paths (update_wide.cu, rollout_wide.cu): C[M][H] = A[M][H] B[H][H]^T, H = 64 KB_, with
// very tall M and K = H of only 128 / 256.  The per-tile kernel above spends most of such a tile on its own prologue, operand
// re-loads per plane product and an epilogue that cannot overlap anything (242 us for 536 MB at H = 256).  Here a CTA keeps its
// NB output columns' worth of B (all planes, all of K) resident in shared memory and walks 128-row tiles: a stage holds every
// plane of one 64-wide K block of A, all plane products are issued from it, two TMEM accumulators let the four epilogue warps
// store tile t (transposed through padded patches: every store is a 128-byte row segment) while tile t + 1 multiplies.
template <int P, int KB_, int NB>
struct SkCfg {
  static constexpr int A_ATOM = 128 * 128, B_ATOM = NB * 128;
  static constexpr int B_BYTES = P * KB_ * B_ATOM, STAGE = P * A_ATOM, STAGES = 2, PATCH = 32 * 33 * 4;
  static constexpr size_t SMEM = (size_t)B_BYTES + STAGES * STAGE + 4 * PATCH + 1024 + 256;
};
// NT: plane products issued per K step, in the order of term_plane_a / term_plane_b - 3 (hi*hi, hi*mid, mid*hi) or 4 (+ mid*mid,
// the rollout's log-prob precision) for two planes, 6 for three
template <int P, int KB_, int NB, int NT>
__global__ void __launch_bounds__(192, 1)
skinny_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C, int M) {
  using Cfg = SkCfg<P, KB_, NB>;
  constexpr int H = 64 * KB_, NBLK = H / NB;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sB = smem;
  unsigned char* sA = smem + Cfg::B_BYTES;
  float* patches = reinterpret_cast<float*>(sA + Cfg::STAGES * Cfg::STAGE);
  uint64_t* bfull = reinterpret_cast<uint64_t*>(sA + Cfg::STAGES * Cfg::STAGE + 4 * Cfg::PATCH);
  uint64_t* full = bfull + 1;
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;    // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (M + 127) / 128;
  const int nb = blockIdx.x % NBLK, t0 = blockIdx.x / NBLK, tstep = gridDim.x / NBLK;

  if (threadIdx.x == 0) {
    mbar_init(bfull, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_fence_init();
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * NB < 32 ? 32 : 2 * NB);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(bfull, Cfg::B_BYTES);
    for (int p = 0; p < P; ++p)
      for (int kb = 0; kb < KB_; ++kb) tma_load_3d(sB + (p * KB_ + kb) * Cfg::B_ATOM, &tmB, 64 * kb, nb * NB, p, bfull);
    uint32_t it = 0;
    for (int tile = t0; tile < ntiles; tile += tstep)
      for (int kb = 0; kb < KB_; ++kb, ++it) {
        const int s = it % Cfg::STAGES;
        mbar_wait(&empty[s], ((it / Cfg::STAGES) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE);
        for (int p = 0; p < P; ++p) tma_load_3d(sA + s * Cfg::STAGE + p * Cfg::A_ATOM, &tmA, 64 * kb, tile * 128, p, &full[s]);
      }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = instr_desc(FMT_BF16, 128, NB, 0, 0);
    mbar_wait(bfull, 0);
    uint32_t it = 0, ti = 0;
    for (int tile = t0; tile < ntiles; tile += tstep, ++ti) {
      const uint32_t acc = ti & 1u;
      mbar_wait(&tempty[acc], ((ti >> 1) & 1u) ^ 1u);
      fence_after_sync();
      for (int kb = 0; kb < KB_; ++kb, ++it) {
        const int s = it % Cfg::STAGES;
        mbar_wait(&full[s], (it / Cfg::STAGES) & 1u);
        fence_after_sync();
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint64_t ad = smem_desc_k_sw128(sA + s * Cfg::STAGE + term_plane_a(t) * Cfg::A_ATOM);
          const uint64_t bd = smem_desc_k_sw128(sB + (term_plane_b(t) * KB_ + kb) * Cfg::B_ATOM);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16(tmem_d + acc * NB, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | t | k) != 0);
        }
        mma_commit(&empty[s]);
      }
      mma_commit(&tfull[acc]);
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    float* patch = patches + (warp - 2) * (Cfg::PATCH / 4);
    uint32_t ti = 0;
    for (int tile = t0; tile < ntiles; tile += tstep, ++ti) {
      const uint32_t acc = ti & 1u;
      mbar_wait(&tfull[acc], (ti >> 1) & 1u);
      fence_after_sync();
      const long long row0 = (long long)tile * 128 + 32 * q;
#pragma unroll 1
      for (int g = 0; g < NB / 32; ++g) {
        float v[32];
        tmem_ld32(tmem_d + acc * NB + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * g), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) patch[lane * 33 + i] = v[i];
        __syncwarp();
#pragma unroll 8
        for (int r = 0; r < 32; ++r)
          if (row0 + r < M) C[(row0 + r) * H + nb * NB + 32 * g + lane] = patch[r * 33 + lane];
        __syncwarp();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, 2 * NB < 32 ? 32 : 2 * NB);
}

template <int P, int KB_, int NB, int NT>
static int launch_skinny(int64_t M, const void* A, size_t a_plane, const void* B, size_t b_plane, float* C, cudaStream_t s) {
  using Cfg = SkCfg<P, KB_, NB>;
  constexpr int H = 64 * KB_, NBLK = H / NB;
  static_assert(Cfg::SMEM <= 232448, "skinny GEMM configuration does not fit shared memory");
  CUtensorMap tmA, tmB;
  const uint64_t dA[3] = {(uint64_t)H, (uint64_t)M, (uint64_t)P}, dB[3] = {(uint64_t)H, (uint64_t)H, (uint64_t)P};
  const uint64_t stA[2] = {(uint64_t)H * 2, (uint64_t)a_plane * 2}, stB[2] = {(uint64_t)H * 2, (uint64_t)b_plane * 2};
  const uint32_t boxA[3] = {64, 128, 1}, boxB[3] = {64, NB, 1};
  int rc;
  if ((rc = make_tensor_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, A, dA, stA, boxA))) return rc;
  if ((rc = make_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, B, dB, stB, boxB))) return rc;
  static DeviceOnce attr;
  if (attr.first()) {
    AUR_CUDA_OK(cudaFuncSetAttribute(skinny_gemm_kernel<P, KB_, NB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr.done();
  }
  const long long ntiles = (M + 127) / 128;
  long long per = sm_count() / NBLK;
  if (per < 1) per = 1;
  if (per > ntiles) per = ntiles;
  skinny_gemm_kernel<P, KB_, NB, NT><<<(unsigned)(per * NBLK), 192, Cfg::SMEM, s>>>(tmA, tmB, C, (int)M);
  AUR_LAUNCH_OK("skinny_gemm_kernel");
  return 0;
}

// C[M][H] (fp32, dense rows) = A planes [M][H] x B planes [H][H]^T for the wide-policy paths; shapes without a resident-B
// configuration fall through to the per-tile kernel
int launch_wide_gemm(int64_t M, int H, const void* A, size_t a_plane, const void* B, size_t b_plane, float* C, int planes, cudaStream_t s,
                     bool mid_mid = false) {
  if (H == 256 && planes == 2) return mid_mid ? launch_skinny<2, 4, 128, 4>(M, A, a_plane, B, b_plane, C, s) : launch_skinny<2, 4, 128, 3>(M, A, a_plane, B, b_plane, C, s);
  if (H == 256 && planes == 3) return launch_skinny<3, 4, 64, 6>(M, A, a_plane, B, b_plane, C, s);
  if (H == 128 && planes == 3) return launch_skinny<3, 2, 128, 6>(M, A, a_plane, B, b_plane, C, s);
  if (H == 128 && planes == 2 && mid_mid) return launch_skinny<2, 2, 128, 4>(M, A, a_plane, B, b_plane, C, s);
  if (mid_mid) { set_error("launch_wide_gemm: no four-product configuration for H = %d", H); return AUR_ERR_UNSUPPORTED; }
  return launch_tc_gemm(M, H, H, A, a_plane, B, b_plane, C, H, planes, s);
}
}  // namespace tc
}  // namespace aur

extern "C" int aur_tc_gemm_bf16(int64_t M, int64_t N, int64_t K, const void* A, const void* B, float* C, void* stream) {
  using namespace aur;
  if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) { set_error("aur_tc_gemm_bf16: bad arguments"); return AUR_ERR_ARG; }
  if (K % 8 != 0) { set_error("aur_tc_gemm_bf16: K must be a multiple of 8 (16-byte row pitch for TMA)"); return AUR_ERR_UNSUPPORTED; }
  return tc::launch_tc_gemm(M, N, K, A, (size_t)K * M, B, (size_t)K * N, C, (int)N, tc::tc_planes(), (cudaStream_t)stream);
}

// Debug/diagnostic: shared-window offset at which dynamic shared memory starts (the first 1 KB of the
// window is system-reserved on sm_90+; the TMEM allocator keeps its bookkeeping there).
namespace aur { namespace tc {
__global__ void smem_base_kernel(unsigned int* out) {
  extern __shared__ unsigned char dyn[];
  if (threadIdx.x == 0) out[0] = smem_u32(dyn);
}
}}
extern "C" int aur_debug_smem_base(unsigned int* out_dev, void* stream) {
  aur::tc::smem_base_kernel<<<1, 32, 1024, (cudaStream_t)stream>>>(out_dev);
  AUR_LAUNCH_OK("smem_base_kernel");
  return 0;
}
