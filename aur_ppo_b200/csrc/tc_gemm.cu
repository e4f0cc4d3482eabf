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
