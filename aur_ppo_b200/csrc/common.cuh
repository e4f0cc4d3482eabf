// Shared device/host helpers for libaurppo.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/aur_ppo.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libaurppo is written for sm_100a (B200) only"
#endif

namespace aur {

// ---- host-side error plumbing -------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int cuda_fail(cudaError_t e, const char* what);

#define AUR_CUDA_OK(expr)                                        \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return aur::cuda_fail(_e, #expr);     \
  } while (0)

#define AUR_LAUNCH_OK(name)                                      \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return aur::cuda_fail(_e, name);      \
    aur::count_launch();                                         \
  } while (0)

int sm_count();

// "do this once per device": function attributes (dynamic shared memory limits) are per device, a process-wide
// flag would leave every device but the first one without them.
struct DeviceOnce {
  uint64_t mask = 0;
  int dev = -1;
  bool first() {
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { dev = -1; return true; }
    return !((mask >> dev) & 1ull);
  }
  void done() { if (dev >= 0) mask |= 1ull << dev; }
};

// Last-CTA ticket counters owned by the LIBRARY (zeroed at allocation, reset by the kernel that used them), one
// block of 1024 words per (device, stream): kernels of one stream run in order, so they can share a block; a
// caller-provided workspace would have to be zero-initialised by contract (it is not part of the ABI contract).
// Returns nullptr on failure (error text set).
unsigned int* stream_tickets(cudaStream_t s);
// Library-owned, grow-only scratch per (device, stream) for entry points without a workspace argument (api.cu); nullptr on failure.
void* stream_scratch(cudaStream_t s, size_t bytes);

// ---- mbarrier / bulk-async (TMA 1-D) PTX wrappers -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (SASS: UBLKCP), completion counted on `bar`.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- warp helpers -------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- Philox4x32-10 (counter-based sampler; Salmon et al. SC'11) -----------------
struct Philox {
  uint32_t c[4];
};
__host__ __device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                         uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  Philox o;
  o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
  return o;
}
// 24-bit uniform in [0,1)
__host__ __device__ __forceinline__ float u01_24(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

}  // namespace aur
