// Exchange areas of the data-parallel update: plain cudaMalloc memory shared between the
// ranks of one node with CUDA IPC, so that the gradient all-reduce is done by OUR kernels over NVLink / NVSwitch
// peer stores (grad_reduce_kernel pushes, adam_kernel gathers; see update.cuh) instead of a library collective.
#include <string.h>

#include "update.cuh"

extern "C" int64_t aur_dp_area_bytes(const aur_policy_desc* desc) {
  if (!desc) return AUR_ERR_ARG;
  return aur::dp_area_bytes(aur::policy_param_count(*desc));
}

extern "C" int aur_dp_alloc(int64_t bytes, void** area_out, void* ipc_handle_out) {
  using namespace aur;
  if (bytes <= 0 || !area_out || !ipc_handle_out) { set_error("aur_dp_alloc: bad arguments"); return AUR_ERR_ARG; }
  static_assert(sizeof(cudaIpcMemHandle_t) == AUR_DP_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  AUR_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  AUR_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
  AUR_CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  AUR_CUDA_OK(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle_out, &h, sizeof(h));
  *area_out = p;
  return 0;
}

extern "C" int aur_dp_open(const void* ipc_handle, void** area_out) {
  using namespace aur;
  if (!ipc_handle || !area_out) { set_error("aur_dp_open: bad arguments"); return AUR_ERR_ARG; }
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  void* p = nullptr;
  AUR_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *area_out = p;
  return 0;
}

extern "C" int aur_dp_close(void* area) {
  using namespace aur;
  if (!area) return 0;
  AUR_CUDA_OK(cudaIpcCloseMemHandle(area));
  return 0;
}

extern "C" int aur_dp_free(void* area) {
  using namespace aur;
  if (!area) return 0;
  AUR_CUDA_OK(cudaFree(area));
  return 0;
}

// out[6] = {spin ns on gradient flags (summed over spinning threads), spins, the same for moment flags, wall ns the Adam kernels
// waited for the slowest peer, Adam launches}; reset != 0 zeroes them
extern "C" int aur_dp_wait_stats(void* area, uint64_t* out6, int32_t reset, void* stream) {
  using namespace aur;
  if (!area || !out6) { set_error("aur_dp_wait_stats: bad arguments"); return AUR_ERR_ARG; }
  unsigned char* w = static_cast<unsigned char*>(area) + DP_OFF_WAIT;
  AUR_CUDA_OK(cudaMemcpyAsync(out6, w, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  if (reset) AUR_CUDA_OK(cudaMemsetAsync(w, 0, 6 * sizeof(uint64_t), (cudaStream_t)stream));
  AUR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

// 0 = healthy, 1 = a kernel gave up waiting for a peer (its results are invalid)
extern "C" int aur_dp_status(const void* area, void* stream) {
  using namespace aur;
  if (!area) { set_error("aur_dp_status: null area"); return AUR_ERR_ARG; }
  uint32_t v = 0;
  AUR_CUDA_OK(cudaMemcpyAsync(&v, static_cast<const unsigned char*>(area) + DP_OFF_STATUS, sizeof(v), cudaMemcpyDeviceToHost,
                              (cudaStream_t)stream));
  AUR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return (int)v;
}
