// Host-side plumbing shared by every entry point of libaurppo.so.
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace aur {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return -(1000 + (int)e);
}
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}


namespace {
struct TicketSlot { int dev; cudaStream_t s; unsigned int* p; };
constexpr size_t TICKET_BYTES = 4096;   // words 0, 1: single-minibatch moments / gradient reduce; 16 ..: aur_ppo_adv_moments_multi (1 + 512)
std::mutex g_ticket_mu;
std::vector<TicketSlot> g_tickets;
}  // namespace

unsigned int* stream_tickets(cudaStream_t s) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("stream_tickets: no current device"); return nullptr; }
  std::lock_guard<std::mutex> lk(g_ticket_mu);
  for (const TicketSlot& t : g_tickets)
    if (t.dev == dev && t.s == s) return t.p;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
    set_error("stream_tickets: first use of a stream inside a graph capture (run one eager call on it first)");
    return nullptr;
  }
  unsigned int* p = nullptr;
  if (cudaMalloc(&p, TICKET_BYTES) != cudaSuccess || cudaMemset(p, 0, TICKET_BYTES) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("stream_tickets: cudaMalloc failed");
    return nullptr;
  }
  g_tickets.push_back({dev, s, p});
  return p;
}

namespace {
struct ScratchSlot { int dev; cudaStream_t s; void* p; size_t bytes; };
std::vector<ScratchSlot> g_scratch;
}  // namespace

// Library-owned scratch per (device, stream) for the entry points whose C signature carries no workspace (aur_rollout at
// hidden widths that run layer by layer, rollout_wide.cu): grows on demand (cudaFree synchronises the device, so a buffer is
// never pulled from under a running kernel), never shrinks; refuses to (re)allocate inside a graph capture.
void* stream_scratch(cudaStream_t s, size_t bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("stream_scratch: no current device"); return nullptr; }
  std::lock_guard<std::mutex> lk(g_ticket_mu);
  ScratchSlot* slot = nullptr;
  for (ScratchSlot& t : g_scratch)
    if (t.dev == dev && t.s == s) slot = &t;
  if (slot && slot->bytes >= bytes) return slot->p;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
    set_error("stream_scratch: allocation inside a graph capture (run one eager call of this shape on the stream first)");
    return nullptr;
  }
  if (slot) { (void)cudaFree(slot->p); slot->p = nullptr; slot->bytes = 0; }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("stream_scratch: cudaMalloc of %zu bytes failed", bytes);
    return nullptr;
  }
  if (slot) { slot->p = p; slot->bytes = bytes; }
  else g_scratch.push_back({dev, s, p, bytes});
  return p;
}

}  // namespace aur

extern "C" int aur_abi_version(void) { return AUR_ABI_VERSION; }
extern "C" const char* aur_last_error(void) { return aur::g_err; }
extern "C" int64_t aur_launch_count(void) { return aur::g_launches; }
extern "C" void aur_launch_count_reset(void) { aur::g_launches = 0; }
