// Error reporting / device info for the C-ABI.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void sgg_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long sgg_launch_count(void) { return g_launches.load(); }

int sgg_set_err(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int sgg_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

extern "C" int sgg_abi_version(void) { return SGG_ABI_VERSION; }
extern "C" const char *sgg_last_error(void) { return g_err; }
extern "C" int sgg_device_info(int out[4]) {
  int dev = 0;
  SGG_CUDA_TRY(cudaGetDevice(&dev));
  SGG_CUDA_TRY(cudaDeviceGetAttribute(&out[0], cudaDevAttrMultiProcessorCount, dev));
  SGG_CUDA_TRY(cudaDeviceGetAttribute(&out[1], cudaDevAttrComputeCapabilityMajor, dev));
  SGG_CUDA_TRY(cudaDeviceGetAttribute(&out[2], cudaDevAttrComputeCapabilityMinor, dev));
  SGG_CUDA_TRY(cudaDeviceGetAttribute(&out[3], cudaDevAttrL2CacheSize, dev));
  return 0;
}

// ---- fork/join helper: one side stream + a ring of timing-less events per process (capturable) ----
namespace sgg {
// side streams per caller stream (so pipelined callers on different streams do not serialise on them);
// idx 0 = object branch, idx 1 = the P = V W_ih^T GEMM of the message-passing loop
cudaStream_t side_stream(cudaStream_t main, int idx) {
  static cudaStream_t keys[32];
  static int kidx[32];
  static cudaStream_t vals[32];
  static int n = 0;
  for (int i = 0; i < n; ++i)
    if (keys[i] == main && kidx[i] == idx) return vals[i];
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  if (n < 32) { keys[n] = main; kidx[n] = idx; vals[n] = s; ++n; return s; }
  keys[0] = main; kidx[0] = idx; vals[0] = s;      // table full: recycle slot 0 (the old stream is leaked, bounded)
  return s;
}
cudaEvent_t next_event() {
  static cudaEvent_t ring[256];
  static bool init = false;
  static unsigned idx = 0;
  if (!init) {
    for (int i = 0; i < 256; ++i) cudaEventCreateWithFlags(&ring[i], cudaEventDisableTiming);
    init = true;
  }
  return ring[(idx++) & 255];
}
// order: everything enqueued on `from` so far happens-before whatever is enqueued on `to` next
int stream_order(cudaStream_t from, cudaStream_t to) {
  cudaEvent_t e = next_event();
  SGG_CUDA_TRY(cudaEventRecord(e, from));
  SGG_CUDA_TRY(cudaStreamWaitEvent(to, e, 0));
  return 0;
}
}  // namespace sgg
