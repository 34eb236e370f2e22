// ABI version + thread-local error string.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace ptrec {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

static std::atomic<cudaStream_t> g_reduce_stream{nullptr};
void set_reduce_stream(cudaStream_t s) { g_reduce_stream.store(s, std::memory_order_relaxed); }

cudaStream_t reduce_stream_after(cudaStream_t producer) {
  cudaStream_t rs = g_reduce_stream.load(std::memory_order_relaxed);
  if (rs == nullptr || rs == producer) return producer;
  // a small ring of timing-less events: an event may be re-recorded while an older wait on it is still pending (the
  // wait captured the record it followed)
  constexpr int kEvents = 32;
  static cudaEvent_t ring[kEvents];
  static std::atomic<int> made{0}, next{0};
  if (made.load(std::memory_order_acquire) == 0) {
    int expected = 0;
    if (made.compare_exchange_strong(expected, -1)) {
      bool ok = true;
      for (int i = 0; i < kEvents; ++i) ok = ok && cudaEventCreateWithFlags(&ring[i], cudaEventDisableTiming) == cudaSuccess;
      made.store(ok ? 1 : -2, std::memory_order_release);
    }
  }
  while (made.load(std::memory_order_acquire) == -1) {
  }
  if (made.load(std::memory_order_acquire) != 1) return producer;
  cudaEvent_t ev = ring[next.fetch_add(1, std::memory_order_relaxed) % kEvents];
  if (cudaEventRecord(ev, producer) != cudaSuccess || cudaStreamWaitEvent(rs, ev, 0) != cudaSuccess) {
    cudaGetLastError();
    return producer;
  }
  return rs;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return PTREC_ECUDA;
}

}  // namespace ptrec

extern "C" int ptrec_abi_version(void) { return PTREC_ABI_VERSION; }
namespace ptrec { void set_reduce_stream(cudaStream_t s); }
extern "C" void ptrec_set_reduce_stream(void* stream) { ptrec::set_reduce_stream((cudaStream_t)stream); }
extern "C" const char* ptrec_last_error(void) { return ptrec::g_err; }
namespace ptrec { long long launches(); }
extern "C" int64_t ptrec_launch_count(void) { return (int64_t)ptrec::launches(); }

// L2 fetch granularity hint (cudaLimitMaxL2FetchGranularity: 32, 64 or 128 bytes).  Random 64-byte row
// reads are promoted to 128-byte DRAM fetches at the default setting (ncu: dram bytes = 2x algorithmic);
// lowering the limit to the row size removes the over-fetch for the gather / scatter kernels.
extern "C" int ptrec_set_l2_fetch_granularity(int32_t bytes) {
  PTREC_CHECK_ARG(bytes == 32 || bytes == 64 || bytes == 128, PTREC_EINVAL, "l2 fetch granularity must be 32, 64 or 128");
  PTREC_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
  return PTREC_OK;
}
extern "C" int ptrec_get_l2_fetch_granularity(void) {
  size_t v = 0;
  if (cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity) != cudaSuccess) return -1;
  return (int)v;
}
