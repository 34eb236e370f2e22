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

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return PTREC_ECUDA;
}

}  // namespace ptrec

extern "C" int ptrec_abi_version(void) { return PTREC_ABI_VERSION; }
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
