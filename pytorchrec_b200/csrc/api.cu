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
