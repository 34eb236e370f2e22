// ABI version + thread-local error string.
#include <stdarg.h>

#include "common.cuh"

namespace ptrec {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return PTREC_ECUDA;
}

}  // namespace ptrec

extern "C" int ptrec_abi_version(void) { return PTREC_ABI_VERSION; }
extern "C" const char* ptrec_last_error(void) { return ptrec::g_err; }
