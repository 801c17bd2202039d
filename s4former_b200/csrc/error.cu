// Error reporting for the C ABI: no exceptions cross the boundary; every entry point returns
// 0 or a negative code and s4_last_error() describes the most recent failure of this thread.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void s4_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int s4_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    s4_set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return S4_ERR_CUDA;
  }
  return S4_OK;
}

extern "C" const char* s4_last_error() { return g_err; }
extern "C" int s4_version() { return 100; }
extern "C" int s4_built_arch() {
  return 100;  // sm_100a only
}
