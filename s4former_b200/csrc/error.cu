// Error reporting for the C ABI: no exceptions cross the boundary; every entry point returns
// 0 or a negative code and s4_last_error() describes the most recent failure of this thread.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void s4_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void s4_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int s4_check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
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
// number of kernels this library has launched in this process (bench.py reports it)
extern "C" long long s4_launch_count() { return g_launches.load(std::memory_order_relaxed); }
