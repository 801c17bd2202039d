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

// ---------------------------------------------------------------------------------------------
// per-kernel-family device timing (bench.py's roofline leg).  Event pairs are pooled; reading
// the report synchronises the device and folds all finished pairs into the per-name totals.
// ---------------------------------------------------------------------------------------------
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

namespace {
struct ProfKind {
  std::string name;
  int unit = 0;
  double ms = 0, work = 0;
  long long launches = 0;
};
struct ProfPair {
  cudaEvent_t a, b;
  int kind;
};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfKind> g_kinds;
std::vector<ProfPair> g_pending;
std::vector<ProfPair> g_free;

void prof_fold_locked() {
  if (g_pending.empty()) return;
  cudaDeviceSynchronize();
  for (auto& pr : g_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pr.a, pr.b) == cudaSuccess) g_kinds[pr.kind].ms += ms;
    g_free.push_back(pr);
  }
  g_pending.clear();
}
}  // namespace

bool s4_prof_on() { return g_prof_on.load(std::memory_order_relaxed) != 0; }

void s4_prof_begin(const char* name, double work, int unit, cudaStream_t st, void** cookie) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int kind = -1;
  for (size_t i = 0; i < g_kinds.size(); ++i)
    if (g_kinds[i].name == name) { kind = (int)i; break; }
  if (kind < 0) {
    g_kinds.push_back(ProfKind{});
    kind = (int)g_kinds.size() - 1;
    g_kinds[kind].name = name;
    g_kinds[kind].unit = unit;
  }
  if (g_pending.size() >= 16384) prof_fold_locked();
  ProfPair pr;
  if (!g_free.empty()) {
    pr = g_free.back();
    g_free.pop_back();
  } else {
    cudaEventCreate(&pr.a);
    cudaEventCreate(&pr.b);
  }
  pr.kind = kind;
  g_kinds[kind].work += work;
  g_kinds[kind].launches += 1;
  cudaEventRecord(pr.a, st);
  ProfPair* heap = new ProfPair(pr);
  *cookie = heap;
}

void s4_prof_end(void* cookie, cudaStream_t st) {
  ProfPair* pr = (ProfPair*)cookie;
  cudaEventRecord(pr->b, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_pending.push_back(*pr);
  delete pr;
}

extern "C" int s4_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (on) {
    prof_fold_locked();
    g_kinds.clear();
  }
  g_prof_on.store(on ? 1 : 0, std::memory_order_relaxed);
  return S4_OK;
}

extern "C" int s4_prof_num_kinds() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  prof_fold_locked();
  return (int)g_kinds.size();
}

extern "C" int s4_prof_get(int idx, char* name, int name_len, int* unit, double* ms, double* work,
                           long long* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  prof_fold_locked();
  if (idx < 0 || idx >= (int)g_kinds.size()) {
    s4_set_error("prof_get: index %d out of range", idx);
    return S4_ERR_ARG;
  }
  const ProfKind& k = g_kinds[idx];
  if (name && name_len > 0) {
    strncpy(name, k.name.c_str(), (size_t)name_len - 1);
    name[name_len - 1] = 0;
  }
  if (unit) *unit = k.unit;
  if (ms) *ms = k.ms;
  if (work) *work = k.work;
  if (launches) *launches = k.launches;
  return S4_OK;
}
