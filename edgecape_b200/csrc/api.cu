// Library-level entry points: version, per-thread error string, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace ec {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int g_pdl = -1;   // -1: read EDGECAPE_PDL on first use
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("EDGECAPE_PDL");
    g_pdl = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return g_pdl != 0;
}

int ensure_dynamic_smem_impl(const void* func, int bytes) {
  static std::map<std::pair<const void*, int>, int> done;   // (kernel, device) -> bytes already granted
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(func, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return 0;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[key] = bytes;
  return (int)e;
}

unsigned long long* overflow_counters() {
  constexpr int MAX_DEV = 64;
  static unsigned long long* ptrs[MAX_DEV] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) {
    set_error("overflow counters: no current CUDA device");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (!ptrs[dev]) {
    unsigned long long* p = nullptr;
    cudaError_t e = cudaMalloc(&p, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(p, 0, 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
      set_error("overflow counters: allocation failed on device %d: %s (the first F16F8 producer call on a device must "
                "be outside a stream capture)", dev, cudaGetErrorString(e));
      cudaGetLastError();
      return nullptr;
    }
    ptrs[dev] = p;
  }
  return ptrs[dev];
}

}  // namespace ec

extern "C" int ec_overflow_count(unsigned long long* out2, int reset) {
  EC_REQUIRE(out2, "ec_overflow_count: null output");
  unsigned long long* p = ec::overflow_counters();
  if (!p) return EC_ERR_CUDA;
  EC_CUDA(cudaDeviceSynchronize());
  EC_CUDA(cudaMemcpy(out2, p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) EC_CUDA(cudaMemset(p, 0, 2 * sizeof(unsigned long long)));
  return EC_OK;
}

extern "C" int ec_version(void) { return 100; }
extern "C" const char* ec_last_error_string(void) { return ec::g_err; }
extern "C" long long ec_launch_count(void) { return ec::g_launches.load(std::memory_order_relaxed); }
extern "C" int ec_set_pdl(int on) {
  ec::g_pdl = on ? 1 : 0;
  return EC_OK;
}
