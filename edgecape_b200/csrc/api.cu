// Library-level entry points: version, per-thread error string, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

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

}  // namespace ec

extern "C" int ec_version(void) { return 100; }
extern "C" const char* ec_last_error_string(void) { return ec::g_err; }
extern "C" long long ec_launch_count(void) { return ec::g_launches.load(std::memory_order_relaxed); }
extern "C" int ec_set_pdl(int on) {
  ec::g_pdl = on ? 1 : 0;
  return EC_OK;
}
