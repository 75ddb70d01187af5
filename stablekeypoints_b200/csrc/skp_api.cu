// Library-level entry points of libskp_b200: version, per-thread error string, launch counter.
#include "skp_common.cuh"
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>

namespace skp {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  // measured on B200 inside the 3-stream step graph: no gain (36.04 vs 36.07 images/s), so the attribute is opt-in
  static const bool on = getenv("SKP_PDL") != nullptr && atoi(getenv("SKP_PDL")) != 0;
  return on;
}

}  // namespace skp

extern "C" int skp_version(void) { return 100; }
extern "C" const char* skp_last_error(void) { return skp::g_err; }
extern "C" int64_t skp_launch_count(void) { return skp::g_launches.load(std::memory_order_relaxed); }
