// capi.cu -- error reporting, launch accounting and version of libpysco_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace psc {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace psc

extern "C" {
const char *psc_last_error(void) { return psc::g_err; }
int psc_version(void) { return 100; }
int64_t psc_launch_count(void) { return (int64_t)psc::g_launches.load(std::memory_order_relaxed); }
}
