// capi.cu -- error reporting, launch accounting and version of libpysco_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

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

// SM count of the current device (148 on a B200), queried once per device; grids of grid-stride / persistent kernels
// are sized in multiples of it.
int num_sms() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsB200;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMsB200;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace psc

extern "C" {
const char *psc_last_error(void) { return psc::g_err; }
int psc_version(void) { return 100; }
int64_t psc_launch_count(void) { return (int64_t)psc::g_launches.load(std::memory_order_relaxed); }
}
