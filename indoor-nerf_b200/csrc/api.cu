// Library-wide plumbing of the C ABI: error text, launch accounting, device queries.
#include <stdarg.h>

#include <atomic>

#include "pn_common.cuh"

namespace pn {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PN_ECUDA;
  }
  return 0;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace pn

extern "C" int pn_abi_version(void) { return PN_ABI_VERSION; }
extern "C" const char *pn_last_error(void) { return pn::g_err; }
extern "C" int64_t pn_launch_count(void) { return pn::g_launches.load(std::memory_order_relaxed); }
