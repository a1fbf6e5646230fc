// Library-wide plumbing of the C ABI: error strings, device queries, launch accounting.
#include "common.cuh"
#include "dfmir_b200.h"
#include <atomic>
#include <stdarg.h>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void dfmir_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void dfmir_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int dfmir_num_sms() {
  // cached per device: one process may drive a model on a device other than the first one it touched
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    int v = 0;
    n = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;  // B200
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

extern "C" const char* dfmir_last_error(void) { return g_err; }
extern "C" int dfmir_abi_version(void) { return DFMIR_ABI_VERSION; }
extern "C" long long dfmir_launch_count(void) { return g_launches.load(); }
extern "C" void dfmir_launch_count_reset(void) { g_launches.store(0); }
