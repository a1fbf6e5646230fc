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
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else
      n = 148;  // B200
  }
  return n;
}

extern "C" const char* dfmir_last_error(void) { return g_err; }
extern "C" int dfmir_abi_version(void) { return DFMIR_ABI_VERSION; }
extern "C" long long dfmir_launch_count(void) { return g_launches.load(); }
extern "C" void dfmir_launch_count_reset(void) { g_launches.store(0); }
