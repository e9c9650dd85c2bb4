// Library-level entry points: error string, device probe.
#include <stdarg.h>

#include "common.cuh"

namespace mnrf {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mnrf

extern "C" int mnrf_abi_version(void) { return MNRF_ABI_VERSION; }
extern "C" const char* mnrf_last_error(void) { return mnrf::g_err; }

extern "C" int mnrf_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { mnrf::set_error("no CUDA device"); return 0; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { mnrf::set_error("no device properties"); return 0; }
  if (prop.major != 10) {
    mnrf::set_error("device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
    return 0;
  }
  return 1;
}

extern "C" int mnrf_num_sms(void) {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}
