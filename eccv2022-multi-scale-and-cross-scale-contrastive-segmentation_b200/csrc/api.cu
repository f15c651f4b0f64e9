// api.cu -- library-level entry points: version, error string, device probe.
#include "common.cuh"
#include <string.h>

namespace mscs {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
}  // namespace mscs

extern "C" const char* mscs_version(void) { return "mscs 0.1.0 (sm_100a, tcgen05/TMA)"; }
extern "C" const char* mscs_last_error(void) { return mscs::get_error(); }
extern "C" int mscs_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10;
}
