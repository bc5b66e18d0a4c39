// libacx: version / error plumbing of the C ABI declared in include/acx.h.
#include <stdarg.h>
#include <string.h>

#include "../../include/acx.h"
#include "common.cuh"

namespace acx {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
}  // namespace acx

extern "C" {

int acx_version(void) { return ACX_VERSION; }

const char* acx_last_error(void) { return acx::last_error(); }

int acx_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    acx::set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return 0;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    acx::set_error("libacx is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return 0;
  }
  return 1;
}

}  // extern "C"
