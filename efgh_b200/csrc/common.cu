// Error reporting and device queries of the efgh_b200 C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace efgh {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

}  // namespace efgh

extern "C" const char *efgh_last_error(void) { return efgh::g_error; }
extern "C" int efgh_version(void) { return 100; }
extern "C" int efgh_device_sm_count(void) {
  int dev = 0, v = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    efgh::set_error("no CUDA device: %s", cudaGetErrorString(e));
    return EFGH_ECUDA;
  }
  return v;
}
