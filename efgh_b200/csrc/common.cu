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

static unsigned long long g_launches = 0;
void note_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

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
extern "C" int efgh_version(void) { return 200; }
namespace efgh { unsigned long long launches(); }
extern "C" int64_t efgh_launch_count(void) { return (int64_t)efgh::launches(); }
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

// Strided host <-> device copy of a (rows, cols) float matrix (cudaMemcpy2DAsync): packs one scan's (3,n) cloud /
// (C,n) feature matrix into its column range of a batch's (3, B*n) / (C, B*n) device matrix with one DMA
// descriptor per scan instead of a staging buffer + copy kernel.  `kind`: 1 = host -> device, 2 = device -> host.
extern "C" int efgh_copy_matrix_async(void *dst, int64_t dst_ld, const void *src, int64_t src_ld, int64_t rows,
                                      int64_t cols, int kind, void *stream) {
  EFGH_REQUIRE(rows >= 0 && cols >= 0 && dst_ld >= cols && src_ld >= cols, "efgh_copy_matrix_async: bad sizes");
  EFGH_REQUIRE(kind == 1 || kind == 2, "efgh_copy_matrix_async: kind must be 1 (H2D) or 2 (D2H)");
  if (rows == 0 || cols == 0) return EFGH_OK;
  EFGH_REQUIRE(dst && src, "efgh_copy_matrix_async: null pointer");
  EFGH_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)dst_ld * 4, src, (size_t)src_ld * 4, (size_t)cols * 4, (size_t)rows,
                                    kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                    static_cast<cudaStream_t>(stream)));
  return EFGH_OK;
}
