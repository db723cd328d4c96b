// Shared helpers for the efgh_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "efgh_b200.h"

namespace efgh {

void set_error(const char *fmt, ...);
int sm_count();
void note_launch();   // bumps the process-wide kernel-launch counter (efgh_launch_count)

// Turns a CUDA error into the C-ABI status code + message.
#define EFGH_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      efgh::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return EFGH_ECUDA;                                                                        \
    }                                                                                           \
  } while (0)

#define EFGH_REQUIRE(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      efgh::set_error(__VA_ARGS__);   \
      return EFGH_EINVAL;             \
    }                                 \
  } while (0)

#define EFGH_LAUNCH_CHECK()              \
  do {                                   \
    efgh::note_launch();                 \
    EFGH_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

// Grid for a grid-stride kernel over `items` work items: enough CTAs to cover the items once, capped at
// a multiple of the SM count (148 on B200) so that a capacity-sized launch does not flood the machine
// with empty CTAs when the device-side count is much smaller than the capacity.
inline int grid_for(int64_t items, int threads, int ctas_per_sm) {
  int64_t need = (items + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ int warp_reduce_min(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_reduce_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename IdxT>
__device__ __forceinline__ int load_idx(const void *base, int64_t i) {
  return (int)__ldg(reinterpret_cast<const IdxT *>(base) + i);
}

}  // namespace efgh
