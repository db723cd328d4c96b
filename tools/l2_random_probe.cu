// Standalone probe: rate of random 16-byte loads from a table of a given size (is an 8-16 MB hash table served from
// the L2 at a higher rate than a 128-512 MB one?), with and without a concurrent streaming store of 180 bytes per
// "vertex" (what k_vertices writes).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/l2_random_probe tools/l2_random_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned)k;
}

__global__ void __launch_bounds__(256) k_probe(const int4 *__restrict__ table, unsigned mask, long long n_items, int probes, int *out,
                                                long long *stream_out, int do_stream) {
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  int acc = 0;
  for (long long i = tid; i < n_items; i += stride) {
    int4 e[4];
    for (int p0 = 0; p0 < probes; p0 += 4) {
#pragma unroll
      for (int t = 0; t < 4; ++t) e[t] = table[hash32((unsigned long long)i * 16 + p0 + t) & mask];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc += e[t].w;
    }
    if (do_stream) {
#pragma unroll
      for (int t = 0; t < 15; ++t) stream_out[(long long)t * n_items + i] = acc + t;   // 120 bytes per item, coalesced rows
    }
  }
  if (acc == 0x7fffffff) *out = acc;
}

int main() {
  int *dout; CK(cudaMalloc(&dout, 4));
  const long long n_items = 1600000;   // vertices of 16 scans at level 0
  long long *dstream; CK(cudaMalloc(&dstream, n_items * 15 * 8));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (long long mb = 8; mb <= 512; mb *= 4) {
    const long long entries = mb * 1024 * 1024 / 16;
    int4 *table; CK(cudaMalloc(&table, entries * 16)); CK(cudaMemset(table, 0, entries * 16));
    for (int do_stream = 0; do_stream <= 1; ++do_stream) {
      k_probe<<<1184, 256>>>(table, (unsigned)(entries - 1), n_items, 16, dout, dstream, do_stream);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a));
      k_probe<<<1184, 256>>>(table, (unsigned)(entries - 1), n_items, 16, dout, dstream, do_stream);
      CK(cudaEventRecord(b));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      printf("table %4lld MB, streaming stores %d: %.1f us for %.1f M probes = %.1f G probes/s\n", mb, do_stream, ms * 1e3, n_items * 16 / 1e6,
             n_items * 16 / (ms * 1e-3) / 1e9);
    }
    CK(cudaFree(table));
  }
  return 0;
}
