// Standalone probe: rate of "add a row of floats to a random row of a matrix in global memory" - the access pattern of the
// level-0 splat (144-byte rows) and of the split-K epilogue of k_conv_tc (128-byte pieces of 1 KB rows) - done
//   (a) with per-lane vector atomics  red.global.add.v4.f32  (REDG.E.ADD.F32x4: what the kernels use), and
//   (b) with the TMA engine:          cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32  (one bulk op per row)
// for several row sizes.  Is the vector-atomic ceiling (~150 G red.v4/s) a limit of the L2's atomic units or of the SM's
// LSU path, i.e. does the bulk path get past it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/bulk_reduce_probe tools/bulk_reduce_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned)k;
}

// (a) one "row add" = ROWB/16 red.v4; thread = (row of the round, 16-byte piece) like k_scatter's item loop
template <int ROWB>
__global__ void __launch_bounds__(256) k_red(float *dst, long long rows, long long pitch_f, long long n_ops) {
  constexpr int P = ROWB / 16;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n_ops * P; i += stride) {
    const long long op = i / P;
    const int piece = (int)(i - op * P);
    const long long row = hash32(op) % rows;
    atomicAdd(reinterpret_cast<float4 *>(dst + row * pitch_f) + piece, make_float4(1.f, 1.f, 1.f, 1.f));
  }
}

// (b) every thread owns one ROWB-byte row in shared memory and adds it to random rows with bulk reduce operations
template <int ROWB>
__global__ void __launch_bounds__(128) k_bulk(float *dst, long long rows, long long pitch_f, long long n_ops, int group) {
  extern __shared__ __align__(128) unsigned char smem[];
  float *mine = reinterpret_cast<float *>(smem + (size_t)threadIdx.x * ROWB);
  for (int i = 0; i < ROWB / 4; ++i) mine[i] = 1.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const unsigned src = (unsigned)__cvta_generic_to_shared(mine);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  int pending = 0;
  for (long long op = tid; op < n_ops; op += stride) {
    const long long row = hash32(op) % rows;
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + row * pitch_f), "r"(src), "r"(ROWB) : "memory");
    if (++pending == group) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");    // (what a kernel that re-fills the source would do)
      pending = 0;
    }
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int ROWB>
void run(float *dst, long long rows, long long pitch_f, long long n_ops, double expect_per_row) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(a));
    k_red<ROWB><<<148 * 8, 256>>>(dst, rows, pitch_f, n_ops);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("row %4d B  red.v4 : %8.1f us  %6.1f M rows  %6.2f G rows/s  %6.1f G red.v4/s  %6.2f TB/s\n", ROWB, ms * 1e3, n_ops / 1e6, n_ops / (ms * 1e-3) / 1e9,
         n_ops * (ROWB / 16) / (ms * 1e-3) / 1e9, n_ops * (double)ROWB / (ms * 1e-3) / 1e12);
  for (int group = 1; group <= 4; group *= 4) {
    CK(cudaFuncSetAttribute(k_bulk<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * ROWB));
    for (int ctas = 1; ctas <= 4; ctas *= 2) {
      if ((size_t)ctas * 128 * ROWB > 200 * 1024) continue;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(a));
        k_bulk<ROWB><<<148 * ctas, 128, 128 * ROWB>>>(dst, rows, pitch_f, n_ops, group);
        CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaEventElapsedTime(&ms, a, b));
      printf("row %4d B  bulk   : %8.1f us  (%d CTAs/SM x 128 thr, commit every %d)  %6.2f G rows/s  %6.2f TB/s  %5.1f ns per op per SM\n", ROWB, ms * 1e3, ctas, group,
             n_ops / (ms * 1e-3) / 1e9, n_ops * (double)ROWB / (ms * 1e-3) / 1e12, ms * 1e6 / (n_ops / 148.0));
    }
  }
  // spot check: total added = n_ops * 2 * (1 + 2 * 3 variants...) is awkward; check one row's value is an integer > 0 instead
  float h[4]; CK(cudaMemcpy(h, dst, 16, cudaMemcpyDeviceToHost));
  printf("            (row 0 holds %.0f %.0f %.0f %.0f)\n", h[0], h[1], h[2], h[3]);
  (void)expect_per_row;
}

// red.v4 only: row size / pitch / destination size sweep (is the level-0 splat's rate - 144-byte rows at a 144-byte pitch,
// 232 MB for 16 scans, 14.5 MB for one - set by DRAM residency, by the unaligned rows, or by the SM's atomic issue rate?)
template <int ROWB>
void run_red(float *dst, long long rows, long long pitch_f, long long n_ops, const char *what) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float ms;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    k_red<ROWB><<<148 * 8, 256>>>(dst, rows, pitch_f, n_ops);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("red.v4 row %4d B pitch %4lld B x %8lld rows (%6.1f MB) %-28s: %8.1f us  %6.2f G rows/s  %6.1f G red.v4/s\n", ROWB, pitch_f * 4, rows,
         rows * pitch_f * 4 / 1e6, what, ms * 1e3, n_ops / (ms * 1e-3) / 1e9, n_ops * (ROWB / 16) / (ms * 1e-3) / 1e9);
}


// scan-local variant: the ops walk the batch scan by scan (as the splat does: points of scan b only touch rows of scan b),
// after a zero-fill of the WHOLE destination (what k_zero does) or of nothing
template <int ROWB>
__global__ void __launch_bounds__(256) k_red_local(float *dst, long long rows_per_scan, long long pitch_f, long long n_ops, long long ops_per_scan) {
  constexpr int P = ROWB / 16;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n_ops * P; i += stride) {
    const long long op = i / P;
    const int piece = (int)(i - op * P);
    const long long scan = op / ops_per_scan;
    const long long row = scan * rows_per_scan + hash32(op) % rows_per_scan;
    atomicAdd(reinterpret_cast<float4 *>(dst + row * pitch_f) + piece, make_float4(1.f, 1.f, 1.f, 1.f));
  }
}
template <int ROWB>
void run_local(float *dst, int scans, long long rows_per_scan, long long pitch_f, long long ops_per_scan, int zero_first, int ctas) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float ms;
  const long long n_ops = ops_per_scan * scans;
  for (int rep = 0; rep < 3; ++rep) {
    if (zero_first) CK(cudaMemsetAsync(dst, 0, scans * rows_per_scan * pitch_f * 4));
    CK(cudaEventRecord(a));
    k_red_local<ROWB><<<148 * ctas, 256>>>(dst, rows_per_scan, pitch_f, n_ops, ops_per_scan);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
  }
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("scan-local red.v4 row %d B, %d scans x %lld rows, zero-fill first %d, %d CTAs/SM: %8.1f us  %6.1f G red.v4/s\n", ROWB, scans, rows_per_scan, zero_first, ctas, ms * 1e3,
         n_ops * (ROWB / 16) / (ms * 1e-3) / 1e9);
}

int main(int argc, char **argv) {
  const long long rows = 1600000;              // vertices of 16 scans at level 0
  const long long pitch_f = 256;               // 1 KB pitch: every row size below fits
  float *dst; CK(cudaMalloc(&dst, rows * pitch_f * 4)); CK(cudaMemset(dst, 0, rows * pitch_f * 4));
  const long long n_ops = 8400000;             // (point, remainder) contributions of 16 scans
  if (argc > 1 && argv[1][0] == 'l') {         // "local": scan-local sweep
    for (int z = 0; z <= 1; ++z)
      for (int ctas = 2; ctas <= 8; ctas *= 2) run_local<144>(dst, 16, 100000, 36, 525000, z, ctas);
    run_local<144>(dst, 16, 100000, 36, 525000, 1, 8);
    return 0;
  }
  if (argc > 1) {                              // "red": the red.v4 sweep only
    run_red<144>(dst, 1600000, 36, n_ops, "16 scans, level-0 S");
    run_red<144>(dst, 800000, 36, n_ops, "8 scans");
    run_red<144>(dst, 400000, 36, n_ops, "4 scans");
    run_red<144>(dst, 100000, 36, n_ops, "1 scan");
    run_red<144>(dst, 1600000, 40, n_ops, "16 scans, rows padded to 160");
    run_red<144>(dst, 100000, 40, n_ops, "1 scan, rows padded to 160");
    run_red<128>(dst, 1600000, 32, n_ops, "16 scans, 128-byte rows");
    run_red<128>(dst, 100000, 32, n_ops, "1 scan, 128-byte rows");
    run_red<160>(dst, 100000, 40, n_ops, "1 scan, 160-byte rows");
    run_red<272>(dst, 1000000, 68, n_ops / 2, "level-1 S, 16 scans");
    run_red<272>(dst, 62500, 68, n_ops / 2, "level-1 S, 1 scan");
    return 0;
  }
  run<128>(dst, rows, pitch_f, n_ops, 0);
  run<144>(dst, rows, pitch_f, n_ops, 0);
  run<256>(dst, rows, pitch_f, n_ops / 2, 0);
  run<512>(dst, rows, pitch_f, n_ops / 4, 0);
  run<1024>(dst, rows, pitch_f, n_ops / 8, 0);
  // the same with a small destination (67k rows x 1 KB = the level-3 convolution output of 16 scans: L2-resident)
  printf("--- destination 67104 rows (69 MB, L2-resident)\n");
  run<128>(dst, 67104, pitch_f, n_ops, 0);
  run<1024>(dst, 67104, pitch_f, n_ops / 8, 0);
  return 0;
}
