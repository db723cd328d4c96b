// Standalone probe (not part of the library): semantics and throughput of the Blackwell TMA row gather
//   cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4
// as a replacement for the LDGSTS gather of k_conv_tc: 4 rows x 128 bytes per instruction, written to shared
// memory in the 128-byte-swizzled K-major layout a tcgen05 A operand wants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_gather_probe tools/tma_gather_probe.cu
// Part 1 checks the landed layout (row order, swizzle, out-of-range columns, row 0) for box {32, 1} / {32, 4} maps;
// part 2 measures gathered bytes/s for a persistent grid with a ring of stages.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
               "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
               : "memory");
}

// ---- part 1: one 128-row x 32-float tile, dumped raw
__global__ void k_layout(const __grid_constant__ CUtensorMap map, const int *rows, int col, float *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = lane; i < 128 * 32; i += 32) reinterpret_cast<float *>(smem + (base - smem_u32(smem)))[i] = -7.f;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (lane == 0) mbar_expect(smem_u32(&bar), 128 * 128);
  __syncwarp();
  gather4(base + lane * 512, &map, col, rows[4 * lane], rows[4 * lane + 1], rows[4 * lane + 2], rows[4 * lane + 3], smem_u32(&bar));
  mbar_wait(smem_u32(&bar), 0);
  for (int i = lane; i < 128 * 32; i += 32) out[i] = reinterpret_cast<float *>(smem + (base - smem_u32(smem)))[i];
}

// ---- part 2: throughput.  One producer warp per CTA keeps `stages` 16 KB tiles in flight; a consumer warp just
// releases them (touching one word), so the number is the TMA gather rate, not a full pipeline.
__global__ void __launch_bounds__(64) k_rate(const __grid_constant__ CUtensorMap map, const int *rows, int n_rows_total, int chunks_per_cta,
                                              int cols, int stages, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[8], empty[8];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long acc = 0;
  if (warp == 0) {
    int s = 0, ph = 0;
    for (int c = 0; c < chunks_per_cta; ++c) {
      mbar_wait(smem_u32(&empty[s]), ph ^ 1);
      if (lane == 0) mbar_expect(smem_u32(&full[s]), 128 * 128);
      __syncwarp();
      const int tile = (blockIdx.x * chunks_per_cta + c) % (n_rows_total / 128);
      const int4 r = *reinterpret_cast<const int4 *>(rows + tile * 128 + 4 * lane);
      gather4(base + s * 16384 + lane * 512, &map, (c * 32) % cols, r.x, r.y, r.z, r.w, smem_u32(&full[s]));
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else {
    int s = 0, ph = 0;
    for (int c = 0; c < chunks_per_cta; ++c) {
      mbar_wait(smem_u32(&full[s]), ph);
      acc += *reinterpret_cast<volatile uint32_t *>(smem + (base - smem_u32(smem)) + s * 16384 + lane * 4);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  if (acc == 0x1234567ull) *sink = acc;
}

int main() {
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int R = 200000, C = 36, ld = 36;
  std::vector<float> hx((size_t)R * ld);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) hx[(size_t)r * ld + c] = (float)(r * 100 + c);
  float *dx; CK(cudaMalloc(&dx, hx.size() * 4)); CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
  std::vector<int> hrows(R);
  srand(1);
  for (int i = 0; i < R; ++i) hrows[i] = rand() % R;
  hrows[0] = 0; hrows[1] = R - 1; hrows[2] = 5; hrows[3] = 5;
  int *drows; CK(cudaMalloc(&drows, R * 4)); CK(cudaMemcpy(drows, hrows.data(), R * 4, cudaMemcpyHostToDevice));
  float *dout; CK(cudaMalloc(&dout, 128 * 32 * 4));
  unsigned long long *dsink; CK(cudaMalloc(&dsink, 8));
  for (int box_rows = 1; box_rows <= 4; box_rows += 3) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {32,%d}: encode rc=%d\n", box_rows, (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    for (int col = 0; col <= 16; col += 16) {
      CK(cudaMemset(dout, 0, 128 * 32 * 4));
      k_layout<<<1, 32, 128 * 128 + 1024>>>(map, drows, col, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  col %d: kernel failed: %s\n", col, cudaGetErrorString(e)); return 2; }
      std::vector<float> ho(128 * 32);
      CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
      // expected: row r of the tile at r*128 bytes, 16-byte unit u at ((u ^ (r & 7)) << 4)
      int bad = 0, first_bad = -1;
      for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 32; ++k) {
          const int u = k >> 2, e4 = k & 3;
          const float got = ho[r * 32 + ((u ^ (r & 7)) << 2) + e4];
          const int c = col + k;
          const float want = c < C ? (float)(hrows[r] * 100 + c) : 0.f;
          if (got != want) { if (first_bad < 0) first_bad = r * 32 + k; ++bad; }
        }
      printf("  col %2d: %d mismatches vs the swizzled-tile model", col, bad);
      if (bad) printf(" (first at row %d k %d: got %.0f)", first_bad / 32, first_bad % 32, ho[(first_bad / 32) * 32 + ((((first_bad % 32) >> 2) ^ ((first_bad / 32) & 7)) << 2) + (first_bad & 3)]);
      printf("; tile[0][0..3] = %.0f %.0f %.0f %.0f, tile row1 unit0 = %.0f\n", ho[0], ho[1], ho[2], ho[3], ho[32 + (1 << 2)]);
    }
    // throughput
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int stages = 2; stages <= 8; stages *= 2) {
      const int chunks = 4000;
      CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024));
      cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      k_rate<<<sms, 64, stages * 16384 + 1024>>>(map, drows, R - R % 128, 200, C, stages, dsink);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a));
      k_rate<<<sms, 64, stages * 16384 + 1024>>>(map, drows, R - R % 128, chunks, C, stages, dsink);
      CK(cudaEventRecord(b));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  rate kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      const double bytes = (double)sms * chunks * 16384;
      printf("  stages %d: %.1f us, %.2f TB/s gathered (16 KB tiles, 128 random rows each), %.0f ns per tile per SM\n", stages, ms * 1e3,
             bytes / (ms * 1e-3) / 1e12, ms * 1e6 / chunks);
    }
  }
  return 0;
}
