// Standalone probe (not part of the library): what does one small tcgen05.mma kind::tf32 (M = 128, K = 8) cost when the
// MMAs form a DEPENDENT chain on one accumulator, and when they are dealt round-robin to several accumulators?
// k_conv_tc and k_wgrad_tc issue 8 - 12 such MMAs per 16 KB of gathered operand; at N = 32 the math is 16 cycles per MMA.
//   A: tensor memory (as in k_conv_tc) or shared memory (K-major);  B: shared memory, K-major 128-byte swizzle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/umma_chain_probe tools/umma_chain_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {      // K-major, SWIZZLE_128B, SBO = 1024 B
  const uint32_t lo = ((addr & 0x3ffff) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

// n_acc accumulators of N columns each; MMA i goes to accumulator i % n_acc.  a_tmem: A operand from tensor memory.
__global__ void __launch_bounds__(128) k_chain(int N, int n_acc, int n_mma, int a_tmem, unsigned long long *ns_out, int b_mn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float *sm = reinterpret_cast<float *>(smem_raw + (base - smem_u32(smem_raw)));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (256 + 128) * 32; i += 128) sm[i] = 0.001f * (float)(i % 97);     // B: 256 rows x 128 B, then A: 128 rows x 128 B
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t a_col = 448;                                        // A stage (64 columns) behind the accumulators
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // b_mn: B MN-major as k_wgrad_tc has it (SWIZZLE_128B_BASE32B, 32-column groups 4096 B apart, 4-row K atoms 512 B apart);
  // the k-step advance inside the blocks below (+2 x 16 B) is not the real one (+1024 B) - timing only, the data are arbitrary
  const uint64_t db = b_mn ? ((uint64_t)((512u >> 4) | (1u << 14) | (1u << 29)) << 32) | (((base & 0x3ffff) >> 4) | ((4096u >> 4) << 16)) : make_desc(base);
  const uint64_t da = make_desc(base + 256 * 128);
  unsigned long long t0 = 0, t1 = 0;
  // the WHOLE warp 0 runs the loop converged and one elected lane issues (as the kernels do: under `if (tid == 0)` ptxas
  // wraps every MMA in a value-uniformising loop)
  if (warp == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    // blocks of 8 MMAs under ONE election with operands prepared outside (as umma_chunk_3x in conv_tc.cu does): the loop
    // overhead per MMA is then one uniform add or less.  n_acc in {1, 2, 4, 8}: MMA j of a block -> accumulator j % n_acc.
    const uint32_t d0 = tmem, d1 = tmem + (uint32_t)((1 % n_acc) * N), d2 = tmem + (uint32_t)((2 % n_acc) * N), d3 = tmem + (uint32_t)((3 % n_acc) * N),
                   d4 = tmem + (uint32_t)((4 % n_acc) * N), d5 = tmem + (uint32_t)((5 % n_acc) * N), d6 = tmem + (uint32_t)((6 % n_acc) * N),
                   d7 = tmem + (uint32_t)((7 % n_acc) * N);
    const uint32_t at = tmem + a_col;
    for (int i = 0; i < n_mma; i += 8) {
      if (a_tmem) {
        asm volatile(
            "{\n\t.reg .pred e;\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 b1, b2, b3;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "add.u32 a1, %8, 8;\n\tadd.u32 a2, %8, 16;\n\tadd.u32 a3, %8, 24;\n\t"
            "add.u64 b1, %9, 2;\n\tadd.u64 b2, %9, 4;\n\tadd.u64 b3, %9, 6;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%8], %9, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [a1], b1, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%2], [a2], b2, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%3], [a3], b3, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%4], [%8], %9, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%5], [a1], b1, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%6], [a2], b2, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%7], [a3], b3, %10, 1;\n\t}"
            ::"r"(d0), "r"(d1), "r"(d2), "r"(d3), "r"(d4), "r"(d5), "r"(d6), "r"(d7), "r"(at), "l"(db), "r"(idesc)
            : "memory");
      } else {
        asm volatile(
            "{\n\t.reg .pred e;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "add.u64 a1, %8, 2;\n\tadd.u64 a2, %8, 4;\n\tadd.u64 a3, %8, 6;\n\t"
            "add.u64 b1, %9, 2;\n\tadd.u64 b2, %9, 4;\n\tadd.u64 b3, %9, 6;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %8, %9, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], a1, b1, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%2], a2, b2, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%3], a3, b3, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%4], %8, %9, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%5], a1, b1, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%6], a2, b2, %10, 1;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::tf32 [%7], a3, b3, %10, 1;\n\t}"
            ::"r"(d0), "r"(d1), "r"(d2), "r"(d3), "r"(d4), "r"(d5), "r"(d6), "r"(d7), "l"(da), "l"(db), "r"(idesc)
            : "memory");
      }
    }
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, 20000;\n\t@p bra Dn;\n\tbra W;\n\tDn:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0 && blockIdx.x == 0) {
    unsigned long long t2;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
    ns_out[0] = t1 - t0;      // issue loop
    ns_out[1] = t2 - t0;      // until the last MMA has retired
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  unsigned long long *dns, hns[2];
  CK(cudaMalloc(&dns, 16));
  const int smem = (256 + 128) * 128 + 1024;
  CK(cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_mma = 4096;
  for (int b_mn = 0; b_mn <= 1; ++b_mn)
  for (int a_tmem = 1; a_tmem >= 0; --a_tmem)
    for (int N = 32; N <= 256; N *= 2)
      for (int n_acc = 1; n_acc * N <= 384 && n_acc <= 2; n_acc *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
          k_chain<<<148, 128, smem>>>(N, n_acc, n_mma, a_tmem, dns, b_mn);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(hns, dns, 16, cudaMemcpyDeviceToHost));
        printf("B %s, A %s, N = %3d, %d accumulator(s): issue %.1f ns per MMA, retire %.1f ns per MMA  (math floor %.1f ns at 1.9 GHz)\n", b_mn ? "MN-major" : "K-major", a_tmem ? "tmem" : "smem", N, n_acc,
               (double)hns[0] / n_mma, (double)hns[1] / n_mma, 128.0 * N / 256.0 / 1.9);
      }
  return 0;
}
