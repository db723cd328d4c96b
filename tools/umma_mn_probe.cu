// Standalone probe (not part of the library): tcgen05.mma kind::tf32 with BOTH operands MN-major in shared memory
// (128-byte swizzle) - the operand form a weight-gradient GEMM needs, because its reduction index (the lattice vertex)
// is the ROW index of both row-major operand arrays:
//     D[k, m] = sum_v A[v, k] * G[v, m]        A: gathered splat rows (v, k), G: output gradients (v, m)
// Checks which of the two descriptor strides (LBO / SBO) addresses the next 32-element MN group and which the next
// 8-row K group, by running an exactly representable problem against the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/umma_mn_probe tools/umma_mn_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int V = 32;      // reduction length (vertices) = 4 MMA K-steps
constexpr int MK = 128;    // M of the MMA (conv-k index)
constexpr int NN = 64;     // N of the MMA (output channels)

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2u) {
  const uint32_t lo = ((addr & 0x3ffff) >> 4) | ((lbo_bytes >> 4) << 16);
  const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (layout << 29);      // version 1; layout 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(128) k_probe(const float *A, const float *G, float *D, int variant, int reps, unsigned long long *ns) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_base = base, g_base = base + (MK / 32) * V * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  // MN-major tiles: group g = 32 consecutive MN elements; inside a group row v (reduction index) is 128 bytes, its 16-byte
  // units XOR-swizzled in 32-byte chunks with (v & 3): the SWIZZLE_128B_BASE32B layout, the only one tf32 MN-major operands support
  if (variant == 2) {
    // sanity variant: the K-major layout the production kernel uses (row = MN index, 32 reduction elements = 128 bytes per row)
    for (int i = tid; i < V * MK; i += 128) {
      const int v = i / MK, k = i % MK;
      const uint32_t off = (uint32_t)k * 128 + ((((uint32_t)v >> 2) ^ ((uint32_t)k & 7u)) << 4) + ((uint32_t)v & 3u) * 4;
      *reinterpret_cast<float *>(smem + off) = A[i];
    }
    for (int i = tid; i < V * NN; i += 128) {
      const int v = i / NN, m = i % NN;
      const uint32_t off = (uint32_t)m * 128 + ((((uint32_t)v >> 2) ^ ((uint32_t)m & 7u)) << 4) + ((uint32_t)v & 3u) * 4;
      *reinterpret_cast<float *>(smem + (g_base - base) + off) = G[i];
    }
  } else {
  for (int i = tid; i < V * MK; i += 128) {
    const int v = i / MK, k = i % MK, g = k / 32, e = k % 32;
    const uint32_t off = (uint32_t)g * V * 128 + (uint32_t)v * 128 + ((((uint32_t)e >> 3) ^ ((uint32_t)v & 3u)) << 5) + ((uint32_t)e & 7u) * 4;
    *reinterpret_cast<float *>(smem + off) = A[i];
  }
  for (int i = tid; i < V * NN; i += 128) {
    const int v = i / NN, m = i % NN, g = m / 32, e = m % 32;
    const uint32_t off = (uint32_t)g * V * 128 + (uint32_t)v * 128 + ((((uint32_t)e >> 3) ^ ((uint32_t)v & 3u)) << 5) + ((uint32_t)e & 7u) * 4;
    *reinterpret_cast<float *>(smem + (g_base - base) + off) = G[i];
  }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  unsigned long long t0 = 0;
  if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (tid == 0) for (int rep = 0; rep < reps; ++rep) {
    // kind::tf32, fp32 accumulate, A and B MN-major (bits 15, 16), N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(MK >> 4) << 24);
    const uint32_t group_stride = V * 128, kstep_stride = 8 * 128, katom_stride = 4 * 128;
    for (int s = 0; s < V / 8 && variant == 2; ++s) {
      const uint32_t idk = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(MK >> 4) << 24);
      const uint64_t da = make_desc(a_base + s * 32, 16, 1024), dg = make_desc(g_base + s * 32, 16, 1024);
      const uint32_t acc = s > 0 || rep > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
                   "l"(dg), "r"(idk), "r"(acc)
                   : "memory");
    }
    for (int s = 0; s < V / 8 && variant != 2; ++s) {
      const uint32_t lbo = variant == 0 ? group_stride : katom_stride, sbo = variant == 0 ? katom_stride : group_stride;
      const uint64_t da = make_desc(a_base + s * kstep_stride, lbo, sbo, 1u), dg = make_desc(g_base + s * kstep_stride, lbo, sbo, 1u);
      const uint32_t acc = s > 0 || rep > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
                   "l"(dg), "r"(idesc), "r"(acc)
                   : "memory");
    }
    if (rep == reps - 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, 20000;\n\t@p bra Dn;\n\tbra W;\n\tDn:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); *ns = t1 - t0; }
  for (int cb = 0; cb < NN; cb += 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + cb)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) D[(size_t)tid * NN + cb + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
  std::vector<float> A(V * MK), G(V * NN), Dref(MK * NN, 0.f), D(MK * NN);
  srand(1);
  for (auto &x : A) x = (float)(rand() % 17 - 8) * 0.25f;
  for (auto &x : G) x = (float)(rand() % 13 - 6) * 0.5f;
  for (int v = 0; v < V; ++v)
    for (int k = 0; k < MK; ++k)
      for (int m = 0; m < NN; ++m) Dref[k * NN + m] += A[v * MK + k] * G[v * NN + m];
  float *dA, *dG, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dG, G.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
  const int smem = (MK / 32 + NN / 32) * V * 128 + 1024;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int variant = 0; variant < 3; ++variant) {
    CK(cudaMemset(dD, 0, D.size() * 4));
    unsigned long long *dns; CK(cudaMalloc(&dns, 8));
    k_probe<<<1, 128, smem>>>(dA, dG, dD, variant, 1, dns);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - Dref[i]); if (d > maxerr) maxerr = d; bad += d != 0; }
    printf("variant %d (LBO = %s stride): max abs err %.4f, %d of %zu wrong; D[0][0..3] = %.2f %.2f %.2f %.2f  ref %.2f %.2f %.2f %.2f\n", variant,
           variant == 0 ? "MN-group" : variant == 1 ? "K-group" : "(K-major sanity)", maxerr, bad, D.size(), D[0], D[1], D[2], D[3], Dref[0], Dref[1], Dref[2], Dref[3]);
  }
  for (int variant = 0; variant < 3; variant += 2) {
    unsigned long long *dns, hns; CK(cudaMalloc(&dns, 8));
    const int reps = 2000;
    k_probe<<<1, 128, smem>>>(dA, dG, dD, variant, reps, dns);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&hns, dns, 8, cudaMemcpyDeviceToHost));
    printf("%s operands: %d MMAs (M=128, N=%d, K=8) in %llu ns = %.1f ns per MMA\n", variant == 0 ? "MN-major" : "K-major", reps * V / 8, NN, hns, (double)hns / (reps * V / 8));
  }
  return 0;
}
