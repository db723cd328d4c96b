// Weight gradient of the lattice convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   dWt[k, m] += sum_h A[h, k] * G[h, m],   A[h, f*C + c] = X[nbr[f,h]+1, c]   (gathered splat rows),  G = dY (masked)
//
// reference: autograd over nets/bilateralNN.py:240-244 (cuDNN wgrad of the (F,1) convolution over the materialised
// (1, C, F, H) gathered tensor).  The reduction runs over the lattice VERTICES, i.e. over the row index of both
// row-major operand arrays, so both tensor-core operands are "MN-major": a gathered row piece (32 consecutive k of one
// vertex, 128 bytes) is already one row of the canonical MN-major shared-memory tile, and so is a 128-byte piece of a
// dY row.  tf32 MN-major operands must use the SWIZZLE_128B_BASE32B layout (32-byte chunks XOR-ed with row & 3; layout
// and descriptor strides established by tools/umma_mn_probe.cu):
//   tile   = groups of 32 MN elements; inside a group row v (vertex) is 128 bytes; LBO = group stride, SBO = 4 rows
//   MMA    = M 128 (conv-k) x N (<= 128 output channels) x K 8 (vertices), both operands from shared memory
//
//   warps 0-7   producers: per stage of 32 vertices cp.async (LDGSTS, zero-fill) 4 x 32 conv-k of every vertex's
//               gathered rows and N floats of its dY row into the stage, then split what landed into the TF32-exact
//               part (which the tensor core obtains by truncating the raw fp32 operand) and the remainder "small",
//               written to a second tile: 3xTF32 = raw*raw + small*raw + raw*small
//   warp  8     MMA issuer (one elected lane), fp32 accumulators in TMEM, two accumulator stages
//   warps 9-12  epilogue: accumulation chains are CUT every 8 stages (256 vertices; the tensor core rounds toward zero
//               on every accumulate) and summed with round-to-nearest adds into a shared-memory running tile; at the end
//               of a work item the 128 x N tile is added to dWt in global memory (red.add, one per vertex split)
// Work item = (128-wide conv-k tile, 128-wide slice of the output channels, vertex range); the vertex count is read
// from device memory and split so that all SMs have work.
#include <stdlib.h>

#include "common.cuh"

// Timing-study switches (tools/wgrad_tc_time.py; results are WRONG with any of them set):
//   bit 0  the MMA issuer issues only the first k-step of a stage        bit 1  producers skip the "small" tiles
//   bit 2  producers skip the copies                                     bit 3  the epilogue only hands the accumulators back
static int g_wgrad_flags = 0;
extern "C" void efgh_debug_set_wgrad_flags(int flags) { g_wgrad_flags = flags; }

namespace efgh {
namespace {

constexpr int kWM = 128;                       // conv-k rows per tile (MMA M)
constexpr int kWStageV = 32;                   // vertices per stage (4 MMA K-steps)
constexpr int kWMaxStages = 6;                 // ring of stages (as many as shared memory allows: 3 at N = 128, 4 at 64, 5 at 32)
#ifndef EFGH_WGRAD_TEAMS
#define EFGH_WGRAD_TEAMS 3
#endif
constexpr int kWTeams = EFGH_WGRAD_TEAMS;      // producer teams of 4 warps; team t fills stages t, t + kWTeams, ...: that many stages are being gathered at a time
constexpr int kWCutStages = 8;                 // stages per accumulation chain (256 vertices)
constexpr int kWProducerWarps = 4 * kWTeams;
constexpr int kWMmaWarp = kWProducerWarps, kWEpiWarp0 = kWProducerWarps + 1;
constexpr int kWThreads = (kWEpiWarp0 + 4) * 32;   // 416

__device__ __forceinline__ uint32_t w_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void w_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void w_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void w_mbar_wait(uint32_t bar, uint32_t parity, uint32_t sleep_ns = 32) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (done) break;
    __nanosleep(32);
  }
}
// non-blocking phase test (same for every lane of a converged warp only if the caller makes it so: see w_slot_free)
__device__ __forceinline__ bool w_mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void w_commit(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(bar)
      : "memory");
}
// One stage (32 vertices = 4 k-steps of 1024 bytes) under ONE election, descriptors advanced inside the block:
//   A_raw x [G_raw | G_small] -> d (N' = 2 Np: main | correction),  A_small x G_raw -> d_corr (N = Np),
// then the commit that frees the stage.  tools/umma_chain_probe.cu: an MMA issued from a block like this costs 15 / 18 / 36 /
// 71 ns at N = 32 / 64 / 128 / 256, ~100 ns of the issuing warp's time from a per-MMA loop (election, predicate, descriptor
// arithmetic each time).  Measured effect here: 1 204 -> 1 187 us over the ten launches of a training step - the kernel is
// bound by its producer teams, the block is simply the cheaper form.
__device__ __forceinline__ void w_umma_stage_ss(uint32_t d, uint32_t d_corr, uint64_t desc_a_raw, uint64_t desc_a_small, uint64_t desc_g_raw,
                                                 uint32_t idesc2, uint32_t idesc, uint32_t acc_first, uint32_t bar_empty, int ksteps) {
  asm volatile(
      "{\n\t"
      ".reg .pred e, pf, p1, p2, p3;\n\t"
      ".reg .b64 a1, a2, a3, s1, s2, s3, g1, g2, g3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %7, 0;\n\t"
      "setp.gt.s32 p1, %9, 1;\n\t"
      "setp.gt.s32 p2, %9, 2;\n\t"
      "setp.gt.s32 p3, %9, 3;\n\t"
      "and.pred p1, p1, e;\n\t"
      "and.pred p2, p2, e;\n\t"
      "and.pred p3, p3, e;\n\t"
      "add.u64 a1, %2, 64;\n\t"
      "add.u64 a2, %2, 128;\n\t"
      "add.u64 a3, %2, 192;\n\t"
      "add.u64 s1, %3, 64;\n\t"
      "add.u64 s2, %3, 128;\n\t"
      "add.u64 s3, %3, 192;\n\t"
      "add.u64 g1, %4, 64;\n\t"
      "add.u64 g2, %4, 128;\n\t"
      "add.u64 g3, %4, 192;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %4, %5, pf;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], %3, %4, %6, 1;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::tf32 [%0], a1, g1, %5, 1;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::tf32 [%1], s1, g1, %6, 1;\n\t"
      "@p2 tcgen05.mma.cta_group::1.kind::tf32 [%0], a2, g2, %5, 1;\n\t"
      "@p2 tcgen05.mma.cta_group::1.kind::tf32 [%1], s2, g2, %6, 1;\n\t"
      "@p3 tcgen05.mma.cta_group::1.kind::tf32 [%0], a3, g3, %5, 1;\n\t"
      "@p3 tcgen05.mma.cta_group::1.kind::tf32 [%1], s3, g3, %6, 1;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
      "}" ::"r"(d), "r"(d_corr), "l"(desc_a_raw), "l"(desc_a_small), "l"(desc_g_raw), "r"(idesc2), "r"(idesc), "r"(acc_first), "r"(bar_empty), "r"(ksteps)
      : "memory");
}
// MN-major, SWIZZLE_128B_BASE32B (layout type 1), descriptor version 1: LBO = stride between 32-element MN groups,
// SBO = stride between 4-row K atoms (512 B with 128-byte rows)
__device__ __forceinline__ uint64_t w_desc(uint32_t addr, uint32_t group_stride) {
  const uint32_t lo = ((addr & 0x3ffff) >> 4) | ((group_stride >> 4) << 16);
  const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
  return ((uint64_t)hi << 32) | lo;
}
// kind::tf32, fp32 accumulate, A and B MN-major (bits 15, 16), M = 128
__device__ __forceinline__ uint32_t w_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kWM >> 4) << 24);
}
__device__ __forceinline__ void w_tmem_ld32(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void w_tmem_st32(uint32_t taddr, const float *v) {
  const uint32_t *r = reinterpret_cast<const uint32_t *>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

struct WgradParams {
  const float *X; int64_t ldX; int C;
  const void *nbr; int64_t nbr_ld; int F;
  int h_host; const int32_t *h_dev;
  const float *G; int64_t ldG; int M;
  float *dWt;                                 // (F*C, M) row-major, accumulated into
  float *dbias;                               // optional (M): column sums of G, accumulated into (by the producers of the kt == 0 items)
  int K, n_kt, n_np, Np;                      // K = F*C; conv-k tiles; output-channel slices of Np columns
  int ring;                                   // stages in the shared-memory ring
  int acc_stages;                             // accumulator stages in tensor memory (each 2 Np columns: main | correction)
  int dbg;
  uint32_t magic_c;
};

// work item -> (vertex split, conv-k tile, output slice); vertex ranges are whole stages
struct WItem {
  int kt, np, st_begin, st_end;
};
__device__ __forceinline__ WItem w_item(const WgradParams &p, int item, int n_stages, int vsplit) {
  const int per = p.n_kt * p.n_np;
  const int vs = item / per, r = item - vs * per;
  WItem it;
  it.np = r / p.n_kt; it.kt = r - it.np * p.n_kt;
  const int base = n_stages / vsplit, rem = n_stages % vsplit;
  it.st_begin = vs * base + min(vs, rem);
  it.st_end = it.st_begin + base + (vs < rem ? 1 : 0);
  return it;
}

template <typename IdxT>
__global__ void __launch_bounds__(kWThreads, 1) k_wgrad_tc(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t w_smem_raw[];
  const uint32_t smem_base = (w_smem_u32(w_smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = w_smem_raw + (smem_base - w_smem_u32(w_smem_raw));
  const int Np = p.Np;
  const uint32_t a_bytes = 4u * 4096u;                         // 4 groups x 32 rows x 128 B
  const uint32_t g_bytes = (uint32_t)(Np / 32) * 4096u;
  const uint32_t stage_bytes = 2u * a_bytes + 2u * g_bytes;   // [A raw | A small | G raw | G small]
  const uint32_t ring = (uint32_t)p.ring;
  const uint32_t sum_base = smem_base + ring * stage_bytes;
  const int pitch = 36;                                        // floats per row of the 32-column staging block of the epilogue
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)ring * stage_bytes + (size_t)kWM * pitch * 4);
  // barriers: full[8] empty[8] acc_full[2] acc_empty[2]
  const uint32_t bar_full = w_smem_u32(bars), bar_empty = bar_full + 64, bar_acc_full = bar_full + 128, bar_acc_empty = bar_full + 144;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // poll intervals of the three roles (timing studies override them: flags bits 8-15 / 16-23 / 24-31, units of 8 ns)
  const uint32_t ns_prod = ((uint32_t)p.dbg >> 8 & 0xffu) ? ((uint32_t)p.dbg >> 8 & 0xffu) * 8u : 32u;
  const uint32_t ns_epi = ((uint32_t)p.dbg >> 16 & 0xffu) ? ((uint32_t)p.dbg >> 16 & 0xffu) * 8u : 32u;
  const uint32_t ns_mma = ((uint32_t)p.dbg >> 24 & 0xffu) ? ((uint32_t)p.dbg >> 24 & 0xffu) * 8u : 32u;
  const int H = p.h_dev ? min(*p.h_dev, p.h_host) : p.h_host;
  const int n_stages = (H + kWStageV - 1) / kWStageV;
  const int per = p.n_kt * p.n_np;
  int vsplit = (3 * (int)gridDim.x) / per;                    // at most three full rounds of work items (no straggler round)
  vsplit = max(1, min(vsplit, (n_stages + kWCutStages - 1) / kWCutStages));
  const int n_items = n_stages > 0 ? per * vsplit : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWMaxStages; ++s) { w_mbar_init(bar_full + 8 * s, kWProducerWarps / kWTeams); w_mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { w_mbar_init(bar_acc_full + 8 * a, 1); w_mbar_init(bar_acc_empty + 8 * a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(w_smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;

  if (warp < kWProducerWarps) {
    // ===================== producers =====================
    // Two teams of 128 threads; team t fills stages t, t + 2, ... of the CTA's stage sequence (all stages of item
    // blockIdx.x, then of item blockIdx.x + gridDim.x, ...).  Warp = 32-wide MN group; a lane copies one 16-byte piece
    // of 8 vertices' row pieces (gathered rows and dY rows), waits for its OWN copies only, derives the "small" tiles
    // from what it copied and arrives - no cross-thread dependency inside a stage.
    const int team = warp / (kWProducerWarps / kWTeams);
    const int g = warp & 3;                                    // this warp's 32-wide MN group (conv-k / output-channel columns)
    const int u = lane & 7, vsub = lane >> 3;                  // 8 lanes cooperate on one 128-byte row piece: every LDGSTS instruction
    const bool has_g = g < Np / 32;                            // covers 4 rows x 128 contiguous bytes (whole cache lines)
    uint32_t doff[8];                                          // piece u of rows 4 i + vsub: 32-byte chunks XOR (row & 3)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t vr = 4u * i + (uint32_t)vsub;
      doff[i] = (uint32_t)g * 4096u + vr * 128u + (((((uint32_t)u >> 1) ^ (vr & 3u))) << 5) + ((uint32_t)u & 1u) * 16u;
    }

    // position in the flattened (item, stage) sequence
    struct Pos { int item, st; WItem it; uint32_t count; };
    auto pos_valid = [&](const Pos &q) { return q.item < n_items; };
    auto pos_step = [&](Pos &q) {                              // one stage forward
      ++q.count;
      if (++q.st >= q.it.st_end) {
        q.item += (int)gridDim.x;
        if (q.item < n_items) { q.it = w_item(p, q.item, n_stages, vsplit); q.st = q.it.st_begin; }
      }
    };
    // Neighbour-table entries of this lane's piece for its 8 vertices at position q; the piece lies inside ONE filter
    // tap (C % 4 == 0), so the lane reads one row of the table.  Eight independent, unconditional loads and NO
    // arithmetic on their results here: this in-order warp would stall for the L2 latency at the first use (a "+ 1"
    // right behind each load serialised the eight loads - 5 us per stage); the values are consumed one stage later.
    auto fetch_rows = [&](const Pos &q, int *raw) {
      const int k = q.it.kt * kWM + 32 * g + 4 * u;
      const int f = min((int)__umulhi((uint32_t)k, p.magic_c), p.F - 1);
      const int h0 = q.st * kWStageV + vsub;
      if (p.nbr) {
        const int64_t base = (int64_t)f * p.nbr_ld;
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = load_idx<IdxT>(p.nbr, base + min(h0 + 4 * i, H - 1));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = h0 + 4 * i;
      }
    };
    Pos pi;                                                    // next stage this team issues
    pi.item = blockIdx.x; pi.count = 0; pi.st = 0;
    pi.it = pi.item < n_items ? w_item(p, pi.item, n_stages, vsplit) : WItem{0, 0, 0, 0};
    pi.st = pi.it.st_begin;
    for (int k = 0; k < team && pos_valid(pi); ++k) pos_step(pi);
    int rows_next[8];
    if (pos_valid(pi)) fetch_rows(pi, rows_next);
    const int64_t xoff = p.nbr ? 0 : -(int64_t)p.ldX;          // without a neighbour table X has no sink row: row h+1 -> h
    bool pending = false;
    uint32_t pend_slot = 0;
    // dbias (column sums of G): the producers of the kt == 0 items see every G row of their output slice exactly once
    // (each (vertex range, output slice) pair has one kt == 0 item; rows beyond H are zero-filled), so they add what they
    // convert anyway: lane = 4 columns x the 8 rows it copied.  Flushed when the output slice changes and at the end.
    const bool want_bias = p.dbias != nullptr && has_g;
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    int csum_np = -1, pend_np = -1;                           // pend_np >= 0: the pending stage belongs to a kt == 0 item
    auto flush_csum = [&]() {
      if (csum_np >= 0) {
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {
          csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o); csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
          csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o); csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
        }
        if (vsub == 0) {
          float *db = p.dbias + csum_np * Np + 32 * g + 4 * u;
          if (csum.x != 0.f) atomicAdd(db, csum.x);
          if (csum.y != 0.f) atomicAdd(db + 1, csum.y);
          if (csum.z != 0.f) atomicAdd(db + 2, csum.z);
          if (csum.w != 0.f) atomicAdd(db + 3, csum.w);
        }
      }
      csum = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // landed stage -> "small" tiles -> publish.  keep_newest: the copy group committed last belongs to the NEXT stage
    auto complete_pending = [&](bool keep_newest) {
      if (keep_newest) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
      const uint32_t sbase = smem_base + pend_slot * stage_bytes;
      const bool sum_g = want_bias && pend_np >= 0;           // (warp-uniform)
      if (sum_g && pend_np != csum_np) { flush_csum(); csum_np = pend_np; }
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        if ((part == 1 && !has_g) || (p.dbg & 2)) break;
        const uint32_t raw = sbase + (part ? 2u * a_bytes : 0u), small = raw + (part ? g_bytes : a_bytes);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(raw + doff[i]));
          if (part == 1 && sum_g) { csum.x += x.x; csum.y += x.y; csum.z += x.z; csum.w += x.w; }
          float4 sm;
          sm.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
          sm.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          sm.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
          sm.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(small + doff[i]), "f"(sm.x), "f"(sm.y), "f"(sm.z), "f"(sm.w) : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // cp.async / st.shared data -> visible to the tensor core
      __syncwarp();
      if (lane == 0) w_mbar_arrive(bar_full + 8 * pend_slot);
      pending = false;
    };
    // A team's consecutive stages are kWTeams apart in the ring.  Issuing the next stage's copies BEFORE converting the
    // pending one hides the copy latency - but only if the next stage's slot is free.  When it is not (always with a ring
    // of two: the slot is the pending one), waiting for it with the pending stage unpublished stalls the MMA issuer that
    // has to free it: measured, every stage then cost one full MMA + wake-up + issue + convert round trip (1.2 - 1.5 us
    // whatever its size).  So: slot free -> issue first, convert while the copies fly; slot busy -> publish first.
    const bool serial = ring <= (uint32_t)kWTeams;
    while (pos_valid(pi) || pending) {
      const bool issue = pos_valid(pi);
      const uint32_t slot = pi.count % ring, slot_par = ((pi.count / ring) & 1u) ^ 1u;
      if (pending && (serial || !issue || !__all_sync(0xffffffffu, w_mbar_test(bar_empty + 8 * slot, slot_par)))) complete_pending(false);
      if (issue) {
        w_mbar_wait(bar_empty + 8 * slot, slot_par, ns_prod);                       // the MMAs that read this slot have retired
        const uint32_t sbase = smem_base + slot * stage_bytes;
        const int h0 = pi.st * kWStageV + vsub;
        const int k = pi.it.kt * kWM + 32 * g + 4 * u;
        const int f = (int)__umulhi((uint32_t)k, p.magic_c);
        const float *xc = p.X + xoff + (k - f * p.C);
        const bool k_ok = k < p.K && !(p.dbg & 4);
        if (!(p.dbg & 32))
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = rows_next[i] + 1;                       // matrix row; 0 = absent neighbour (zero-filled)
          const uint32_t nbytes = (k_ok && h0 + 4 * i < H && row > 0) ? 16u : 0u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + doff[i]), "l"(nbytes ? xc + (int64_t)row * p.ldX : p.X),
                       "r"(nbytes)
                       : "memory");
        }
        if (has_g && !(p.dbg & 32)) {
          const float *gc = p.G + pi.it.np * Np + 32 * g + 4 * u;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int h = h0 + 4 * i;
            const uint32_t nbytes = (h < H && !(p.dbg & 4)) ? 16u : 0u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + 2u * a_bytes + doff[i]), "l"(gc + (int64_t)min(h, H - 1) * p.ldG),
                         "r"(nbytes)
                         : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int issued_np = pi.it.kt == 0 ? pi.it.np : -1;
        // this team's next stage: prefetch its row indices while the copies fly
        for (int k = 0; k < kWTeams && pos_valid(pi); ++k) pos_step(pi);
        if (pos_valid(pi) && !(p.dbg & 64)) fetch_rows(pi, rows_next);
        if (pending) complete_pending(true);
        pending = true;
        pend_slot = slot;
        pend_np = issued_np;
      }
    }
    if (want_bias) flush_csum();
  } else if (warp == kWMmaWarp) {
    // ===================== MMA issuer =====================
    const uint32_t tmem = __shfl_sync(0xffffffffu, *s_tmem, 0);
    const uint32_t idesc = w_idesc(Np), idesc2 = w_idesc(2 * Np);
    uint32_t count = 0, cuts = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const WItem it = w_item(p, item, n_stages, vsplit);
      for (int st = it.st_begin; st < it.st_end;) {
        const uint32_t as = p.acc_stages == 2 ? (cuts & 1u) : 0u, aph = p.acc_stages == 2 ? ((cuts >> 1) & 1u) : (cuts & 1u);
        w_mbar_wait(bar_acc_empty + 8 * as, aph ^ 1u, ns_mma);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // 3xTF32 = A_raw G_raw + A_raw G_small + A_small G_raw.  The G tile holds [raw groups | small groups] back to back at
        // the same group stride, so A_raw x [G_raw | G_small] is ONE MMA with N' = 2 Np into two adjacent accumulators
        // (main | correction); A_small x G_raw joins the correction accumulator: 8 MMAs per stage instead of 12, issued as
        // one block (w_umma_stage_ss).  The epilogue adds the two accumulators with round-to-nearest adds.
        const uint32_t d = tmem + as * (uint32_t)(2 * Np), d_corr = d + (uint32_t)Np;
        const int cut_end = min(it.st_end, st + kWCutStages);
        for (bool first = true; st < cut_end; ++st, ++count) {
          const uint32_t slot = count % ring;
          w_mbar_wait(bar_full + 8 * slot, (count / ring) & 1u, ns_mma);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sbase = smem_base + slot * stage_bytes;
          const uint32_t a_raw = sbase, a_small = sbase + a_bytes, g_raw = sbase + 2u * a_bytes;
          w_umma_stage_ss(d, d_corr, w_desc(a_raw, 4096u), w_desc(a_small, 4096u), w_desc(g_raw, 4096u), idesc2, idesc, !first, bar_empty + 8 * slot,
                          (p.dbg & 1) ? 1 : kWStageV / 8);
          first = false;
        }
        w_commit(bar_acc_full + 8 * as);
        ++cuts;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                                    // TMEM lane quarter of this warp (warps 9..12 -> 1, 2, 3, 0)
    const uint32_t pitch_b = (uint32_t)pitch * 4u;
    const uint32_t tile = sum_base + (uint32_t)(q * 32) * pitch_b;
    uint32_t cuts = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const WItem it = w_item(p, item, n_stages, vsplit);
      const int n_cut = (it.st_end - it.st_begin + kWCutStages - 1) / kWCutStages;
      for (int c = 0; c < n_cut; ++c, ++cuts) {
        const uint32_t as = p.acc_stages == 2 ? (cuts & 1u) : 0u, aph = p.acc_stages == 2 ? ((cuts >> 1) & 1u) : (cuts & 1u);
        const bool last = c == n_cut - 1;
        w_mbar_wait(bar_acc_full + 8 * as, aph, ns_epi);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int cb = 0; cb < Np; cb += 32) {
          if (p.dbg & 8) break;
          float vv[32];
          {
            const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)(2 * Np) + cb;
            float cc[32];
            w_tmem_ld32(acc, vv);
            w_tmem_ld32(acc + (uint32_t)Np, cc);                  // correction products (round-to-nearest add)
#pragma unroll
            for (int i = 0; i < 32; ++i) vv[i] += cc[i];
          }
          const uint32_t my_row = tile + (uint32_t)lane * pitch_b;
          {
            // running sum of the item's chain cuts in tensor memory: each thread adds into its own lane's columns - no shared
            // memory until the last cut, whose 32-column block goes through the small staging tile for the transposed reductions
            const uint32_t racc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.acc_stages * 2 * Np) + (uint32_t)cb;
            if (c > 0) {
              float rr[32];
              w_tmem_ld32(racc, rr);
#pragma unroll
              for (int i = 0; i < 32; ++i) vv[i] += rr[i];
            }
            if (!last) { w_tmem_st32(racc, vv); continue; }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (uint32_t)u * 16u), "f"(vv[4 * u]), "f"(vv[4 * u + 1]),
                         "f"(vv[4 * u + 2]), "f"(vv[4 * u + 3])
                         : "memory");
          __syncwarp();
          // transposed read-back: every reduction instruction covers 4 rows x 128 contiguous bytes of dWt
          const int k_warp = it.kt * kWM + q * 32;
#pragma unroll
          for (int it8 = 0; it8 < 8; ++it8) {
            const int row = it8 * 4 + (lane >> 3);
            float4 o;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                         : "r"(tile + (uint32_t)row * pitch_b + (uint32_t)(lane & 7) * 16u));
            const int k = k_warp + row;
            if (k < p.K) atomicAdd(reinterpret_cast<float4 *>(p.dWt + (int64_t)k * p.M + it.np * Np + cb + 4 * (lane & 7)), o);
          }
          __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) w_mbar_arrive(bar_acc_empty + 8 * as);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kWMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace
}  // namespace efgh

using namespace efgh;

static size_t wgrad_tc_smem(int Np, int *ring_out) {
  const size_t stage = 2 * 4 * 4096 + 2 * (size_t)(Np / 32) * 4096;
  const size_t fixed = (size_t)kWM * 36 * 4 + 256 + 1024;
  int ring = (int)((226 * 1024 - fixed) / stage);
  if (ring > kWMaxStages) ring = kWMaxStages;
  if (ring_out) *ring_out = ring;
  return ring * stage + fixed;
}

extern "C" int efgh_bcl_conv_wgrad_tc_supported(int C, int F, int M) {
  if (C < 4 || C % 4 != 0 || F < 1 || M < 32 || M > 256 || M % 32 != 0) return 0;
  if (M > 128 && M % 128 != 0) return 0;
  return 1;
}

extern "C" int efgh_bcl_conv_wgrad_tc(const float *X, int64_t ldX, int C, const void *nbr, int idx_bits, int64_t nbr_ld, int F,
                                      int64_t h, const int32_t *h_dev, const float *dY, int64_t ldY, int M, float *dWt,
                                      float *dbias, void *stream) {
  if (!nbr) F = 1;
  EFGH_REQUIRE(efgh_bcl_conv_wgrad_tc_supported(C, F, M), "efgh_bcl_conv_wgrad_tc: unsupported shape C=%d F=%d M=%d", C, F, M);
  EFGH_REQUIRE(h >= 0 && h < (1ll << 30), "efgh_bcl_conv_wgrad_tc: bad h");
  if (h == 0) return EFGH_OK;
  EFGH_REQUIRE(X && dY && dWt, "efgh_bcl_conv_wgrad_tc: null pointer");
  EFGH_REQUIRE(ldX % 4 == 0 && ldY % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(dY) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dWt) & 15) == 0,
               "efgh_bcl_conv_wgrad_tc: X, dY and dWt must be 16-byte aligned with leading dimensions multiple of 4");
  EFGH_REQUIRE(idx_bits == 32 || idx_bits == 64, "efgh_bcl_conv_wgrad_tc: idx_bits must be 32 or 64");
  WgradParams p;
  p.X = X; p.ldX = ldX; p.C = C; p.nbr = nbr; p.nbr_ld = nbr_ld; p.F = F; p.h_host = (int)h; p.h_dev = h_dev;
  p.G = dY; p.ldG = ldY; p.M = M; p.dWt = dWt; p.dbias = dbias; p.K = F * C;
  p.n_kt = (p.K + kWM - 1) / kWM;
  p.Np = M > 128 ? 128 : M;
  p.n_np = M / p.Np;
  p.magic_c = (uint32_t)(((1ull << 32) + (uint64_t)C - 1) / (uint64_t)C);
  p.dbg = g_wgrad_flags;
  // tensor memory (512 columns): acc_stages x (main | correction) x Np + the running sum of the chain cuts (Np) - keeping
  // that sum in tensor memory instead of a 128 x Np shared-memory tile is what leaves room for a ring of 5 / 4 / 3 stages
  p.acc_stages = (2 * 2 * p.Np + p.Np) <= 512 ? 2 : 1;
  const size_t smem = wgrad_tc_smem(p.Np, &p.ring);
  EFGH_REQUIRE(p.ring >= 2, "efgh_bcl_conv_wgrad_tc: no room for two stages");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = sm_count();
  if (idx_bits == 32) {
    auto kern = k_wgrad_tc<int32_t>;
    EFGH_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kWThreads, smem, s>>>(p);
  } else {
    auto kern = k_wgrad_tc<int64_t>;
    EFGH_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kWThreads, smem, s>>>(p);
  }
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}
