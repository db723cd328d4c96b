// Lattice convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   Y[h, m] (+)= act(bias[m] + sum_{f,c} X[nbr[f,h]+1, c] * W[(f,c), m])
//
// reference nets/bilateralNN.py:240-244 gathers a (1, C, F, H) tensor (218 MB at level 0) and runs a cuDNN
// (F,1) convolution over it.  Here the gather is the A-operand path of a warp-specialised GEMM and the
// gathered tensor only ever exists 128 rows x 32 floats at a time:
//
//   warps 0-15  gather producers, four teams of 128 threads (thread = one vertex row of the 128-row tile);
//               team t handles K chunks t, t+4, ... of every work item, so four chunks are always being
//               prepared concurrently (one warp per scheduler per team; a single warp is latency-bound).
//               Per K chunk (32 floats = 128 B of the row) a team
//                 1. cp.async's (LDGSTS, zero-fill for absent neighbours) whole 128-byte pieces of neighbour
//                    rows of the normalised, vertex-major splat matrix from L2 into a warp-PRIVATE slot of a
//                    shared-memory ring (8 lanes per row piece = one cache line per 8 lanes) - several chunks
//                    in flight, no barrier needed because only the issuing warp reads its slot back;
//                 2. each thread re-reads its row's landed 128 B, applies the optional deferred bias + ReLU of a split-K
//                    producer, splits each value into a TF32-exact "big" part and the fp32 remainder
//                    "small", and writes both straight into TENSOR MEMORY with tcgen05.st (lane = row,
//                    32 columns each) - the A operand never goes back through shared memory;
//   warp  16    MMA issuer: one thread issues tcgen05.mma kind::tf32 with A from TMEM and B from shared
//               memory, M=128 x N x K=8, fp32 accumulators in TMEM.  3xTF32:
//               D += A_small*B_big + A_big*B_small + A_big*B_big (the dropped small*small term is 2^-22
//               relative), or one TF32 pass when nsplit == 1;
//   warp  17    weight loader: one thread issues cp.async.bulk (TMA engine, UBLKCP) per K chunk; the
//               weights were packed once (k_pack_weights) into the swizzled K-major image, big | small;
//   warps 18-21 epilogue: tcgen05.ld accumulators -> registers, then either bias + activation + 16-byte
//               row stores, or red.add.v4 of the raw partial sum into Y (split-K / chain cut, see below).
//
// Tensor-core accumulation rounds toward zero, so the error of a long accumulation chain grows linearly
// with its length; chains are CUT every kGroupChunks chunks and the partial sums are added with
// round-to-nearest adds.  Two ways, chosen by the output width:
//   N <= 128  on chip: a work item is a whole 128-vertex tile; the epilogue warps drain every cut's accumulator
//             (the other accumulator stage is being filled meanwhile) into a running sum held in shared memory -
//             the same tile that transposes the result for coalesced stores - and write Y ONCE, with bias and
//             activation applied.  No atomics, no zero-fill of Y, no deferred bias;
//   N == 256  (no room for a 128 KB running sum next to the weight stages) in L2: work item = (tile, K group),
//             partial sums added with red.add.v4 into a zero-filled Y, bias + activation deferred to the consumer.
//             The same split keeps all 148 SMs busy on the small lattices of the deep levels.
// Persistent CTAs, one per SM, stride over the items; the vertex count is read from device memory.
//
// TMEM map (512 columns): [accumulators: acc_stages x nacc x N] [A operand: 4 teams x (32 big | 32 small)].
#include "common.cuh"

static unsigned long long *g_conv_trace = nullptr;
// Debug hook (not part of the public header): device buffer of 5 roles x 512 x 2 u64 that CTA 0 fills with
// (event, globaltimer) pairs, or NULL to disable.
extern "C" void efgh_debug_set_conv_trace(unsigned long long *buf) { g_conv_trace = buf; }
// Debug / timing-study switches (tools/conv_tc_time.py, tools/tf32_truncation_probe.py; not part of the public header):
//   bit 1        publish a converted chunk after the next copy issue instead of at once
//   bit 2        one-pass TF32 only: hand the tensor core unmasked fp32 A
//   bits 4-7     cap on raw-row slots      bits 8-13   chunks per K group
//   bits 16-23   producers' poll sleep / 8 ns (0xff = none)      bits 24-31   relaxed waits' sleep / 8 ns
static int g_conv_flags = 0;
extern "C" void efgh_debug_set_conv_flags(int flags) { g_conv_flags = flags; }

namespace efgh {
namespace {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                  // floats per K chunk = one 128-byte row piece
constexpr int kTeams = 4;                    // producer teams; team t converts chunks t, t+4, ... of every item
constexpr int kProducerWarps = 4 * kTeams;
constexpr int kMmaWarp = kProducerWarps, kTmaWarp = kProducerWarps + 1, kEpiWarp0 = kProducerWarps + 2;
constexpr int kThreads = (kEpiWarp0 + 4) * 32;  // 704
constexpr int kMaxRaw = 1;                   // raw-row slots per producer warp.  One: with 16 producer warps per SM the copy latency is
                                             // hidden by the other warps, deeper rings were measured no faster inside the kernel, and the
                                             // 64-128 KB of shared memory they cost made the whole launch sequence 3.6 % slower
constexpr int kRawSlotBytes = kTeams * kTileM * 128;  // one raw slot for all teams (512 threads x 128 B)
constexpr int kEpiRowFloats = 36;            // padded row of the epilogue staging tile (bank-conflict free)
constexpr int kBiasFloats = 256 + 512;       // output bias (N <= 256) + input bias (C <= 512) staged in shared memory

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"   // suspend-time hint: sleep in hardware instead of
      "@p bra LAB_DONE;\n\t"                                            // spinning - a spinning warp steals issue slots from
      "bra LAB_WAIT;\n\t"                                               // the one producer warp per scheduler that has work
      "LAB_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(20000u)
      : "memory");
}
// Same, but yields issue slots between polls: used by roles whose wake-up latency is not critical (their
// spin loops would otherwise compete with the producer warps on the same scheduler).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t ns = 64) {
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
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem descriptor].  Executed by the WHOLE (converged) MMA warp; one elected lane
// issues.  Keeping the warp converged lets the compiler hold the operands in uniform registers - under an
// `if (lane == 0)` it wraps every tcgen05.mma in a value-uniformising loop.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// One K chunk of the 3xTF32 convolution with separate correction accumulators, issued under ONE election: 4 k-steps x
// [A_big x (B_big | B_small) -> d_main (N' = 2N), A_small x B_big -> d_sb (N)], then the two commits that free the
// team's A stage and the weight stage.  The MMA issuer's per-chunk instruction chain gates every team's A-stage round
// trip (a variant with three more uniform-register moves per MMA was measured 5 % slower), so the election, the
// predicate set-up and the operand increments happen once per chunk, not once per MMA.
__device__ __forceinline__ void umma_chunk_3x(uint32_t d_main, uint32_t d_sb, uint32_t a_big, uint64_t desc_b, uint32_t idesc2,
                                               uint32_t idesc, uint32_t acc_main, uint32_t acc_sb, uint32_t bar_a, uint32_t bar_b) {
  asm volatile(
      "{\n\t"
      ".reg .pred e, pm, ps;\n\t"
      ".reg .b32 a1, a2, a3, s0, s1, s2, s3;\n\t"
      ".reg .b64 b1, b2, b3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pm, %6, 0;\n\t"
      "setp.ne.b32 ps, %7, 0;\n\t"
      "add.u32 a1, %2, 8;\n\t"
      "add.u32 a2, %2, 16;\n\t"
      "add.u32 a3, %2, 24;\n\t"
      "add.u32 s0, %2, 32;\n\t"
      "add.u32 s1, %2, 40;\n\t"
      "add.u32 s2, %2, 48;\n\t"
      "add.u32 s3, %2, 56;\n\t"
      "add.u64 b1, %3, 2;\n\t"
      "add.u64 b2, %3, 4;\n\t"
      "add.u64 b3, %3, 6;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %4, pm;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [s0], %3, %5, ps;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a1], b1, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [s1], b1, %5, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a2], b2, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [s2], b2, %5, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a3], b3, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [s3], b3, %5, 1;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
      "}" ::"r"(d_main), "r"(d_sb), "r"(a_big), "l"(desc_b), "r"(idesc2), "r"(idesc), "r"(acc_main), "r"(acc_sb), "r"(bar_a), "r"(bar_b)
      : "memory");
}
// The same for ONE accumulator (N >= 128: no tensor-memory columns to spare for correction accumulators): 4 k-steps x
// [A_small x B_big, A_big x B_small, A_big x B_big] -> d, then the two commits, under one election with the operand
// increments done once.  tools/umma_chain_probe.cu: an MMA issued from a block like this costs 15 / 18 / 36 / 71 ns at
// N = 32 / 64 / 128 / 256 (the math floor), ~100 ns of the issuing warp's time from a per-MMA loop.  Measured effect here:
// level-2 conv1 (N = 128) 455 -> 445 us per 16 scans - the issuer was not the bound, the block is simply the cheaper form.
__device__ __forceinline__ void umma_chunk_3x_1acc(uint32_t d, uint32_t a_big, uint64_t desc_bb, uint64_t desc_bs, uint32_t idesc,
                                                    uint32_t acc_first, uint32_t bar_a, uint32_t bar_b) {
  asm volatile(
      "{\n\t"
      ".reg .pred e, pf;\n\t"
      ".reg .b32 a1, a2, a3, s0, s1, s2, s3;\n\t"
      ".reg .b64 b1, b2, b3, c1, c2, c3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %5, 0;\n\t"
      "add.u32 a1, %1, 8;\n\t"
      "add.u32 a2, %1, 16;\n\t"
      "add.u32 a3, %1, 24;\n\t"
      "add.u32 s0, %1, 32;\n\t"
      "add.u32 s1, %1, 40;\n\t"
      "add.u32 s2, %1, 48;\n\t"
      "add.u32 s3, %1, 56;\n\t"
      "add.u64 b1, %2, 2;\n\t"
      "add.u64 b2, %2, 4;\n\t"
      "add.u64 b3, %2, 6;\n\t"
      "add.u64 c1, %3, 2;\n\t"
      "add.u64 c2, %3, 4;\n\t"
      "add.u64 c3, %3, 6;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [s0], %2, %4, pf;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %3, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [s1], b1, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a1], c1, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a1], b1, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [s2], b2, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a2], c2, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a2], b2, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [s3], b3, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a3], c3, %4, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [a3], b3, %4, 1;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
      "}" ::"r"(d), "r"(a_big), "l"(desc_bb), "l"(desc_bs), "r"(idesc), "r"(acc_first), "r"(bar_a), "r"(bar_b)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(bar)
      : "memory");
}

// K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO = 64 x 16 B), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr & 0x3ffff) >> 4) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
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

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.1f * v;
  return v;
}

// Optional timeline trace (debug): role r of CTA 0 appends (event, globaltimer) pairs to trace[r*kTraceCap...]
constexpr int kTraceCap = 512;
// Role timeline hook: compiled in only with -DEFGH_CONV_TRACE (tools/conv_tc_trace.py); the production kernel is
// instruction-issue bound and every trace point costs four instructions per K chunk.
__device__ __forceinline__ void trace_ev(unsigned long long *trace, int role, int &n, int ev) {
#ifndef EFGH_CONV_TRACE
  (void)trace; (void)role; (void)n; (void)ev;
  return;
#endif
  if (trace && blockIdx.x == 0 && n < kTraceCap) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(size_t)role * kTraceCap * 2 + 2 * n] = (unsigned long long)ev;
    trace[(size_t)role * kTraceCap * 2 + 2 * n + 1] = t;
    ++n;
  }
}

struct ConvParams {
  unsigned long long *trace; int dbg_flags;
  const float *X; int64_t ldX; int C;
  const float *in_bias; int in_act;     // optional input transform x = act(x + in_bias[c]) (deferred epilogue of a split-K producer)
  const void *nbr; int64_t nbr_ld; int F;
  int h_host; const int32_t *h_dev;
  const float *Wimg; const float *bias; int N; int act;
  float *Y; int64_t ldY;
  int n_chunks; int n_groups;
  int cut_chunks;                       // > 0: on-chip chain cuts - an item's chunks are accumulated in runs of <= cut_chunks, summed by the epilogue
  int b_stages, raw_slots, acc_stages, nacc;   // nacc: independent accumulators per stage (see the MMA issuer)
  uint32_t magic_c;                     // ceil(2^32 / C): k / C == __umulhi(k, magic_c) for the k range used here
  int accumulate;                       // 1: red.add raw partial sums into pre-zeroed Y (bias/act deferred); 0: store act(bias + acc)
};

// K chunks [begin, end) of group g when n_chunks are dealt as evenly as possible to n_groups
__device__ __forceinline__ void group_range(int n_chunks, int n_groups, int g, int &begin, int &end) {
  const int base = n_chunks / n_groups, rem = n_chunks % n_groups;
  begin = g * base + min(g, rem);
  end = begin + base + (g < rem ? 1 : 0);
}

// Position in one producer team's chunk sequence.  The CTA's chunks (all chunks of item blockIdx.x, then of
// item blockIdx.x + gridDim.x, ...) are numbered 0, 1, 2, ...; team t takes numbers t, t+kTeams, ...
// A team crosses into a new item every 1-2 chunks, so the crossing must be cheap: the chunk range of every K
// group comes from a shared-memory table (gb[g] .. gb[g+1]) and item -> (tile, group) is advanced
// incrementally (no integer division in the loop).
constexpr int kMaxGroups = 256;
struct TeamPos {
  int item, j, j_end, tile, g;
  __device__ __forceinline__ bool valid(int n_items) const { return item < n_items; }
  // carry j >= j_end over into the following items
  __device__ __forceinline__ void settle(const ConvParams &p, int n_items, const int *gb, int dq, int dr) {
    while (item < n_items && j >= j_end) {
      const int over = j - j_end;
      item += (int)gridDim.x;
      tile += dq; g += dr;
      if (g >= p.n_groups) { g -= p.n_groups; ++tile; }
      if (item >= n_items) break;
      j = gb[g] + over; j_end = gb[g + 1];
    }
  }
  __device__ __forceinline__ void init(const ConvParams &p, int n_items, int team, const int *gb, int dq, int dr) {
    item = (int)blockIdx.x; j = j_end = 0;
    tile = item / p.n_groups; g = item - tile * p.n_groups;
    if (item < n_items) {
      j = gb[g] + team; j_end = gb[g + 1];
      settle(p, n_items, gb, dq, dr);
    }
  }
  __device__ __forceinline__ void advance(const ConvParams &p, int n_items, const int *gb, int dq, int dr) {
    j += kTeams;
    settle(p, n_items, gb, dq, dr);
  }
};

// COMBINE: an item's contraction is cut into several accumulation chains that the epilogue sums in shared memory.  A
// compile-time switch: without it (one chain per item - 1x1 convolutions, K groups added in L2) the epilogue keeps its
// constant-pitch staging tile and the MMA issuer its per-item loop; every instruction either executes per item gates the
// other roles (measured: the run-time version cost the short 1x1 convolutions 35 %).
template <typename IdxT, int NSPLIT, bool COMBINE>
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [B ring: b_stages x (B_big | B_small?)] [raw ring: raw_slots x 512 threads x 128 B] [epilogue staging] [barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int N = p.N;
  const uint32_t b_bytes = (uint32_t)N * 128u * (NSPLIT == 3 ? 2 : 1);
  const uint32_t raw_base = smem_base + (uint32_t)p.b_stages * b_bytes;
  const uint32_t epi_base = raw_base + (uint32_t)p.raw_slots * kRawSlotBytes;   // 4 epilogue warps x 32 rows x epi_pitch floats
  const int epi_pitch = COMBINE ? N + 4 : kEpiRowFloats;                        // on-chip cuts: the whole 128 x N running sum lives here
  float *s_bias = reinterpret_cast<float *>(smem + (size_t)p.b_stages * b_bytes + (size_t)p.raw_slots * kRawSlotBytes + (size_t)(4 * 32 * 4) * epi_pitch);
  float *s_in_bias = s_bias + 256;
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_bias + kBiasFloats);
  // barrier map (8 B each): A_full[4 teams] A_empty[4] (2 spare each) B_full[4] B_empty[4] acc_full[2] acc_empty[2]
  const uint32_t bar_a_full = smem_u32(bars), bar_a_empty = bar_a_full + 8 * 6, bar_b_full = bar_a_empty + 8 * 6,
                 bar_b_empty = bar_b_full + 8 * 4, bar_acc_full = bar_b_empty + 8 * 4, bar_acc_empty = bar_acc_full + 16;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 24);
  int *s_gb = reinterpret_cast<int *>(bars + 26);        // [n_groups + 1] first chunk of every K group
  int *s_rx = s_gb + kMaxGroups + 4 + (threadIdx.x >> 5) * 64;   // per producer warp: row indices of the chunk being issued (2 taps x 32)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // producers poll their TMEM A stage's barrier with a sleep in between: the polling loops were 18 % of the kernel's
  // instructions (flags bits 16-23 override, 0xff = hardware try_wait loop without sleeping)
  const uint32_t spin_f = (uint32_t)p.dbg_flags >> 16 & 0xffu;
  const uint32_t spin_ns = spin_f == 0xffu ? 0u : (spin_f ? spin_f * 8u : 128u);
  const uint32_t relax_ns = ((uint32_t)p.dbg_flags >> 24 & 0xffu) ? ((uint32_t)p.dbg_flags >> 24 & 0xffu) * 8u : 64u;
  const int H = p.h_dev ? min(*p.h_dev, p.h_host) : p.h_host;
  const int n_tiles = (H + kTileM - 1) / kTileM;
  const int n_items = n_tiles * p.n_groups;
  const int K = p.F * p.C;
  constexpr uint32_t kAStageCols = NSPLIT == 3 ? 64 : 32;
  const uint32_t a_ring_col = (uint32_t)(p.acc_stages * p.nacc * N);

  for (int i = threadIdx.x; i <= p.n_groups; i += kThreads) {
    int gbeg = p.n_chunks, gend;
    if (i < p.n_groups) group_range(p.n_chunks, p.n_groups, i, gbeg, gend);
    s_gb[i] = gbeg;
  }
  int *s_cut = s_gb + kMaxGroups + 4 + kProducerWarps * 64;    // COMBINE: [0] number of cuts, [1 + c] first chunk of cut c (last entry: n_chunks)
  if (COMBINE && threadIdx.x == 0) {
    const int n_cut = max(1, (p.n_chunks + p.cut_chunks - 1) / p.cut_chunks);
    s_cut[0] = n_cut;
    for (int c = 0; c <= n_cut; ++c) {
      int b, e;
      group_range(p.n_chunks, n_cut, min(c, n_cut - 1), b, e);
      s_cut[1 + c] = c < n_cut ? b : p.n_chunks;
    }
  }
  const int it_dq = (int)gridDim.x / p.n_groups, it_dr = (int)gridDim.x % p.n_groups;   // item += gridDim.x in (tile, group) terms
  for (int i = threadIdx.x; i < N; i += kThreads) s_bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
  if (p.in_bias)
    for (int i = threadIdx.x; i < p.C; i += kThreads) s_in_bias[i] = __ldg(p.in_bias + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) { mbar_init(bar_a_full + 8 * i, 4); mbar_init(bar_a_empty + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_b_full + 8 * i, 1); mbar_init(bar_b_empty + 8 * i, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_acc_full + 8 * a, 1); mbar_init(bar_acc_empty + 8 * a, 4); }
    fence_barrier_init();
  }
  if (warp == kTmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < kProducerWarps) {
    // ===================== gather producers =====================
    const int team = warp >> 2;
    const int r = (warp & 3) * 32 + lane;                                     // tile row owned by this thread
    const uint32_t swz = (uint32_t)(lane & 7);
    const uint32_t warp_raw = raw_base + (uint32_t)warp * 4096u;              // this warp's 32 rows x 128 B (+ slot * kRawSlotBytes)
    const uint32_t my_raw = warp_raw + (uint32_t)lane * 128u;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;             // TMEM lanes this warp may touch
    const uint32_t a_t = tmem_base + lane_base + a_ring_col + (uint32_t)team * kAStageCols;
    const bool tracer = r == 0 && team < 2;
    int ntrace = 0;

    // Matrix rows (taps f, f+1) of this thread's vertex for the chunk at `pos`: neighbour index + 1, so 0 means
    // "absent" and addresses the all-zero sink row; without a neighbour table (1x1 convolution) the vertex's own row,
    // clamped into range (rows >= H are computed on garbage and dropped by the epilogue).  The loaded values are
    // used two chunks later - any arithmetic on them right here would stall this in-order warp for the L2 latency,
    // so the "+ 1" happens when they are shuffled out.
    auto fetch_rows = [&](const TeamPos &pos, int &ra, int &rb) {
      ra = rb = -1;
      if (!pos.valid(n_items)) return;
      const int h = pos.tile * kTileM + r;
      if (!p.nbr) { ra = min(h, H - 1) - 1; return; }
      if (h >= H) return;
      const int f = (int)__umulhi((uint32_t)(pos.j * kChunkK), p.magic_c);
      ra = load_idx<IdxT>(p.nbr, f * p.nbr_ld + h);
      if (f + 1 < p.F) rb = load_idx<IdxT>(p.nbr, (f + 1) * p.nbr_ld + h);
    };

    const uint32_t always = p.nbr ? 0u : 1u;                // without a sink row every (in-K) piece is copied

    TeamPos pi, pc, pp;                                     // issue / convert / row-prefetch positions
    pi.init(p, n_items, team, s_gb, it_dq, it_dr);
    pc = pi; pp = pi;
    int r0a, r0b, r1a, r1b, r2a, r2b;                       // row pairs for pi, pi+1, pi+2
    fetch_rows(pp, r0a, r0b);
    pp.advance(p, n_items, s_gb, it_dq, it_dr); fetch_rows(pp, r1a, r1b);
    pp.advance(p, n_items, s_gb, it_dq, it_dr); fetch_rows(pp, r2a, r2b);
    uint32_t n_inflight = 0, slot_i = 0, slot_c = 0, a_ph = 0;
    bool st_pending = false;

    while (pc.valid(n_items)) {
      while (pi.valid(n_items) && (int)n_inflight < p.raw_slots) {
        // 8 lanes cooperate on one row (8 x 16 B = one full 128-byte line), 4 rows per instruction, so every
        // LDGSTS touches whole cache lines; the neighbour row index comes from the lane that owns the row.
        const uint32_t dst = warp_raw + slot_i * kRawSlotBytes;
        const int k0 = pi.j * kChunkK;
        const int f = (int)__umulhi((uint32_t)k0, p.magic_c);
        const int cu = k0 - f * p.C + 4 * (lane & 7);
        const bool wrap = cu >= p.C;
        const int c_u = wrap ? cu - p.C : cu;
        const bool in_k = k0 + 4 * (lane & 7) < K;
        // The 8 lanes of a row group need the row index held by the lane that owns the row: exchanged through 256
        // bytes of warp-private shared memory (2 stores + 8 loads instead of 16 shuffles + 8 selects).
        __syncwarp();
        s_rx[lane] = r0a + 1;                                 // matrix row; 0 = absent neighbour = the all-zero sink row
        s_rx[32 + lane] = r0b + 1;
        __syncwarp();
        const int *rx = s_rx + (wrap ? 32 : 0) + (lane >> 3);
        int rows[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rows[i] = rx[4 * i];
        if (tracer) trace_ev(p.trace, team, ntrace, 5);
        const uint32_t dst_lane = dst + ((uint32_t)lane >> 3) * 128u;
        const uint32_t u_lane = (uint32_t)lane & 7u;
        const uint64_t xc = reinterpret_cast<uint64_t>(p.X + c_u);
        const uint32_t ldx4 = (uint32_t)p.ldX * 4u;           // row pitch in bytes (32-bit); mad.wide.u32 forms the full 64-bit offset
        const uint32_t kmask = in_k ? 0xffffffffu : 0u;
        // Per copy: one wide multiply-add (address) and the zero-fill predicate - this loop is instruction-issue bound;
        // a 64-bit multiply per row was a quarter of the kernel's instructions.  Absent neighbours (row 0) and K padding
        // are zero-FILLED (source size 0), never read: copying the sink row instead was measured 6x slower - every SM
        // hammering the same 144 bytes serialises in one L2 slice.
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t R7 = ((uint32_t)(4 * i) + ((uint32_t)lane >> 3)) & 7u;
          const uint32_t nbytes = (((uint32_t)rows[i] | always) & kmask) ? 16u : 0u;
          uint64_t src;
          asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(src) : "r"((uint32_t)rows[i]), "r"(ldx4), "l"(xc));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_lane + (uint32_t)i * 512u + ((u_lane ^ R7) << 4)), "l"(src),
                       "r"(nbytes)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (tracer) trace_ev(p.trace, team, ntrace, 6);
        pi.advance(p, n_items, s_gb, it_dq, it_dr);
        r0a = r1a; r0b = r1b; r1a = r2a; r1b = r2b;
        if (tracer) trace_ev(p.trace, team, ntrace, 7);
        pp.advance(p, n_items, s_gb, it_dq, it_dr); fetch_rows(pp, r2a, r2b);
        ++n_inflight;
        if (++slot_i == (uint32_t)p.raw_slots) slot_i = 0;
        if (tracer) trace_ev(p.trace, team, ntrace, 1);
      }
      if (st_pending) {                                     // previous chunk's TMEM stores had a whole issue to complete
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        { __syncwarp(); if (lane == 0) mbar_arrive(bar_a_full + 8 * team); }   // one arrival per warp: every arrival wakes the CTA's barrier sleepers
        st_pending = false;
      }
      if (n_inflight >= 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else if (n_inflight == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();                                         // the row pieces were fetched by 8 different lanes of this warp
      if (tracer) trace_ev(p.trace, team, ntrace, 2);

      // ---- landed chunk -> registers -> (bias/act) -> big | small -> tensor memory, 16 columns at a time
      const uint32_t src = my_raw + slot_c * kRawSlotBytes;
      const int cb0 = p.in_bias ? (pc.j * kChunkK - (int)__umulhi((uint32_t)(pc.j * kChunkK), p.magic_c) * p.C) : 0;
      bool waited = false;
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[4 * u]), "=f"(v[4 * u + 1]), "=f"(v[4 * u + 2]), "=f"(v[4 * u + 3])
                       : "r"(src + ((((uint32_t)(4 * hlf + u)) ^ swz) << 4)));
        }
        if (p.in_bias) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int c = cb0 + 16 * hlf + 4 * u;
            if (c >= p.C) c -= p.C;
            const float4 b = *reinterpret_cast<const float4 *>(s_in_bias + c);
            v[4 * u] = act_apply(v[4 * u] + b.x, p.in_act);
            v[4 * u + 1] = act_apply(v[4 * u + 1] + b.y, p.in_act);
            v[4 * u + 2] = act_apply(v[4 * u + 2] + b.z, p.in_act);
            v[4 * u + 3] = act_apply(v[4 * u + 3] + b.w, p.in_act);
          }
        }
        uint32_t big[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) big[i] = __float_as_uint(v[i]) & 0xffffe000u;
        if (NSPLIT == 1 && (p.dbg_flags & 4)) {             // probe: hand the tensor core unmasked fp32 (does kind::tf32 truncate?)
#pragma unroll
          for (int i = 0; i < 16; ++i) big[i] = __float_as_uint(v[i]);
        }
        if (!waited) {                                      // MMA done with this team's previous chunk
          if (spin_ns) mbar_wait_relaxed(bar_a_empty + 8 * team, a_ph ^ 1, spin_ns); else mbar_wait(bar_a_empty + 8 * team, a_ph ^ 1);
          tc_fence_after();
          waited = true;
        }
        tmem_st16(a_t + 16 * hlf, big);
        if (NSPLIT == 3) {
#pragma unroll
          for (int i = 0; i < 16; ++i) big[i] = __float_as_uint(v[i] - __uint_as_float(big[i]));
          tmem_st16(a_t + 32 + 16 * hlf, big);
        }
      }
      if (!(p.dbg_flags & 2)) {                             // publish at once: the MMA of this chunk gates the team's next A stage
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        { __syncwarp(); if (lane == 0) mbar_arrive(bar_a_full + 8 * team); }   // one arrival per warp: every arrival wakes the CTA's barrier sleepers
      } else {
        st_pending = true;
      }
      a_ph ^= 1;
      if (tracer) trace_ev(p.trace, team, ntrace, 3);
      pc.advance(p, n_items, s_gb, it_dq, it_dr);
      --n_inflight;
      if (++slot_c == (uint32_t)p.raw_slots) slot_c = 0;
    }
    if (st_pending) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      { __syncwarp(); if (lane == 0) mbar_arrive(bar_a_full + 8 * team); }   // one arrival per warp: every arrival wakes the CTA's barrier sleepers
    }
  } else if (warp == kTmaWarp) {
    // ===================== weight loader (TMA bulk copies) =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int gi = item % p.n_groups;
        const int j_begin = s_gb[gi], j_end = s_gb[gi + 1];
        for (int j = j_begin; j < j_end; ++j) {
          mbar_wait_relaxed(bar_b_empty + 8 * s, ph ^ 1, relax_ns);
          mbar_arrive_expect_tx(bar_b_full + 8 * s, b_bytes);
          // several 8 KB copies in flight per stage: one large bulk copy is latency-bound
          for (uint32_t off = 0; off < b_bytes; off += 8192u)
            bulk_g2s(smem_base + s * b_bytes + off, reinterpret_cast<const uint8_t *>(p.Wimg) + (size_t)j * b_bytes + off,
                     min(8192u, b_bytes - off), bar_b_full + 8 * s);
          if (++s == (uint32_t)p.b_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    {
      // shfl-broadcast marks the value warp-uniform for ptxas, so descriptors and TMEM addresses are computed in the
      // uniform datapath instead of costing an R2UR per operand per MMA
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
      const uint32_t idesc = make_idesc(N);
      uint32_t tcount = 0, sb = 0, phb = 0, pha_bits = 0, seq = 0;   // tcount: accumulator hand-overs (one per cut)
      int ntrace = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int gi = item % p.n_groups;
        const int i_begin = s_gb[gi], i_end = s_gb[gi + 1];
        const int n_cut = COMBINE ? s_cut[0] : 1;              // (COMBINE runs with one K group: every item has the same cuts)
        for (int c = 0; c < n_cut; ++c, ++tcount) {
        const int j_begin = COMBINE ? s_cut[1 + c] : i_begin, j_end = COMBINE ? s_cut[2 + c] : i_end;
        const uint32_t as = p.acc_stages == 2 ? (tcount & 1) : 0;
        const uint32_t aph = p.acc_stages == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
        if (lane == 0) trace_ev(p.trace, 3, ntrace, 100 + (int)tcount);
        mbar_wait(bar_acc_empty + 8 * as, aph ^ 1);      // epilogue drained this accumulator stage
        tc_fence_after();
        // 3xTF32 = A_big*B_big + A_big*B_small + A_small*B_big.  The weight image holds [big rows | small rows]
        // back to back, so when TMEM has room for separate accumulators (nacc >= 2) ONE descriptor with N' = 2N
        // makes A_big*[B_big | B_small] a single MMA writing two adjacent accumulators (bb | bs): 8 MMAs per
        // chunk instead of 12 - the warp is bound by MMA issue, not by tensor throughput, at these widths.  The
        // epilogue adds the accumulators with round-to-nearest adds; the tensor core rounds toward zero on every
        // accumulate, so keeping the tiny correction products out of the main chain also shortens it (measured
        // error 1.0e-6 instead of 1.7e-6 per layer).
        const uint32_t tmem_d0 = tmem_base + as * (uint32_t)(p.nacc * N);
        const uint32_t d_sb = tmem_d0 + (uint32_t)((p.nacc == 3 ? 2 : 1) * N);   // nacc 2: shares the bs accumulator
        const uint32_t idesc2 = make_idesc(2 * N);
        for (int j = j_begin; j < j_end; ++j) {
          const uint32_t team = seq++ & (kTeams - 1);
          mbar_wait(bar_a_full + 8 * team, (pha_bits >> team) & 1);
          if (lane == 0) trace_ev(p.trace, 3, ntrace, 1);
          mbar_wait(bar_b_full + 8 * sb, phb);
          tc_fence_after();
          if (lane == 0) trace_ev(p.trace, 3, ntrace, 2);
          const uint32_t a_big = tmem_base + a_ring_col + team * kAStageCols;      // (the remainder half follows at + 32 columns)
          const uint32_t b_big = smem_base + sb * b_bytes, b_small = b_big + (uint32_t)N * 128u;
          const uint64_t dbb = make_desc(b_big);
          const uint32_t first = j == j_begin;
          if (NSPLIT == 3 && p.nacc >= 2) {
            umma_chunk_3x(tmem_d0, d_sb, a_big, dbb, idesc2, idesc, !first, !(first && p.nacc == 3), bar_a_empty + 8 * team,
                          bar_b_empty + 8 * sb);
          } else if (NSPLIT == 3) {
            umma_chunk_3x_1acc(tmem_d0, a_big, dbb, make_desc(b_small), idesc, !first, bar_a_empty + 8 * team, bar_b_empty + 8 * sb);
          } else {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) umma_tf32_ts(tmem_d0, a_big + 8 * k4, dbb + 2 * k4, idesc, !(first && k4 == 0));
          }
          if (NSPLIT != 3) {
            umma_commit(bar_a_empty + 8 * team);             // frees the team's TMEM A stage when these MMAs retire
            umma_commit(bar_b_empty + 8 * sb);               // ... and the weight stage
          }
          if (lane == 0) trace_ev(p.trace, 3, ntrace, 3);
          pha_bits ^= 1u << team;
          if (++sb == (uint32_t)p.b_stages) { sb = 0; phb ^= 1; }
        }
        umma_commit(bar_acc_full + 8 * as);                  // accumulator complete -> epilogue
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    uint32_t tcount = 0;
    int ntrace = 0;
    const bool tracer = threadIdx.x == kEpiWarp0 * 32;
    const uint32_t pitch_b = (uint32_t)epi_pitch * 4u;                       // bytes per staged row
    const uint32_t stage = epi_base + (uint32_t)(warp - kEpiWarp0) * (32u * pitch_b);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = item / p.n_groups;
      const int n_cut = COMBINE ? s_cut[0] : 1;
      const int h_warp = tile * kTileM + q * 32;
      for (int c = 0; c < n_cut; ++c, ++tcount) {
      const uint32_t as = p.acc_stages == 2 ? (tcount & 1) : 0;
      const uint32_t aph = p.acc_stages == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
      const bool last = c == n_cut - 1;
      if (tracer) trace_ev(p.trace, 4, ntrace, 100 + (int)tcount);
      mbar_wait_relaxed(bar_acc_full + 8 * as, aph, relax_ns);
      tc_fence_after();
      if (tracer) trace_ev(p.trace, 4, ntrace, 1);
      // Each thread holds 32 consecutive columns of ITS row after tcgen05.ld.  The 32 x N block of this warp is staged
      // in shared memory, row = lane: (a) cuts of a long contraction are summed there with round-to-nearest adds
      // (every thread touches only its own row: no synchronisation until the last cut), (b) the last cut reads it back
      // transposed so that every global store / reduction instruction covers 4 rows x 128 contiguous bytes.
      for (int cb = 0; cb < N; cb += 32) {
        float v[32];
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)(p.nacc * N) + cb;
        tmem_ld32(tacc, v);
        for (int d = 1; d < p.nacc; ++d) {                    // add the independent partial accumulators (round-to-nearest adds)
          float w[32];
          tmem_ld32(tacc + (uint32_t)(d * N), w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += w[i];
        }
        if (tracer) trace_ev(p.trace, 4, ntrace, 10);
        const uint32_t my_row = stage + (uint32_t)lane * pitch_b + (COMBINE ? (uint32_t)cb * 4u : 0u);
        if (COMBINE && c > 0) {                               // running sum of the earlier cuts of this tile
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 r;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(my_row + (uint32_t)u * 16u));
            v[4 * u] += r.x; v[4 * u + 1] += r.y; v[4 * u + 2] += r.z; v[4 * u + 3] += r.w;
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (uint32_t)u * 16u),
                       "f"(v[4 * u]), "f"(v[4 * u + 1]), "f"(v[4 * u + 2]), "f"(v[4 * u + 3])
                       : "memory");
        if (COMBINE && !last) continue;
        // bias of the 4 columns this lane will store after the transposition (one 16-byte load per block)
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!p.accumulate) bv = *(reinterpret_cast<const float4 *>(s_bias + cb) + (lane & 7));
        __syncwarp();
        if (tracer) trace_ev(p.trace, 4, ntrace, 11);
        float4 o[8];
        const uint32_t blk = stage + (COMBINE ? (uint32_t)cb * 4u : 0u);
#pragma unroll
        for (int it8 = 0; it8 < 8; ++it8) {                   // all shared loads first, then all global stores: one warp cannot hide
          const int row = it8 * 4 + (lane >> 3);              // the load latency of an interleaved load/store sequence
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(o[it8].x), "=f"(o[it8].y), "=f"(o[it8].z), "=f"(o[it8].w)
                       : "r"(blk + (uint32_t)row * pitch_b + (uint32_t)(lane & 7) * 16u));
        }
        float *ybase = p.Y + (int64_t)(h_warp + (lane >> 3)) * p.ldY + cb + 4 * (lane & 7);
        const int h_lane = h_warp + (lane >> 3);
        if (p.accumulate) {
#pragma unroll
          for (int it8 = 0; it8 < 8; ++it8)
            if (h_lane + 4 * it8 < H) atomicAdd(reinterpret_cast<float4 *>(ybase + (int64_t)(4 * it8) * p.ldY), o[it8]);
        } else {
          if (p.act == 1) {
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
              o[it8].x = fmaxf(o[it8].x + bv.x, 0.f); o[it8].y = fmaxf(o[it8].y + bv.y, 0.f);
              o[it8].z = fmaxf(o[it8].z + bv.z, 0.f); o[it8].w = fmaxf(o[it8].w + bv.w, 0.f);
            }
          } else {
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
              o[it8].x += bv.x; o[it8].y += bv.y; o[it8].z += bv.z; o[it8].w += bv.w;
              if (p.act == 2) {
                o[it8].x = o[it8].x > 0.f ? o[it8].x : 0.1f * o[it8].x; o[it8].y = o[it8].y > 0.f ? o[it8].y : 0.1f * o[it8].y;
                o[it8].z = o[it8].z > 0.f ? o[it8].z : 0.1f * o[it8].z; o[it8].w = o[it8].w > 0.f ? o[it8].w : 0.1f * o[it8].w;
              }
            }
          }
#pragma unroll
          for (int it8 = 0; it8 < 8; ++it8)
            if (h_lane + 4 * it8 < H) *reinterpret_cast<float4 *>(ybase + (int64_t)(4 * it8) * p.ldY) = o[it8];
        }
        __syncwarp();                                         // (without on-chip cuts the next block re-uses this staging tile)
        if (tracer) trace_ev(p.trace, 4, ntrace, 12);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
      if (tracer) trace_ev(p.trace, 4, ntrace, 2);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Packs W (K, N) row-major (k = f*C + c) into the shared-memory image of the B operand:
//   for every 32-wide K chunk j: [big: N rows x 128 B, 16-byte units XOR-swizzled by (n & 7)] [small: same]
// big = value truncated to TF32 (10-bit mantissa), small = value - big (exact in fp32).  K is zero-padded.
__global__ void k_pack_weights(const float *__restrict__ Wt, int K, int N, int nsplit, float *__restrict__ img, int n_chunks) {
  const int64_t total = (int64_t)n_chunks * N * kChunkK;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % kChunkK);
    const int n = (int)((i / kChunkK) % N);
    const int j = (int)(i / ((int64_t)kChunkK * N));
    const int k = j * kChunkK + kk;
    const float v = k < K ? Wt[(int64_t)k * N + n] : 0.f;
    const float big = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const int unit = kk >> 2, e = kk & 3;
    const int64_t pos = (int64_t)n * 32 + ((unit ^ (n & 7)) << 2) + e;
    const int64_t chunk_floats = (int64_t)N * 32 * (nsplit == 3 ? 2 : 1);
    img[j * chunk_floats + pos] = big;
    if (nsplit == 3) img[j * chunk_floats + (int64_t)N * 32 + pos] = v - big;
  }
}

__global__ void k_bias_act(float *__restrict__ Y, int64_t ldY, int M, int h_host, const int32_t *h_dev,
                           const float *__restrict__ bias, int act) {
  const int H = h_dev ? min(*h_dev, h_host) : h_host;
  const int64_t total = (int64_t)H * M;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t h = i / M;
    const int m = (int)(i - h * M);
    float *y = Y + h * ldY + m;
    *y = act_apply(*y + (bias ? __ldg(bias + m) : 0.f), act);
  }
}

}  // namespace
}  // namespace efgh

using namespace efgh;

extern "C" int efgh_bcl_bias_act(float *Y, int64_t ldY, int M, int64_t h, const int32_t *h_dev, const float *bias, int act,
                                 void *stream) {
  EFGH_REQUIRE(M > 0 && h >= 0 && h < (1ll << 30) && ldY >= M, "efgh_bcl_bias_act: bad sizes");
  if (h == 0) return EFGH_OK;
  EFGH_REQUIRE(Y, "efgh_bcl_bias_act: null pointer");
  k_bias_act<<<grid_for(h * M, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(Y, ldY, M, (int)h, h_dev, bias, act);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

// Output widths up to this keep the running sum of an item's chain cuts in shared memory (128 rows x N floats).
constexpr int kOnChipMaxN = 128;

// Resource plan for output width N: TMEM = acc_stages*N + 4 teams x (32|64) A columns <= 512;
// shared memory = b_stages weight tiles + raw_slots x 64 KB of warp-private row slots + the epilogue tile.
static bool conv_tc_plan(int N, int nsplit, bool combine, ConvParams *p, size_t *smem_out) {
  const int a_cols = nsplit == 3 ? 64 : 32;
  const int room = 512 - kTeams * a_cols;                // TMEM columns left for accumulators
  int nacc = 1, acc_stages = 1;
  // two accumulator stages first (epilogue of item i overlaps the MMAs of item i+1), then spare columns go to
  // separate accumulators for the 3xTF32 correction products (shorter chains, smaller rounding error)
  if (nsplit == 3 && 6 * N <= room) { nacc = 3; acc_stages = 2; }
  else if (nsplit == 3 && 4 * N <= room) { nacc = 2; acc_stages = 2; }
  else if (2 * N <= room) { nacc = 1; acc_stages = 2; }
  if (acc_stages * nacc * N > room) return false;
  const size_t b_bytes = (size_t)N * 128 * (nsplit == 3 ? 2 : 1);
  const size_t epi_bytes = (size_t)4 * 32 * 4 * (combine ? N + 4 : kEpiRowFloats);
  const size_t fixed = 1024 + 256 + (kMaxGroups + 4) * 4 + kProducerWarps * 256 + 256 /* cut table */ + epi_bytes + kBiasFloats * 4;
  const size_t budget = 226 * 1024 - fixed;
  int b_stages = b_bytes >= 32 * 1024 ? 2 : 4;
  if ((size_t)b_stages * b_bytes + kRawSlotBytes > budget) return false;
  int raw = (int)((budget - (size_t)b_stages * b_bytes) / kRawSlotBytes);
  if (raw > kMaxRaw) raw = kMaxRaw;
  if (p) { p->b_stages = b_stages; p->raw_slots = raw; p->acc_stages = acc_stages; p->nacc = nacc; }
  if (smem_out) *smem_out = (size_t)b_stages * b_bytes + (size_t)raw * kRawSlotBytes + fixed;
  return true;
}

extern "C" int efgh_bcl_conv_tc_supported(int C, int F, int M, int nsplit) {
  if (!(nsplit == 1 || nsplit == 3)) return 0;
  if (C < 32 || C > 512 || C % 4 != 0 || M < 32 || M > 256 || M % 32 != 0 || F < 1) return 0;
  return conv_tc_plan(M, nsplit, M <= kOnChipMaxN, nullptr, nullptr) ? 1 : 0;
}

extern "C" size_t efgh_bcl_packed_weight_bytes(int K, int M, int nsplit) {
  const size_t chunks = (size_t)(K + kChunkK - 1) / kChunkK;
  return chunks * (size_t)M * 128 * (nsplit == 3 ? 2 : 1);
}

extern "C" int efgh_bcl_pack_weights(const float *Wt, int K, int M, int nsplit, float *Wimg, void *stream) {
  EFGH_REQUIRE(Wt && Wimg && K > 0 && M > 0 && (nsplit == 1 || nsplit == 3), "efgh_bcl_pack_weights: bad arguments");
  const int n_chunks = (K + kChunkK - 1) / kChunkK;
  k_pack_weights<<<grid_for((int64_t)n_chunks * M * kChunkK, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(Wt, K, M, nsplit, Wimg,
                                                                                                               n_chunks);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

// Number of partial sums per output that the kernel adds IN GLOBAL MEMORY for a contraction of length K and output
// width M.  The tensor core accumulates in TMEM with round-toward-zero, so its error grows linearly with the length
// of an accumulation chain; chains are cut every kGroupChunks x 32 terms.  For M <= 128 the cuts are summed on chip
// (returns 1: Y is written once, bias and activation applied); for wider outputs they are K groups added in L2
// (red.add.f32 into a zero-filled Y; the same split keeps all SMs busy on small lattices).
constexpr int kGroupChunks = 8;
extern "C" int efgh_bcl_conv_tc_groups(int K, int M) {
  if (M <= kOnChipMaxN) return 1;
  const int chunks = (K + kChunkK - 1) / kChunkK;
  return (chunks + kGroupChunks - 1) / kGroupChunks;
}

extern "C" int efgh_bcl_conv_tc(const float *X, int64_t ldX, int C, const float *in_bias, int in_act, const void *nbr,
                                int idx_bits, int64_t nbr_ld, int F, int64_t h, const int32_t *h_dev, const float *Wimg,
                                const float *bias, int M, int act, float *Y, int64_t ldY, int nsplit, int accumulate,
                                void *stream) {
  if (!nbr) F = 1;
  EFGH_REQUIRE(efgh_bcl_conv_tc_supported(C, F, M, nsplit), "efgh_bcl_conv_tc: unsupported shape C=%d F=%d M=%d nsplit=%d", C, F, M, nsplit);
  EFGH_REQUIRE(h >= 0 && h < (1ll << 30), "efgh_bcl_conv_tc: bad h");
  if (h == 0) return EFGH_OK;
  EFGH_REQUIRE(X && Wimg && Y, "efgh_bcl_conv_tc: null pointer");
  EFGH_REQUIRE(ldX % 4 == 0 && ldY % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(Wimg) & 15) == 0 && (reinterpret_cast<uintptr_t>(in_bias) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
               "efgh_bcl_conv_tc: X, Y, Wimg and in_bias must be 16-byte aligned with leading dimensions multiple of 4");
  EFGH_REQUIRE(idx_bits == 32 || idx_bits == 64, "efgh_bcl_conv_tc: idx_bits must be 32 or 64");
  ConvParams p;
  p.trace = g_conv_trace; p.dbg_flags = g_conv_flags;
  p.X = X; p.ldX = ldX; p.C = C; p.nbr = nbr; p.nbr_ld = nbr_ld; p.F = F;
  p.h_host = (int)h; p.h_dev = h_dev; p.Wimg = Wimg; p.bias = bias; p.N = M; p.act = act; p.Y = Y; p.ldY = ldY;
  p.in_bias = in_bias; p.in_act = in_act; p.accumulate = accumulate;
  p.n_chunks = (F * C + kChunkK - 1) / kChunkK;
  p.n_groups = efgh_bcl_conv_tc_groups(F * C, M);
  p.cut_chunks = M <= kOnChipMaxN ? kGroupChunks : 0;
  if ((g_conv_flags >> 8) & 63) {                            // timing studies: chunks per chain
    if (p.cut_chunks) p.cut_chunks = (g_conv_flags >> 8) & 63;
    else p.n_groups = (p.n_chunks + ((g_conv_flags >> 8) & 63) - 1) / ((g_conv_flags >> 8) & 63);
  }
  p.magic_c = (uint32_t)(((1ull << 32) + (uint64_t)C - 1) / (uint64_t)C);
  EFGH_REQUIRE(p.n_groups <= kMaxGroups, "efgh_bcl_conv_tc: K=%d too long (%d K groups, at most %d)", F * C, p.n_groups, kMaxGroups);
  EFGH_REQUIRE(accumulate || p.n_groups == 1,
               "efgh_bcl_conv_tc: K=%d needs %d partial sums; call with accumulate=1 on a zero-filled Y", F * C, p.n_groups);
  EFGH_REQUIRE(ldX < (1ll << 30), "efgh_bcl_conv_tc: ldX too large");   // row pitch in bytes fits 32 bits; mad.wide.u32 gives the 64-bit offset
  const bool combine = p.cut_chunks > 0 && p.n_chunks > p.cut_chunks;     // more than one accumulation chain per item
  EFGH_REQUIRE(!combine || (p.n_chunks + p.cut_chunks - 1) / p.cut_chunks <= 60, "efgh_bcl_conv_tc: K=%d too long for on-chip chain cuts", F * C);
  size_t smem = 0;
  conv_tc_plan(M, nsplit, combine, &p, &smem);
  if ((g_conv_flags >> 4) & 15) p.raw_slots = min(p.raw_slots, (g_conv_flags >> 4) & 15);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t items = ((h + kTileM - 1) / kTileM) * p.n_groups;
  const int grid = (int)(items < sm_count() ? items : sm_count());
#define EFGH_LAUNCH_TC(IDX, NS, CB)                                                                               \
  do {                                                                                                            \
    auto kern = k_conv_tc<IDX, NS, CB>;                                                                           \
    EFGH_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<grid, kThreads, smem, s>>>(p);                                                                         \
  } while (0)
#define EFGH_LAUNCH_TC2(IDX, NS) do { if (combine) EFGH_LAUNCH_TC(IDX, NS, true); else EFGH_LAUNCH_TC(IDX, NS, false); } while (0)
  if (idx_bits == 32) {
    if (nsplit == 3) EFGH_LAUNCH_TC2(int32_t, 3); else EFGH_LAUNCH_TC2(int32_t, 1);
  } else {
    if (nsplit == 3) EFGH_LAUNCH_TC2(int64_t, 3); else EFGH_LAUNCH_TC2(int64_t, 1);
  }
#undef EFGH_LAUNCH_TC2
#undef EFGH_LAUNCH_TC
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}
