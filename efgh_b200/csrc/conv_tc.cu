// Lattice convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   Y[h, m] = act(bias[m] + sum_{f,c} X[nbr[f,h]+1, c] * scale[row] * W[(f,c), m])
//
// reference nets/bilateralNN.py:240-244: advanced-index gather that materialises (1, C, F, H), then a cuDNN
// (F,1) convolution.  Here the gather IS the A-operand loader of a warp-specialised GEMM:
//
//   warps 0-7   gather producers, two teams of 128 (thread = tile row): cp.async 16-byte pieces of neighbour
//               rows of the (already normalised) vertex-major splat matrix from L2 straight into the
//               128-byte-swizzled K-major A tile (the UMMA canonical layout), several chunks in flight; then
//               split each landed value in place into a TF32-exact "big" part and the fp32 remainder "small"
//   warp  8     MMA issuer: one thread issues tcgen05.mma kind::tf32, M=128 x N x K=8, accumulators in TMEM.
//               3xTF32: D += A_small*B_big + A_big*B_small + A_big*B_big  (fp32-equivalent accuracy; the
//               dropped small*small term is 2^-22 relative), or a single TF32 pass when nsplit == 1
//   warp  9     weight loader: one thread issues cp.async.bulk (TMA engine, UBLKCP) per K chunk; the weights
//               were packed once (k_pack_weights) into the exact swizzled shared-memory image, big | small
//   warps 10-13 epilogue: tcgen05.ld TMEM -> registers, bias + activation, row-contiguous 16-byte stores
//
// Two TMEM accumulator stages (2 x N <= 512 columns) let the epilogue of tile t overlap the mainloop of
// tile t+1; a ring of shared-memory stages (full/empty mbarriers) decouples gather, weight load and MMA.
// Persistent CTAs, one per SM, stride over 128-vertex tiles; the vertex count is read from device memory.
#include "common.cuh"

static unsigned long long *g_conv_trace = nullptr;
// Debug hook (not part of the public header): device buffer of 5 roles x 512 x 2 u64 that CTA 0 fills with
// (event, globaltimer) pairs, or NULL to disable.
extern "C" void efgh_debug_set_conv_trace(unsigned long long *buf) { g_conv_trace = buf; }

namespace efgh {
namespace {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                  // floats per K chunk = one 128-byte swizzle row
constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kMmaWarp = 8, kTmaWarp = 9, kEpiWarp0 = 10;
constexpr int kThreads = (kEpiWarp0 + 4) * 32;  // 448
constexpr int kMaxStages = 6;
constexpr int kABytes = kTileM * 128;        // one A tile (128 rows x 128 B)

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.1f * v;
  return v;
}

// Optional timeline trace (debug): role r of CTA 0 appends (event, globaltimer) pairs to trace[r*kTraceCap...]
constexpr int kTraceCap = 512;
__device__ __forceinline__ void trace_ev(unsigned long long *trace, int role, int &n, int ev) {
  if (trace && blockIdx.x == 0 && n < kTraceCap) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(size_t)role * kTraceCap * 2 + 2 * n] = (unsigned long long)ev;
    trace[(size_t)role * kTraceCap * 2 + 2 * n + 1] = t;
    ++n;
  }
}

struct ConvParams {
  unsigned long long *trace;
  const float *X; int64_t ldX; int C;
  const float *in_bias; int in_act;     // optional input transform x = act(x + in_bias[c]) (deferred epilogue of a split-K producer)
  const void *nbr; int64_t nbr_ld; int F;
  int h_host; const int32_t *h_dev;
  const float *Wimg; const float *bias; int N; int act;
  float *Y; int64_t ldY;
  int n_chunks; int n_groups; int stages; int tmem_cols;
  int accumulate;                       // 1: red.add raw partial sums into pre-zeroed Y (bias/act deferred); 0: store act(bias + acc)
};

// K chunks [begin, end) of group g when n_chunks are dealt as evenly as possible to n_groups
__device__ __forceinline__ void group_range(int n_chunks, int n_groups, int g, int &begin, int &end) {
  const int base = n_chunks / n_groups, rem = n_chunks % n_groups;
  begin = g * base + min(g, rem);
  end = begin + base + (g < rem ? 1 : 0);
}

template <typename IdxT, int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A_big | A_small? | B_big | B_small?)] [barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int N = p.N;
  const uint32_t a_bytes = kABytes * (NSPLIT == 3 ? 2 : 1);
  const uint32_t b_bytes = (uint32_t)N * 128u * (NSPLIT == 3 ? 2 : 1);
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kMaxStages,
                 bar_acc_full = bar_empty + 8 * kMaxStages, bar_acc_empty = bar_acc_full + 16;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.h_dev ? min(*p.h_dev, p.h_host) : p.h_host;
  const int n_tiles = (H + kTileM - 1) / kTileM;
  const int n_items = n_tiles * p.n_groups;
  const int K = p.F * p.C;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 128 + 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, 128);
    }
    fence_barrier_init();
  }
  if (warp == kTmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < kProducerWarps) {
    // ===================== gather producers =====================
    // Two teams of 128 threads (thread = tile row) alternate over this CTA's work items, so one team's
    // pipeline drain / neighbour-table load overlaps the other team's copies.  Per K chunk a thread
    //   1. cp.async's (LDGSTS, zero-fill for absent neighbours) the 8 x 16-byte pieces of its row straight
    //      from the vertex-major matrix into the swizzled A tile - up to `depth` chunks in flight, no
    //      registers held across the L2 latency,
    //   2. once its own copies have landed, re-reads them, applies the optional deferred bias + activation,
    //      splits every value into the TF32-exact "big" part (written back in place) and the fp32 remainder
    //      "small" (second tile), fences the generic->async proxy and arrives on the stage's full barrier.
    const int team = warp >> 2;
    const int r = threadIdx.x & (kTileM - 1);
    const uint32_t swz = (uint32_t)(r & 7);
    const int depth = min(4, p.stages - 1);
    uint32_t it = 0;       // global chunk sequence number of this CTA (both teams track all items)
    int k_item = 0, ntrace = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k_item) {
      const int tile = item / p.n_groups, grp = item - tile * p.n_groups;
      int j_begin, j_end;
      group_range(p.n_chunks, p.n_groups, grp, j_begin, j_end);
      if ((k_item & 1) != team) { it += (uint32_t)(j_end - j_begin); continue; }
      const int h = tile * kTileM + r;
      const bool has_next = item + (int)gridDim.x < n_items;
      // Neighbour rows of this thread's vertex for filter taps f, f+1, f+2 live in registers and are refilled
      // two taps ahead of use (coalesced 512-byte reads across the team), so no index table / barrier is needed.
      int f_cur = (j_begin * kChunkK) / p.C;
      int c_cur = j_begin * kChunkK - f_cur * p.C;
      auto fetch_row = [&](int f) -> int {
        if (h >= H || f >= p.F) return -1;
        if (!p.nbr) return h;
        const int row = load_idx<IdxT>(p.nbr, f * p.nbr_ld + h) + 1;
        return row == 0 ? -1 : row;                                          // sink row: all zeros
      };
      int row0 = fetch_row(f_cur), row1 = fetch_row(f_cur + 1), row2 = fetch_row(f_cur + 2);
      // Stage slots are claimed in global chunk order: this team may start claiming only after the other team
      // has claimed every slot of the previous item (mbarrier parity cannot tell phases two apart).
      if (k_item > 0) asm volatile("bar.sync %0, 256;" ::"r"(3 + (team ^ 1)) : "memory");
      bool handed_off = false;

      auto issue = [&](int j, uint32_t seq) {
        const uint32_t s = seq % p.stages, ph = (seq / p.stages) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t a_big = smem_base + s * stage_bytes + (uint32_t)r * 128u;
        int k = j * kChunkK;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int row = k < K ? row0 : -1;
          const float *src = row >= 0 ? p.X + (int64_t)row * p.ldX + c_cur : p.X;
          const uint32_t nbytes = row >= 0 ? 16u : 0u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_big + (((uint32_t)u ^ swz) << 4)), "l"(src), "r"(nbytes)
                       : "memory");
          k += 4; c_cur += 4;
          if (c_cur >= p.C) {
            c_cur = 0; ++f_cur;
            row0 = row1; row1 = row2; row2 = fetch_row(f_cur + 2);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      auto convert = [&](int j, uint32_t seq) {
        const uint32_t s = seq % p.stages;
        const uint32_t a_big = smem_base + s * stage_bytes + (uint32_t)r * 128u, a_small = a_big + kABytes;
        int c = (j * kChunkK) % p.C;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t off = ((uint32_t)u ^ swz) << 4;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a_big + off) : "memory");
          if (p.in_bias) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(p.in_bias + c));
            v.x = act_apply(v.x + b.x, p.in_act); v.y = act_apply(v.y + b.y, p.in_act);
            v.z = act_apply(v.z + b.z, p.in_act); v.w = act_apply(v.w + b.w, p.in_act);
          }
          float4 big;
          big.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          big.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          big.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          big.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_big + off), "f"(big.x), "f"(big.y), "f"(big.z), "f"(big.w) : "memory");
          if (NSPLIT == 3) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_small + off), "f"(v.x - big.x), "f"(v.y - big.y),
                         "f"(v.z - big.z), "f"(v.w - big.w)
                         : "memory");
          }
          c += 4;
          if (c >= p.C) c = 0;
        }
        fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core (async proxy)
        mbar_arrive(bar_full + 8 * s);
      };

      int ji = j_begin, jc = j_begin;
      const bool tracer = (threadIdx.x & 127) == 0;
      if (tracer) trace_ev(p.trace, team, ntrace, 100 + k_item);
      while (jc < j_end) {
        while (ji < j_end && ji - jc < depth) {
          issue(ji, it + (uint32_t)(ji - j_begin)); ++ji;
          if (tracer) trace_ev(p.trace, team, ntrace, 1);
        }
        if (ji == j_end && !handed_off) {
          handed_off = true;
          if (has_next) asm volatile("bar.arrive %0, 256;" ::"r"(3 + team) : "memory");
        }
        const int pending = ji - jc - 1;         // younger groups that may still be in flight
        if (pending >= 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
        else if (pending == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (tracer) trace_ev(p.trace, team, ntrace, 2);
        convert(jc, it + (uint32_t)(jc - j_begin));
        if (tracer) trace_ev(p.trace, team, ntrace, 3);
        ++jc;
      }
      it += (uint32_t)(j_end - j_begin);
    }
  } else if (warp == kTmaWarp) {
    // ===================== weight loader (TMA bulk copies) =====================
    if (lane == 0) {
      uint32_t it = 0;
      int ntrace = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int j_begin, j_end;
        group_range(p.n_chunks, p.n_groups, item % p.n_groups, j_begin, j_end);
        for (int j = j_begin; j < j_end; ++j, ++it) {
          const uint32_t s = it % p.stages, ph = (it / p.stages) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          trace_ev(p.trace, 2, ntrace, 1);
          mbar_arrive_expect_tx(bar_full + 8 * s, b_bytes);
          bulk_g2s(smem_base + s * stage_bytes + a_bytes, reinterpret_cast<const uint8_t *>(p.Wimg) + (size_t)j * b_bytes, b_bytes,
                   bar_full + 8 * s);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(N);
      uint32_t it = 0, tcount = 0;
      int ntrace = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
        int j_begin, j_end;
        group_range(p.n_chunks, p.n_groups, item % p.n_groups, j_begin, j_end);
        const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
        trace_ev(p.trace, 3, ntrace, 100 + (int)tcount);
        mbar_wait(bar_acc_empty + 8 * as, aph ^ 1);      // epilogue drained this accumulator stage
        tc_fence_after();
        trace_ev(p.trace, 3, ntrace, 1);
        const uint32_t tmem_d = tmem_base + as * (uint32_t)N;
        for (int j = j_begin; j < j_end; ++j, ++it) {
          const uint32_t s = it % p.stages, ph = (it / p.stages) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          trace_ev(p.trace, 3, ntrace, 2);
          const uint32_t a_big = smem_base + s * stage_bytes, a_small = a_big + kABytes;
          const uint32_t b_big = a_big + a_bytes, b_small = b_big + (uint32_t)N * 128u;
          const uint64_t dab = make_desc(a_big), dbb = make_desc(b_big);
          uint32_t acc = j > j_begin;
          if (NSPLIT == 3) {
            const uint64_t das = make_desc(a_small), dbs = make_desc(b_small);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) { umma_tf32(tmem_d, das + 2 * k4, dbb + 2 * k4, idesc, acc); acc = 1; }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) umma_tf32(tmem_d, dab + 2 * k4, dbs + 2 * k4, idesc, 1);
          }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) { umma_tf32(tmem_d, dab + 2 * k4, dbb + 2 * k4, idesc, acc); acc = 1; }
          umma_commit(bar_empty + 8 * s);              // frees the smem stage when these MMAs retire
        }
        umma_commit(bar_acc_full + 8 * as);            // accumulator complete -> epilogue
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    uint32_t tcount = 0;
    int ntrace = 0;
    const bool tracer = threadIdx.x == kEpiWarp0 * 32;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
      const int tile = item / p.n_groups;
      const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
      if (tracer) trace_ev(p.trace, 4, ntrace, 100 + (int)tcount);
      mbar_wait(bar_acc_full + 8 * as, aph);
      tc_fence_after();
      if (tracer) trace_ev(p.trace, 4, ntrace, 1);
      const int h = tile * kTileM + q * 32 + lane;
      float *yrow = p.Y + (int64_t)h * p.ldY;
      for (int cb = 0; cb < N; cb += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)N + cb, v);
        if (h < H && p.accumulate) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            atomicAdd(reinterpret_cast<float4 *>(yrow + cb + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        } else if (h < H) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o;
            o.x = act_apply(v[i + 0] + (p.bias ? __ldg(p.bias + cb + i + 0) : 0.f), p.act);
            o.y = act_apply(v[i + 1] + (p.bias ? __ldg(p.bias + cb + i + 1) : 0.f), p.act);
            o.z = act_apply(v[i + 2] + (p.bias ? __ldg(p.bias + cb + i + 2) : 0.f), p.act);
            o.w = act_apply(v[i + 3] + (p.bias ? __ldg(p.bias + cb + i + 3) : 0.f), p.act);
            *reinterpret_cast<float4 *>(yrow + cb + i) = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty + 8 * as);
      if (tracer) trace_ev(p.trace, 4, ntrace, 2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
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

static int conv_tc_stages(int N, int F, int nsplit, size_t *smem_out) {
  const size_t stage = (size_t)(kABytes + N * 128) * (nsplit == 3 ? 2 : 1);
  const size_t fixed = 8 * (2 * kMaxStages + 4) + 16 + 1024;
  (void)F;
  int stages = (int)((220 * 1024 - fixed) / stage);
  if (stages > kMaxStages) stages = kMaxStages;
  if (smem_out) *smem_out = fixed + (size_t)stages * stage;
  return stages;
}

extern "C" int efgh_bcl_conv_tc_supported(int C, int F, int M, int nsplit) {
  if (!(nsplit == 1 || nsplit == 3)) return 0;
  if (C <= 0 || C % 4 != 0 || M < 16 || M > 256 || M % 32 != 0 || F < 1) return 0;
  return conv_tc_stages(M, F, nsplit, nullptr) >= 2;
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

// Number of K groups (= partial sums per output) the kernel will use for a contraction of length K.  The
// tensor core accumulates in TMEM with round-toward-zero, so its error grows linearly with the length of
// an accumulation chain; chains are cut every kGroupChunks x 32 terms and the partial sums are added in L2
// (red.add.f32, round-to-nearest).  The same cut is the split-K that keeps all SMs busy on small lattices.
constexpr int kGroupChunks = 8;
extern "C" int efgh_bcl_conv_tc_groups(int K) {
  const int chunks = (K + kChunkK - 1) / kChunkK;
  return (chunks + kGroupChunks - 1) / kGroupChunks;
}

extern "C" int efgh_bcl_conv_tc(const float *X, int64_t ldX, int C, const float *in_bias,
                                int in_act, const void *nbr, int idx_bits, int64_t nbr_ld, int F, int64_t h,
                                const int32_t *h_dev, const float *Wimg, const float *bias, int M, int act, float *Y,
                                int64_t ldY, int nsplit, int accumulate, void *stream) {
  if (!nbr) F = 1;
  EFGH_REQUIRE(efgh_bcl_conv_tc_supported(C, F, M, nsplit), "efgh_bcl_conv_tc: unsupported shape C=%d F=%d M=%d nsplit=%d", C, F, M, nsplit);
  EFGH_REQUIRE(h >= 0 && h < (1ll << 30), "efgh_bcl_conv_tc: bad h");
  if (h == 0) return EFGH_OK;
  EFGH_REQUIRE(X && Wimg && Y, "efgh_bcl_conv_tc: null pointer");
  EFGH_REQUIRE(ldX % 4 == 0 && ldY % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(Wimg) & 15) == 0 && (reinterpret_cast<uintptr_t>(in_bias) & 15) == 0,
               "efgh_bcl_conv_tc: X, Y and Wimg must be 16-byte aligned with leading dimensions multiple of 4");
  EFGH_REQUIRE(idx_bits == 32 || idx_bits == 64, "efgh_bcl_conv_tc: idx_bits must be 32 or 64");
  ConvParams p;
  p.trace = g_conv_trace;
  p.X = X; p.ldX = ldX; p.C = C; p.nbr = nbr; p.nbr_ld = nbr_ld; p.F = F;
  p.h_host = (int)h; p.h_dev = h_dev; p.Wimg = Wimg; p.bias = bias; p.N = M; p.act = act; p.Y = Y; p.ldY = ldY;
  p.in_bias = in_bias; p.in_act = in_act; p.accumulate = accumulate;
  p.n_chunks = (F * C + kChunkK - 1) / kChunkK;
  p.n_groups = efgh_bcl_conv_tc_groups(F * C);
  EFGH_REQUIRE(accumulate || p.n_groups == 1,
               "efgh_bcl_conv_tc: K=%d needs %d partial sums; call with accumulate=1 on a zero-filled Y", F * C, p.n_groups);
  size_t smem = 0;
  p.stages = conv_tc_stages(M, F, nsplit, &smem);
  int cols = 32;
  while (cols < 2 * M) cols <<= 1;
  p.tmem_cols = cols;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t items = ((h + kTileM - 1) / kTileM) * p.n_groups;
  const int grid = (int)(items < sm_count() ? items : sm_count());
#define EFGH_LAUNCH_TC(IDX, NS)                                                                                   \
  do {                                                                                                            \
    auto kern = k_conv_tc<IDX, NS>;                                                                               \
    EFGH_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<grid, kThreads, smem, s>>>(p);                                                                         \
  } while (0)
  if (idx_bits == 32) {
    if (nsplit == 3) EFGH_LAUNCH_TC(int32_t, 3); else EFGH_LAUNCH_TC(int32_t, 1);
  } else {
    if (nsplit == 3) EFGH_LAUNCH_TC(int64_t, 3); else EFGH_LAUNCH_TC(int64_t, 1);
  }
#undef EFGH_LAUNCH_TC
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}
