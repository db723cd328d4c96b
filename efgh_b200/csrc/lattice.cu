// Permutohedral lattice build on the GPU (d = 3).
//
// Replaces, per level, reference nets/generate_data.py:128-179 + nets/transforms.py:125-184:
//   k_clear    reset the hash table / scan state, latch the device-side point count
//   k_points   elevate -> round -> rank -> barycentric -> 4 simplex keys per point
//              (generate_data.py:56-112), key box (:135-136), hash insert keeping per key the
//              SMALLEST stream position p = 4*point + remainder (the order transforms.py:152-166 walks)
//   k_assign   single-pass decoupled-look-back scan over "p is the first occurrence of its key":
//              rank among first occurrences == the reference's insertion index hash_cnt1
//   k_vertices lattice offsets per point, blur neighbours per vertex (transforms.py:168-180, including
//              key2int's mixed-radix aliasing for keys outside the key box), next-level coordinates
//              (generate_data.py:175-178)
//
// Data layout in HBM (all sized by the caller through efgh_lattice_workspace_bytes):
//   table  {u64 key, i32 first_pos, i32 index}[T]  16-byte entries: packed key (3 coordinates x 21 bits; the 4th
//                   is minus their sum), smallest stream position that inserted it, vertex index (valid after
//                   k_assign) - one 16-byte load per probe returns everything a later pass needs
//   slots  int4[n]  table slot of each of the point's 4 keys (so later passes never re-probe)
//   vkeys  u64[4n]  packed key of vertex h, insertion order
//   tiles  u64[..]  look-back status words
// T = power of two >= 8 * n (load factor <= 0.5 even when every key is distinct), so for one 131k-point
// scan the whole table (16 MB) stays resident in the 126 MB L2 and the random probes never reach HBM.
//
// Float semantics are part of the contract (keys must be bit-exact): every operation below is an
// explicit round-to-nearest intrinsic so that nvcc cannot contract or reassociate it; the chain
// reproduces MKL sgemm's k-ascending FMA order that the reference's torch.matmul resolves to.
#include <stdarg.h>

#include "common.cuh"

namespace efgh {

namespace {

constexpr int kBias = 1 << 20;
constexpr unsigned long long kEmpty = ~0ull;
constexpr int kPointThreads = 256;
constexpr int kTile = 1024;  // points per scan tile (4 keys each)

struct __align__(16) Entry {
  unsigned long long key;
  int first_pos;
  int index;
};

// Several scans in one launch sequence ("ragged batch").  The point streams of B scans are
// concatenated (scan b = points [pt_start[b], pt_start[b+1])); every scan keeps its OWN hash table (region b of
// `tstride` entries), its own key box and its own insertion order, so per scan the result is what the single-scan
// path produces - only the vertex indices are global: vertex_start[b] + local index.  The kernels downstream (splat,
// convolution) then treat the batch as one big lattice.  info (int32, device):
//   [0, B]                vertex_start (written by k_assign; it is the next level's pt_start)
//   [tm_off, tm_off + B)  per-scan table mask
//   [box_off + 8 b ...)   per-scan key box: min[4], max[4]   (one 32-byte sector per scan)
constexpr int kMaxBatch = 64;
struct Batch {
  const int32_t *pt_start;   // nullptr = single scan
  int32_t *info;
  int B, tm_off, box_off;
  long long tstride;
  // optional vertex -> contributions lists (CSR) for the gather-form splat: voff[h] .. voff[h+1] index `contrib`,
  // whose entries are stream positions 4 * point + remainder
  int32_t *voff;             // [0, h_cap]: offsets; [h_cap + 1]: number of heavy vertices; then their indices
  int32_t *contrib;
  int32_t *cursor;           // workspace, h_cap ints
  float *point_rows;         // optional (n, 8) point-major copy: el_minus_gr[0..3], barycentric[0..3]
  int h_cap;
};
// A vertex with more than kHeavy contributions is listed so that the gather-form splat can put a whole CTA on it instead
// of a few lanes.  At level 0 those are common: the ground rings next to the sensor put 50-100 points into one lattice
// cell (a 16-scan batch lists a few thousand vertices); an overflowing list makes single lanes walk lists of several
// hundred contributions serially - measured 4x on the whole level-0 splat.
constexpr int kHeavy = 64;
constexpr int kMaxHeavy = 16384;
inline int batch_tm_off(int B) { return (B + 1 + 7) & ~7; }
inline int batch_box_off(int B) { return batch_tm_off(B) + ((B + 7) & ~7); }

// scan that owns stream element i: largest b with start[b] <= i (start[] ascending, start[0] = 0)
__device__ __forceinline__ int find_scan(const int *start, int B, int i) {
  int lo = 0, hi = B;                                   // invariant: start[lo] <= i < start[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (start[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

struct Workspace {
  Entry *table;
  int4 *slots;
  unsigned long long *vkeys;
  unsigned long long *tiles;
  int32_t *cursor;
  int64_t table_cap;
  size_t bytes;
};

inline int64_t pow2_ceil(int64_t v) {
  int64_t p = 1024;
  while (p < v) p <<= 1;
  return p;
}

inline size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

// n_cap: points of the whole launch; B scans with table_entries hash-table entries each (single scan: 8 per point)
Workspace carve(void *base, int64_t n_cap, int B = 1, int64_t table_entries = -1) {
  Workspace w;
  if (n_cap < 1) n_cap = 1;
  w.table_cap = table_entries >= 1 ? pow2_ceil(table_entries) : pow2_ceil(8 * n_cap);   // per scan
  size_t off = 0;
  char *b = static_cast<char *>(base);
  w.table = reinterpret_cast<Entry *>(b + off); off = align_up(off + sizeof(Entry) * w.table_cap * B);
  w.slots = reinterpret_cast<int4 *>(b + off); off = align_up(off + sizeof(int4) * n_cap);
  w.vkeys = reinterpret_cast<unsigned long long *>(b + off); off = align_up(off + sizeof(unsigned long long) * 4 * n_cap);
  w.tiles = reinterpret_cast<unsigned long long *>(b + off); off = align_up(off + sizeof(unsigned long long) * ((n_cap + kTile - 1) / kTile + 1));
  w.cursor = reinterpret_cast<int32_t *>(b + off); off = align_up(off + sizeof(int32_t) * 4 * n_cap);
  w.bytes = off;
  return w;
}

__device__ __forceinline__ unsigned long long pack_key(int k0, int k1, int k2) {
  return ((unsigned long long)(unsigned)(k0 + kBias) << 42) | ((unsigned long long)(unsigned)(k1 + kBias) << 21) |
         (unsigned long long)(unsigned)(k2 + kBias);
}

__device__ __forceinline__ void unpack_key(unsigned long long p, int &k0, int &k1, int &k2) {
  k0 = (int)((p >> 42) & 0x1fffff) - kBias;
  k1 = (int)((p >> 21) & 0x1fffff) - kBias;
  k2 = (int)(p & 0x1fffff) - kBias;
}

// 32-bit mix of the packed key (murmur3 finaliser over the folded halves): a third of the instructions of a 64-bit
// finaliser - the lattice kernels execute it 4 times per point and 15 times per vertex.
__device__ __forceinline__ unsigned hash_key(unsigned long long k) {
  unsigned h = (unsigned)k ^ ((unsigned)(k >> 32) * 0x9e3779b1u);
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

// table size for n points: power of two >= 8n (>= 1024), never above the carved capacity
__device__ __forceinline__ int table_mask_for(int n, int64_t table_cap) {
  long long want = 8ll * n;
  long long t = 1024;
  while (t < want && t < table_cap) t <<= 1;
  return (int)(t - 1);
}

__global__ void k_clear(efgh_lattice_state *st, const int32_t *n_dev, int n_host, int64_t table_cap,
                        Entry *table, unsigned long long *tiles, int n_tiles_cap, Batch bt) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int4 empty = make_int4(-1, -1, 0x7fffffff, -1);   // key = ~0, first_pos = INT_MAX, index = -1
  int n, mask = 0;
  if (bt.pt_start) {
    n = min(max(bt.pt_start[bt.B], 0), n_host);
    for (int b = 0; b < bt.B; ++b) {
      const int mb = table_mask_for(bt.pt_start[b + 1] - bt.pt_start[b], table_cap);
      Entry *tb = table + (long long)b * bt.tstride;
      for (int64_t i = tid; i <= mb; i += stride) reinterpret_cast<int4 *>(tb)[i] = empty;
      if (tid == 0) bt.info[bt.tm_off + b] = mb;
      if (tid < 8) bt.info[bt.box_off + 8 * b + tid] = tid < 4 ? 0x7fffffff : -0x7fffffff - 1;
    }
  } else {
    n = n_dev ? min(max(*n_dev, 0), n_host) : n_host;
    mask = table_mask_for(n, table_cap);
    for (int64_t i = tid; i <= mask; i += stride) reinterpret_cast<int4 *>(table)[i] = empty;
  }
  int n_tiles = (n + kTile - 1) / kTile;
  for (int64_t i = tid; i < n_tiles && i < n_tiles_cap; i += stride) tiles[i] = 0ull;
  if (tid == 0) {
    if (bt.voff) bt.voff[bt.h_cap + 1] = 0;
    st->n = n;
    st->hash_cnt = 0;
    st->status = 0;
    st->table_mask = mask;
    for (int c = 0; c < 4; ++c) {
      st->key_min[c] = 0x7fffffff;
      st->key_max[c] = -0x7fffffff - 1;
    }
    st->tile_counter = 0;
  }
}

// Elevation matrix E (4x3) as float32 bit patterns (generate_data.py:15-20) and
// expected_std = 4*sqrt(2/3) (generate_data.py:19) rounded to float32.
#define EFGH_E_A __int_as_float(0x3f3504f3)   //  0.70710677  1/sqrt(2)
#define EFGH_E_B __int_as_float(0x3ed105eb)   //  0.40824828  1/sqrt(6)
#define EFGH_E_C __int_as_float(0x3e93cd3a)   //  0.28867513  1/sqrt(12)
#define EFGH_E_B2 __int_as_float(0xbf5105eb)  // -2/sqrt(6)
#define EFGH_E_C3 __int_as_float(0xbf5db3d7)  // -3/sqrt(12) (rounded product)
#define EFGH_STD __int_as_float(0x405105ec)   //  3.2659864

__device__ __forceinline__ int canonical(int i, int j) { return (j <= 3 - i) ? j : j - 4; }

// Minimum CTAs per SM (register caps).  These kernels are latency-bound on random table accesses: k_assign at 83
// registers ran ONE 512-thread CTA per SM and k_vertices at 78 registers three 256-thread CTAs; capping both at 64
// registers doubles k_assign's occupancy (measured: lattice stages -6 %); tighter caps spill and lose.
#ifndef EFGH_LB_POINTS
#define EFGH_LB_POINTS 4
#endif
__global__ void __launch_bounds__(kPointThreads, EFGH_LB_POINTS)
k_points(const float *__restrict__ pts, int64_t pts_ld, float scale, float *__restrict__ bary,
         float *__restrict__ elmgr, int64_t out_ld, efgh_lattice_state *st, Entry *table_all, int4 *slots, Batch bt) {
  const int n = st->n;
  unsigned mask = (unsigned)st->table_mask;
  Entry *table = table_all;
  int kmin[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
  int kmax[4] = {-0x7fffffff - 1, -0x7fffffff - 1, -0x7fffffff - 1, -0x7fffffff - 1};
  int bad = 0;
  __shared__ int s_start[kMaxBatch + 1];
  __shared__ int s_min[4][kPointThreads / 32], s_max[4][kPointThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool batched = bt.pt_start != nullptr;
  if (batched) {
    for (int b = threadIdx.x; b <= bt.B; b += blockDim.x) s_start[b] = bt.pt_start[b];
    __syncthreads();
  }

  const float E[4][3] = {{EFGH_E_A, EFGH_E_B, EFGH_E_C},
                         {-EFGH_E_A, EFGH_E_B, EFGH_E_C},
                         {0.0f, EFGH_E_B2, EFGH_E_C},
                         {0.0f, 0.0f, EFGH_E_C3}};

  // (the loop is CTA-uniform - every thread runs the same number of rounds - because the batched path ends each
  //  round with a CTA-wide key-box reduction)
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    int scan = 0;
    if (batched) {
      scan = find_scan(s_start, bt.B, min(i, n - 1));
      table = table_all + (long long)scan * bt.tstride;
      mask = (unsigned)bt.info[bt.tm_off + scan];
    }
    const long long slot_base = (long long)scan * bt.tstride;
    if (i < n) {
    float p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __fmul_rn(__ldg(pts + a * pts_ld + i), scale);  // generate_data.py:130
    float el[4], gr[4], d0[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = __fmul_rn(E[c][0], p[0]);                  // :67 k-ascending FMA chain (MKL sgemm)
      acc = __fmaf_rn(E[c][1], p[1], acc);
      acc = __fmaf_rn(E[c][2], p[2], acc);
      el[c] = __fmul_rn(acc, EFGH_STD);
      gr[c] = __fmul_rn(rintf(__fmul_rn(el[c], 0.25f)), 4.0f);  // :70 (x/4 == x*0.25 exactly)
      d0[c] = __fsub_rn(el[c], gr[c]);                       // :72
    }
    int rank[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {                            // :73-78 stable descending sort, inverted
      int r = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) r += (d0[j] > d0[c]) || (d0[j] == d0[c] && j < c);
      rank[c] = r;
    }
    const int rs = (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(gr[0], gr[1]), gr[2]), gr[3]), 0.25f);  // :81
    float b[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    float diff[4];
    int gi[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {                            // :83-93
      if (rs > 0 && rank[c] >= 4 - rs) { gr[c] = __fsub_rn(gr[c], 4.0f); rank[c] -= 4; }
      else if (rs < 0 && rank[c] < -rs) { gr[c] = __fadd_rn(gr[c], 4.0f); rank[c] += 4; }
      rank[c] += rs;
      diff[c] = __fsub_rn(el[c], gr[c]);                     // :96
      if (!(fabsf(gr[c]) < (float)(kBias - 8))) { bad = 1; gr[c] = 0.f; }
      gi[c] = (int)gr[c];                                    // :97
      rank[c] &= 3;                                          // only reachable for non-finite input
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {                            // :100 slot 3-rank gets +diff
#pragma unroll
      for (int s = 0; s < 4; ++s) if (3 - rank[c] == s) b[s] = __fadd_rn(b[s], diff[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {                            // :101 slot 4-rank gets -diff
#pragma unroll
      for (int s = 1; s < 5; ++s) if (4 - rank[c] == s) b[s] = __fsub_rn(b[s], diff[c]);
    }
#pragma unroll
    for (int s = 0; s < 5; ++s) b[s] = __fmul_rn(b[s], 0.25f);   // :102
    b[0] = __fadd_rn(b[0], __fadd_rn(1.0f, b[4]));               // :103
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      bary[c * out_ld + i] = b[c];
      elmgr[c * out_ld + i] = diff[c];
    }
    if (bt.point_rows) {
      float4 *pr = reinterpret_cast<float4 *>(bt.point_rows + (int64_t)i * 8);
      pr[0] = make_float4(diff[0], diff[1], diff[2], diff[3]);
      pr[1] = make_float4(b[0], b[1], b[2], b[3]);
    }

    // key box of the point's 4 keys (generate_data.py:135-136): over the remainders r = 0..3 coordinate c takes the values
    // greedy[c] + canonical[rank[c]][r] = greedy[c] + {-rank[c], ..., 3 - rank[c]}
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      kmin[c] = min(kmin[c], gi[c] - rank[c]);
      kmax[c] = max(kmax[c], gi[c] + 3 - rank[c]);
    }
    int s4[4];
    unsigned long long key4[4], cur4[4];
    unsigned h4[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {                            // :106 keys[c, n, r] = greedy[c] + canonical[rank[c], r]
      int k[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) k[c] = gi[c] + canonical(rank[c], r);
      key4[r] = pack_key(k[0], k[1], k[2]);
      h4[r] = hash_key(key4[r]) & mask;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) cur4[r] = table[h4[r]].key;   // four independent probes in flight (L2 latency bound)
#pragma unroll
    for (int r = 0; r < 4; ++r)                                // ... and four independent claims of empty slots
      if (cur4[r] == kEmpty) cur4[r] = atomicCAS(&table[h4[r]].key, kEmpty, key4[r]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const unsigned long long key = key4[r];
      unsigned h = h4[r];
      unsigned long long cur = cur4[r];                        // kEmpty here means "we just claimed it"
      unsigned probes = 0;
      while (cur != kEmpty && cur != key) {                    // collision: linear probing
        h = (h + 1) & mask;
        if (++probes > mask) { bad |= 4; break; }
        cur = table[h].key;
        if (cur == kEmpty) cur = atomicCAS(&table[h].key, kEmpty, key);
      }
      atomicMin(&table[h].first_pos, 4 * i + r);
      if (bt.voff) atomicAdd(&table[h].index, 1);              // index = (number of contributions) - 1 until k_assign
      s4[r] = (int)(slot_base + h);
    }
    slots[i] = make_int4(s4[0], s4[1], s4[2], s4[3]);
    }  // i < n

    if (batched) {
      // per-scan key box.  Almost every CTA round lies inside one scan: CTA reduction, 8 guarded atomics.  A round
      // that straddles a scan boundary (at most B - 1 of them per launch) falls back to guarded per-thread atomics.
      const int first = find_scan(s_start, bt.B, min(base, n - 1));
      const int last = find_scan(s_start, bt.B, min(base + (int)blockDim.x - 1, n - 1));
      if (first == last) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          int mn = warp_reduce_min(kmin[c]), mx = warp_reduce_max(kmax[c]);
          if (lane == 0) { s_min[c][wid] = mn; s_max[c][wid] = mx; }
        }
        __syncthreads();
        if (threadIdx.x < 4) {
          int c = threadIdx.x, mn = s_min[c][0], mx = s_max[c][0];
          for (int w = 1; w < kPointThreads / 32; ++w) { mn = min(mn, s_min[c][w]); mx = max(mx, s_max[c][w]); }
          int *box = bt.info + bt.box_off + 8 * first;
          if (mn <= mx) {
            if (mn < *(volatile int *)&box[c]) atomicMin(&box[c], mn);
            if (mx > *(volatile int *)&box[4 + c]) atomicMax(&box[4 + c], mx);
          }
        }
        __syncthreads();
      } else if (i < n) {
        int *box = bt.info + bt.box_off + 8 * scan;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (kmin[c] < *(volatile int *)&box[c]) atomicMin(&box[c], kmin[c]);
          if (kmax[c] > *(volatile int *)&box[4 + c]) atomicMax(&box[4 + c], kmax[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) { kmin[c] = 0x7fffffff; kmax[c] = -0x7fffffff - 1; }
    }
  }

  if (batched) {
    if (bad) atomicOr(&st->status, (bad & 1 ? EFGH_ST_KEY_RANGE : 0) | (bad & 4 ? EFGH_ST_TABLE_FULL : 0));
    return;
  }
  // key box: warp shuffle -> shared -> 8 global atomics per CTA
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int mn = warp_reduce_min(kmin[c]), mx = warp_reduce_max(kmax[c]);
    if (lane == 0) { s_min[c][wid] = mn; s_max[c][wid] = mx; }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    int c = threadIdx.x, mn = s_min[c][0], mx = s_max[c][0];
    for (int w = 1; w < kPointThreads / 32; ++w) { mn = min(mn, s_min[c][w]); mx = max(mx, s_max[c][w]); }
    // The 8 box words share one 32-byte sector, so every atomic on them serialises in one L2 slice; only CTAs that
    // actually widen the box issue one (the plain read may be stale, but the box only ever widens - a stale value
    // just costs a redundant atomic).
    if (mn <= mx) {
      if (mn < *(volatile int *)&st->key_min[c]) atomicMin(&st->key_min[c], mn);
      if (mx > *(volatile int *)&st->key_max[c]) atomicMax(&st->key_max[c], mx);
    }
  }
  if (bad) atomicOr(&st->status, (bad & 1 ? EFGH_ST_KEY_RANGE : 0) | (bad & 4 ? EFGH_ST_TABLE_FULL : 0));
}

// Single-pass scan (decoupled look-back) over first-occurrence flags; two points (8 keys) per thread, stream
// order preserved (thread t owns points 2t, 2t+1 of the tile).
// status word: bits 63..62 = 1 aggregate ready / 2 inclusive prefix ready, low 32 bits = value.
constexpr int kAssignThreads = kTile / 2;
#ifndef EFGH_LB_ASSIGN
#define EFGH_LB_ASSIGN 2
#endif
__global__ void __launch_bounds__(kAssignThreads, EFGH_LB_ASSIGN)
k_assign(efgh_lattice_state *st, const int4 *__restrict__ slots, Entry *table,
         unsigned long long *__restrict__ vkeys, unsigned long long *tiles, int h_cap, Batch bt) {
  __shared__ int s_tile;
  __shared__ unsigned long long s_prefix;
  __shared__ unsigned long long s_warp[kAssignThreads / 32];
  const int n = st->n;
  const int n_tiles = (n + kTile - 1) / kTile;
  if (threadIdx.x == 0) s_tile = atomicAdd(&st->tile_counter, 1);
  __syncthreads();
  const int tile = s_tile;
  if (tile >= n_tiles) return;

  const int i0 = tile * kTile + 2 * threadIdx.x;
  int sl[8];
  int fl[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int4 s = make_int4(0, 0, 0, 0);
    if (i0 + q < n) s = slots[i0 + q];
    sl[4 * q] = s.x; sl[4 * q + 1] = s.y; sl[4 * q + 2] = s.z; sl[4 * q + 3] = s.w;
  }
  unsigned long long kk[8];
  int cc[8];                                                   // contributions of the key (valid for first occurrences)
  int cnt = 0, ccnt = 0;
  {
    int4 ent[8];
#pragma unroll
    for (int e = 0; e < 8; ++e)                                                    // eight independent 16-byte L2 reads in flight
      ent[e] = (i0 + (e >> 2) < n) ? *reinterpret_cast<const int4 *>(table + sl[e]) : make_int4(0, 0, -1, 0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      fl[e] = ent[e].z == 4 * (i0 + (e >> 2)) + (e & 3);                           // first_pos == own stream position
      cnt += fl[e];
      cc[e] = ent[e].w + 1;
      ccnt += fl[e] ? cc[e] : 0;
      kk[e] = ((unsigned long long)(unsigned)ent[e].y << 32) | (unsigned)ent[e].x;
    }
  }

  // block exclusive scan of (cnt, ccnt), packed low / high 32 bits
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned long long mine = (unsigned long long)(unsigned)cnt | ((unsigned long long)(unsigned)ccnt << 32);
  unsigned long long inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    unsigned long long v = lane < kAssignThreads / 32 ? s_warp[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane < kAssignThreads / 32) s_warp[lane] = v;  // inclusive warp totals
  }
  __syncthreads();
  const unsigned long long block_total = s_warp[kAssignThreads / 32 - 1];
  const unsigned long long local = inc - mine + (wid ? s_warp[wid - 1] : 0ull);
  // look-back word: bits 63..62 state, bits 59..30 contributions, bits 29..0 vertices (both < 2^30)
  const unsigned long long block_word = (block_total & 0x3fffffffull) | ((block_total >> 32) << 30);

  if (wid == 0) {
    // warp-parallel look-back: lane l inspects tile (base - l); stop at the first inclusive prefix
    volatile unsigned long long *vt = tiles;
    unsigned long long prefix = 0;
    if (tile > 0) {
      if (lane == 0) vt[tile] = (1ull << 62) | block_word;
      int base = tile - 1;
      while (true) {
        const int j = base - lane;
        unsigned long long w = 2ull << 62;                 // virtual tile before tile 0: inclusive prefix 0
        if (j >= 0) { do { w = vt[j]; } while ((w >> 62) == 0); }
        const unsigned incl = __ballot_sync(0xffffffffu, (w >> 62) == 2);
        const int first = __ffs(incl) - 1;                  // nearest tile that already knows its inclusive prefix
        unsigned long long contrib = (first < 0 || lane <= first) ? (w & 0x0fffffffffffffffull) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        prefix += contrib;                                  // the two 30-bit fields cannot carry into each other
        if (first >= 0) break;
        base -= 32;
      }
    }
    if (lane == 0) {
      __threadfence();
      vt[tile] = (2ull << 62) | (prefix + block_word);
      s_prefix = prefix;
      if (tile == n_tiles - 1) {
        const unsigned long long tw = prefix + block_word;
        const int total = (int)(tw & 0x3fffffffull);
        st->hash_cnt = total;
        if (bt.pt_start) bt.info[bt.B] = total;             // vertex_start[B]
        if (bt.voff && total <= h_cap) bt.voff[total] = (int)(tw >> 30);
        if (total > h_cap) atomicOr(&st->status, EFGH_ST_VERTEX_CAP);
      }
    }
  }
  __syncthreads();
  int idx = (int)(s_prefix & 0x3fffffffull) + (int)(unsigned)(local & 0xffffffffull);
  int coff = (int)(s_prefix >> 30) + (int)(local >> 32);
  if (bt.pt_start) {
    // vertex_start[b] = number of vertices created before scan b's first point (scans are never empty)
    const int c0 = fl[0] + fl[1] + fl[2] + fl[3];
    for (int b = 0; b < bt.B; ++b) {
      const int ps = bt.pt_start[b];
      if (ps == i0) bt.info[b] = idx;
      else if (ps == i0 + 1) bt.info[b] = idx + c0;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (fl[e]) {
      table[sl[e]].index = idx;
      if (idx < h_cap) {
        vkeys[idx] = kk[e];
        if (bt.voff) {
          bt.voff[idx] = coff; bt.cursor[idx] = coff; coff += cc[e];
          if (cc[e] > kHeavy) {
            const int slot = atomicAdd(&bt.voff[bt.h_cap + 1], 1);
            if (slot < kMaxHeavy) bt.voff[bt.h_cap + 2 + slot] = idx;
          }
        }
      }
      ++idx;
    }
}

#ifndef EFGH_LB_VERTICES
#define EFGH_LB_VERTICES 4
#endif
__global__ void __launch_bounds__(256, EFGH_LB_VERTICES)
k_vertices(efgh_lattice_state *__restrict__ st, int n_cap, int h_cap, const int4 *__restrict__ slots,
           const Entry *__restrict__ table_all, const unsigned long long *__restrict__ vkeys, int64_t *__restrict__ loff, int32_t *__restrict__ loff32,
           int64_t off_ld, const int32_t *__restrict__ foffs, int F, int64_t *__restrict__ nbr,
           int32_t *__restrict__ nbr32, int64_t nbr_ld, float *__restrict__ next_pts, int64_t next_ld,
           float next_divisor, Batch bt) {
  const int n = min(st->n, n_cap);
  const int H = min(st->hash_cnt, h_cap);
  unsigned mask = (unsigned)st->table_mask;
  const int stride = gridDim.x * blockDim.x;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ int s_start[kMaxBatch + 1];
  const bool batched = bt.pt_start != nullptr;
  if (batched) {
    for (int b = threadIdx.x; b <= bt.B; b += blockDim.x) s_start[b] = bt.info[b];    // vertex_start
    __syncthreads();
  }

  const Entry *table = table_all;
  // lattice offsets: pc1_lattice_offset[r, n] = vertex index of the point's r-th key (transforms.py:166)
  // (slots are global table positions, so this part is the same for one scan and for a batch)
  if (loff || loff32) {
    for (int i = tid; i < n; i += stride) {
      const int4 s = slots[i];
      int v0 = table_all[s.x].index, v1 = table_all[s.y].index, v2 = table_all[s.z].index, v3 = table_all[s.w].index;
      // More vertices than the caller's capacity (EFGH_ST_VERTEX_CAP is set, the level's results are void): indices
      // beyond the capacity become -2 so that no consumer - the splat adds 1 and skips negative rows - ever leaves
      // the vertex-side arrays.
      v0 = v0 < h_cap ? v0 : -2; v1 = v1 < h_cap ? v1 : -2; v2 = v2 < h_cap ? v2 : -2; v3 = v3 < h_cap ? v3 : -2;
      if (loff) {
        long long *lo = reinterpret_cast<long long *>(loff);
        __stcs(lo + i, (long long)v0); __stcs(lo + off_ld + i, (long long)v1);
        __stcs(lo + 2 * off_ld + i, (long long)v2); __stcs(lo + 3 * off_ld + i, (long long)v3);
      }
      if (loff32) {
        loff32[i] = v0; loff32[off_ld + i] = v1; loff32[2 * off_ld + i] = v2; loff32[3 * off_ld + i] = v3;
      }
      if (bt.contrib) {                                      // vertex -> contributions lists (order within a vertex is arbitrary)
        const int v[4] = {v0, v1, v2, v3};
        int pos[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) pos[r] = v[r] >= 0 ? atomicAdd(&bt.cursor[v[r]], 1) : -1;
#pragma unroll
        for (int r = 0; r < 4; ++r) if (pos[r] >= 0) bt.contrib[pos[r]] = 4 * i + r;
      }
    }
  }

  int kmin[4], kmax[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { kmin[c] = st->key_min[c]; kmax[c] = st->key_max[c]; }
  const float E[4][3] = {{EFGH_E_A, EFGH_E_B, EFGH_E_C},
                         {-EFGH_E_A, EFGH_E_B, EFGH_E_C},
                         {0.0f, EFGH_E_B2, EFGH_E_C},
                         {0.0f, 0.0f, EFGH_E_C3}};

  // Four threads per vertex, each owning taps g, g+4, g+8, ...: a thread keeps four independent table probes in
  // flight (the probes are L2-latency bound).  A CTA round covers 64 consecutive vertices x 4 tap groups
  // (thread = 64 g + vertex): stores to nbr[f, :] stay coalesced along the vertex index and all probes of a
  // vertex - and of its scan - happen at the same time, so a batch's hash tables are walked one after the other
  // instead of four times each.
  const bool want_nbr = F > 0 && (nbr || nbr32);
  const int groups = want_nbr ? 4 : 1;
  const int per_round = blockDim.x / groups;                // vertices per CTA round
  const int g = threadIdx.x / per_round;
  int scan_end = 0;
  for (int h = blockIdx.x * per_round + (threadIdx.x - g * per_round); h < H; h += gridDim.x * per_round) {
    if (batched && h >= scan_end) {                            // the vertex's own scan: its table, its key box
      const int scan = find_scan(s_start, bt.B, h);              // (h only grows: looked up again when it leaves the scan)
      scan_end = s_start[scan + 1];
      table = table_all + (long long)scan * bt.tstride;
      mask = (unsigned)bt.info[bt.tm_off + scan];
      const int *box = bt.info + bt.box_off + 8 * scan;
#pragma unroll
      for (int c = 0; c < 4; ++c) { kmin[c] = box[c]; kmax[c] = box[4 + c]; }
    }
    // Spans of the key box.  The reference packs keys with int64 arithmetic that wraps (numba, transforms.py:62-78);
    // unsigned arithmetic reproduces the wrap without undefined behaviour.  A box so wide that the in-box packing
    // itself overflows 2^63 (spans of ~55 000 in every coordinate - far outside any LiDAR cloud) is flagged.
    const unsigned long long s1 = (unsigned long long)((long long)kmax[1] - kmin[1] + 1), s2 = (unsigned long long)((long long)kmax[2] - kmin[2] + 1),
                             s3 = (unsigned long long)((long long)kmax[3] - kmin[3] + 1), s0 = (unsigned long long)((long long)kmax[0] - kmin[0] + 1);
    const unsigned long long s01 = s0 * s1, s23 = s2 * s3;            // each span < 2^22: the pair products are exact
    const bool box_ok = __umul64hi(s01, s23) == 0 && (long long)(s01 * s23) >= 0;
    const unsigned long long box = s01 * s23;
    int k[4];
    unpack_key(vkeys[h], k[0], k[1], k[2]);
    k[3] = -(k[0] + k[1] + k[2]);
    bool aliased = false, wide_box = false;
    if (want_nbr) {
      for (int f0 = g; f0 < F; f0 += 16) {
        unsigned long long want[4];
        unsigned hh[4];
        int res[4];
        bool live[4], ali[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int f = f0 + 4 * t;
          live[t] = false; ali[t] = false; res[t] = -1; want[t] = 0; hh[t] = 0;
          if (f >= F) continue;
          const int4 o = __ldg(reinterpret_cast<const int4 *>(foffs) + f);
          int q[4] = {k[0] + o.x, k[1] + o.y, k[2] + o.z, k[3] + o.w};
          bool in_box = true;
#pragma unroll
          for (int c = 0; c < 4; ++c) in_box = in_box && q[c] >= kmin[c] && q[c] <= kmax[c];
          if (!in_box) {
            // transforms.py:62-78: the reference looks the neighbour up by its mixed-radix packed integer,
            // which can alias a DIFFERENT in-box key when the neighbour lies outside the key box.
            if (!box_ok) { wide_box = true; continue; }
            unsigned long long P = ((((unsigned long long)(long long)(q[0] - kmin[0])) * s1 + (unsigned long long)(long long)(q[1] - kmin[1])) * s2 +
                                    (unsigned long long)(long long)(q[2] - kmin[2])) * s3 + (unsigned long long)(long long)(q[3] - kmin[3]);
            if (!(P < box)) continue;                          // (as int64: negative or beyond every stored key)
            const unsigned long long a3 = P % s3; P /= s3;      // int2key (transforms.py:81-92) on a non-negative value
            const unsigned long long a2 = P % s2; P /= s2;
            const unsigned long long a1 = P % s1; P /= s1;
            q[0] = (int)P + kmin[0]; q[1] = (int)a1 + kmin[1]; q[2] = (int)a2 + kmin[2]; q[3] = (int)a3 + kmin[3];
            ali[t] = true;
          }
          if (q[0] + q[1] + q[2] + q[3] != 0) continue;        // not a lattice point: cannot be in the table
          want[t] = pack_key(q[0], q[1], q[2]);
          hh[t] = hash_key(want[t]) & mask;
          live[t] = true;
        }
        int4 e[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) e[t] = live[t] ? *reinterpret_cast<const int4 *>(table + hh[t]) : make_int4(-1, -1, 0, -1);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (!live[t]) continue;
          unsigned long long cur = ((unsigned long long)(unsigned)e[t].y << 32) | (unsigned)e[t].x;
          int idx = e[t].w;
          unsigned hcur = hh[t];
          for (unsigned probes = 0; cur != want[t] && cur != kEmpty && probes <= mask; ++probes) {   // collisions: rare
            hcur = (hcur + 1) & mask;
            const int4 e2 = *reinterpret_cast<const int4 *>(table + hcur);
            cur = ((unsigned long long)(unsigned)e2.y << 32) | (unsigned)e2.x;
            idx = e2.w;
          }
          res[t] = (cur == want[t] && idx < h_cap) ? idx : -1;   // (idx >= h_cap only when the vertex capacity overflowed)
          if (ali[t] && res[t] >= 0) aliased = true;           // an out-of-box key aliased onto an existing vertex:
                                                               // the neighbour table loses its mirror symmetry
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int f = f0 + 4 * t;
          if (f >= F) continue;
          if (nbr) __stcs(reinterpret_cast<long long *>(nbr) + f * nbr_ld + h, (long long)res[t]);   // reference-format copy: streamed past the L2
          if (nbr32) nbr32[f * nbr_ld + h] = res[t];
        }
      }
    }
    if (aliased) atomicOr(&st->status, EFGH_ST_ALIASED);
    if (wide_box) atomicOr(&st->status, EFGH_ST_KEY_RANGE);
    if (next_pts && g == 0) {                                // generate_data.py:177-178
      float q[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) q[c] = __fdiv_rn((float)k[c], next_divisor);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float acc = __fmul_rn(E[0][a], q[0]);
        acc = __fmaf_rn(E[1][a], q[1], acc);
        acc = __fmaf_rn(E[2][a], q[2], acc);
        acc = __fmaf_rn(E[3][a], q[3], acc);
        next_pts[a * next_ld + h] = acc;
      }
    }
  }
}

}  // namespace

}  // namespace efgh

using namespace efgh;

extern "C" size_t efgh_lattice_workspace_bytes(int64_t n_cap) { return carve(nullptr, n_cap).bytes; }

extern "C" int64_t efgh_lattice_table_entries(int64_t n_scan, int64_t h_scan) {
  if (n_scan < 1) n_scan = 1;
  int64_t v = 4 * n_scan;                                 // a scan of n points has at most 4n vertices
  if (h_scan >= 1 && h_scan < v) v = h_scan;
  return pow2_ceil(2 * v);                                // load factor <= 0.5
}

extern "C" size_t efgh_lattice_batch_workspace_bytes(int B, int64_t table_entries, int64_t n_cap_total) {
  if (B < 1) B = 1;
  if (table_entries < 1) table_entries = 1;
  return carve(nullptr, n_cap_total, B, table_entries).bytes;
}

extern "C" int64_t efgh_lattice_vertex_offsets_ints(int64_t h_cap) { return h_cap + 2 + kMaxHeavy; }

extern "C" int64_t efgh_lattice_batch_info_ints(int B) { return B < 1 ? 0 : (int64_t)batch_box_off(B) + 8 * (int64_t)B; }

namespace {

int lattice_points_impl(const char *who, const float *pts, int64_t pts_ld, int64_t n, const int32_t *n_dev, float scale,
                        float *barycentric, float *el_minus_gr, int64_t out_ld, int64_t h_cap, efgh_lattice_state *state,
                        void *workspace, size_t workspace_bytes, void *stream, Batch bt, int64_t table_entries) {
  EFGH_REQUIRE(n >= 0 && n < (1ll << 28), "%s: n=%lld out of range", who, (long long)n);
  EFGH_REQUIRE(state && workspace && (n == 0 || (pts && barycentric && el_minus_gr)), "%s: null pointer", who);
  EFGH_REQUIRE(pts_ld >= n && out_ld >= n, "%s: leading dimension smaller than n", who);
  const int B = bt.pt_start ? bt.B : 1;
  Workspace w = carve(workspace, n, B, bt.pt_start ? table_entries : -1);
  if (w.bytes > workspace_bytes) {
    set_error("%s: workspace %zu < %zu bytes", who, workspace_bytes, w.bytes);
    return EFGH_ENOMEM;
  }
  EFGH_REQUIRE(w.table_cap * B < (1ll << 31), "%s: hash tables of the batch exceed 2^31 entries", who);
  bt.tstride = w.table_cap;
  bt.cursor = w.cursor;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n_tiles_cap = (int)((n + kTile - 1) / kTile) + 1;
  k_clear<<<grid_for(w.table_cap, 256, 8), 256, 0, s>>>(state, n_dev, (int)n, w.table_cap, w.table, w.tiles, n_tiles_cap, bt);
  EFGH_LAUNCH_CHECK();
  if (n == 0) return EFGH_OK;
  k_points<<<grid_for(n, kPointThreads, 8), kPointThreads, 0, s>>>(pts, pts_ld, scale, barycentric, el_minus_gr,
                                                                   out_ld, state, w.table, w.slots, bt);
  EFGH_LAUNCH_CHECK();
  k_assign<<<(int)((n + kTile - 1) / kTile), kAssignThreads, 0, s>>>(state, w.slots, w.table, w.vkeys, w.tiles,
                                                            (int)(h_cap < (1ll << 30) ? h_cap : (1ll << 30)), bt);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

int lattice_vertices_impl(const char *who, int64_t n, int64_t *lattice_offset, int32_t *lattice_offset32, int64_t off_ld,
                          const int32_t *filter_offsets, int F, int64_t h, int64_t *blur_neighbors,
                          int32_t *blur_neighbors32, int64_t nbr_ld, float *next_pts, int64_t next_ld, float next_divisor,
                          efgh_lattice_state *state, void *workspace, size_t workspace_bytes, void *stream, Batch bt,
                          int64_t table_entries) {
  EFGH_REQUIRE(n >= 0 && n < (1ll << 28) && h >= 0 && h < (1ll << 30), "%s: sizes out of range", who);
  EFGH_REQUIRE(state && workspace, "%s: null pointer", who);
  EFGH_REQUIRE(F <= 0 || filter_offsets, "%s: filter_offsets is null", who);
  EFGH_REQUIRE((!lattice_offset && !lattice_offset32) || off_ld >= n, "%s: off_ld < n", who);
  EFGH_REQUIRE((!blur_neighbors && !blur_neighbors32) || nbr_ld >= h, "%s: nbr_ld < h", who);
  EFGH_REQUIRE(!next_pts || next_ld >= h, "%s: next_ld < h", who);
  const int B = bt.pt_start ? bt.B : 1;
  Workspace w = carve(workspace, n, B, bt.pt_start ? table_entries : -1);
  if (w.bytes > workspace_bytes) {
    set_error("%s: workspace %zu < %zu bytes", who, workspace_bytes, w.bytes);
    return EFGH_ENOMEM;
  }
  bt.tstride = w.table_cap;
  bt.cursor = w.cursor;
  if (n == 0) return EFGH_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t vitems = h * (F > 0 ? 4 : 1);
  const int64_t items = n > vitems ? n : vitems;
  k_vertices<<<grid_for(items, 256, 8), 256, 0, s>>>(state, (int)n, (int)h, w.slots, w.table, w.vkeys,
                                                     lattice_offset, lattice_offset32, off_ld, filter_offsets, F,
                                                     blur_neighbors, blur_neighbors32, nbr_ld, next_pts, next_ld,
                                                     next_divisor, bt);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

int make_batch(const char *who, const int32_t *scan_start, int B, int64_t table_entries, int32_t *batch_info, Batch *bt,
               int32_t *voff, int32_t *contrib, float *point_rows, int64_t h_cap) {
  EFGH_REQUIRE(B >= 1 && B <= kMaxBatch, "%s: batch of %d scans (1..%d supported)", who, B, kMaxBatch);
  EFGH_REQUIRE(scan_start && batch_info && table_entries >= 1, "%s: null scan_start / batch_info or table_entries < 1", who);
  bt->pt_start = scan_start; bt->info = batch_info; bt->B = B;
  bt->tm_off = batch_tm_off(B); bt->box_off = batch_box_off(B); bt->tstride = 0;
  bt->voff = voff; bt->contrib = contrib; bt->cursor = nullptr; bt->point_rows = point_rows;
  bt->h_cap = (int)(h_cap < (1ll << 30) ? h_cap : (1ll << 30));
  EFGH_REQUIRE(!point_rows || (reinterpret_cast<uintptr_t>(point_rows) & 15) == 0, "%s: point_rows must be 16-byte aligned", who);
  return EFGH_OK;
}

}  // namespace

extern "C" int efgh_lattice_points(const float *pts, int64_t pts_ld, int64_t n, const int32_t *n_dev, float scale,
                                   float *barycentric, float *el_minus_gr, int64_t out_ld, int64_t h_cap,
                                   efgh_lattice_state *state, void *workspace, size_t workspace_bytes,
                                   void *stream) {
  Batch bt = {nullptr, nullptr, 1, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, 0};
  return lattice_points_impl("efgh_lattice_points", pts, pts_ld, n, n_dev, scale, barycentric, el_minus_gr, out_ld, h_cap,
                             state, workspace, workspace_bytes, stream, bt, -1);
}

extern "C" int efgh_lattice_vertices(int64_t n, int64_t *lattice_offset, int32_t *lattice_offset32, int64_t off_ld,
                                     const int32_t *filter_offsets, int F, int64_t h, int64_t *blur_neighbors,
                                     int32_t *blur_neighbors32, int64_t nbr_ld, float *next_pts, int64_t next_ld,
                                     float next_divisor, efgh_lattice_state *state, void *workspace,
                                     size_t workspace_bytes, void *stream) {
  Batch bt = {nullptr, nullptr, 1, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, 0};
  return lattice_vertices_impl("efgh_lattice_vertices", n, lattice_offset, lattice_offset32, off_ld, filter_offsets, F, h,
                               blur_neighbors, blur_neighbors32, nbr_ld, next_pts, next_ld, next_divisor, state, workspace,
                               workspace_bytes, stream, bt, -1);
}

extern "C" int efgh_lattice_points_batch(const float *pts, int64_t pts_ld, int64_t n_cap_total, const int32_t *scan_start,
                                         int B, int64_t table_entries, float scale, float *barycentric, float *el_minus_gr,
                                         int64_t out_ld, int64_t h_cap, efgh_lattice_state *state, int32_t *batch_info,
                                         int32_t *vertex_offsets, float *point_rows, void *workspace, size_t workspace_bytes,
                                         void *stream) {
  Batch bt;
  if (int rc = make_batch("efgh_lattice_points_batch", scan_start, B, table_entries, batch_info, &bt, vertex_offsets, nullptr,
                          point_rows, h_cap))
    return rc;
  return lattice_points_impl("efgh_lattice_points_batch", pts, pts_ld, n_cap_total, nullptr, scale, barycentric,
                             el_minus_gr, out_ld, h_cap, state, workspace, workspace_bytes, stream, bt, table_entries);
}

extern "C" int efgh_lattice_vertices_batch(int64_t n_cap_total, const int32_t *scan_start, int B, int64_t table_entries,
                                           int64_t *lattice_offset, int32_t *lattice_offset32, int64_t off_ld,
                                           const int32_t *filter_offsets, int F, int64_t h, int64_t *blur_neighbors,
                                           int32_t *blur_neighbors32, int64_t nbr_ld, float *next_pts, int64_t next_ld,
                                           float next_divisor, efgh_lattice_state *state, int32_t *batch_info,
                                           int32_t *contributions, void *workspace, size_t workspace_bytes, void *stream) {
  Batch bt;
  if (int rc = make_batch("efgh_lattice_vertices_batch", scan_start, B, table_entries, batch_info, &bt, nullptr, contributions,
                          nullptr, h))
    return rc;
  return lattice_vertices_impl("efgh_lattice_vertices_batch", n_cap_total, lattice_offset, lattice_offset32, off_ld,
                               filter_offsets, F, h, blur_neighbors, blur_neighbors32, nbr_ld, next_pts, next_ld,
                               next_divisor, state, workspace, workspace_bytes, stream, bt, table_entries);
}
