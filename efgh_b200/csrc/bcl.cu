// Bilateral convolution layer kernels (reference nets/bilateralNN.py:148-263), fp32 CUDA-core path.
//
//   k_scatter   splat (:176-191) + density sum (:193-207); adjoint of slice
//   k_inv_norm  1 / (wsum + 1e-5)  (:210)
//   k_gather    slice (:251-261); adjoint of splat
//   k_conv      neighbour gather (:240-242) fused with the (F,1) / (1,1) convolution (:244): the
//               reference's (1, C, F, H) gathered tensor (218 MB at level 0) is never materialised -
//               rows of the vertex-major splat matrix go straight from L2 into the shared-memory tile
//   k_dgrad / k_wgrad   backward of k_conv (reference: autograd over index + cuDNN)
//
// Lattice-side matrices are vertex-major (row = vertex, C contiguous floats) so one neighbour is one
// contiguous 4*C-byte read and a splat contribution is a handful of 16-byte vector reductions.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace efgh {
namespace {

constexpr int kTP = 32;  // points per tile in scatter / gather

// ---------------------------------------------------------------------------------------------
// scatter: S[off[r,n]+shift, :] += w[r,n] * [feat ; feat2][:, n]
// Two sources are concatenated along channels on the fly (reference nets/enet.py:113-137 materialises the
// torch.cat); either may be channel-major (stride_n == 1) or point-major.
// ---------------------------------------------------------------------------------------------
// Pointwise stem fused into the splat's tile load (reference nets/enet.py:24-28,111 and
// nets/net_utils.py:35-43): the second feature source is not read from memory but computed per point as
// act(W3 act(W2 act(W1 p + b1) + b2) + b3), act = LeakyReLU(slope) (slope 0 = ReLU), from the point's cin
// coordinates - the (32, N) stem output of E-Net never exists in HBM.
// Packed weights: W1 (c1 x cin) b1 (c1) W2 (c2 x c1) b2 (c2) W3 (c3 x c2) b3 (c3), row-major, c* <= 32, cin <= 4.
struct Stem {
  const float *pts; int64_t pts_ld; const float *w; int cin, c1, c2, c3; float slope;
};
__host__ __device__ inline int stem_weight_floats(int cin, int c1, int c2, int c3) {
  return c1 * cin + c1 + c2 * c1 + c2 + c3 * c2 + c3;
}

template <typename IdxT, int TP>
__global__ void __launch_bounds__(256)
k_scatter(const float *__restrict__ feat, int64_t sc, int64_t sn, int C1, const float *__restrict__ feat2, int64_t sc2,
          int64_t sn2, int C2, int n_host, const int32_t *n_dev, const float *__restrict__ w, int64_t w_ld,
          const void *__restrict__ off, int64_t off_ld, int shift, float *S, int64_t ldS, float *wsum, Stem stem) {
  extern __shared__ float smem[];
  const int C = C1 + C2;
  float *tile = smem;                                   // [C][TP+1]
  float *s_w = smem + (size_t)C * (TP + 1);             // [4][TP]
  int *s_row = reinterpret_cast<int *>(s_w + 4 * TP);   // [4][TP]
  float *s_stem = reinterpret_cast<float *>(s_row + 4 * TP);          // weights | pts [4][TP+1] | h1 [32][TP+1] | h2 [32][TP+1]
  const int stem_nw = stem.pts ? stem_weight_floats(stem.cin, stem.c1, stem.c2, stem.c3) : 0;
  if (stem.pts) {
    // weights transposed to [k][o] so that a thread reads the 4 outputs it owns with one 16-byte broadcast load
    int base = 0;
    const int cins[3] = {stem.cin, stem.c1, stem.c2}, couts[3] = {stem.c1, stem.c2, stem.c3};
    for (int l = 0; l < 3; ++l) {
      const int ci = cins[l], co = couts[l];
      for (int i = threadIdx.x; i < ci * co; i += blockDim.x) {
        const int o = i / ci, k = i - o * ci;
        s_stem[base + k * co + o] = __ldg(stem.w + base + i);
      }
      for (int i = threadIdx.x; i < co; i += blockDim.x) s_stem[base + ci * co + i] = __ldg(stem.w + base + ci * co + i);
      base += ci * co + co;
    }
    __syncthreads();
  }
  const int n = n_dev ? min(*n_dev, n_host) : n_host;
  const int n_tiles = (n + TP - 1) / TP;
  const bool vec = (C % 4 == 0) && (ldS % 4 == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0);
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int n0 = t * TP;
    const int np = min(TP, n - n0);
#pragma unroll
    for (int src = 0; src < 2; ++src) {
      const float *f = src ? feat2 : feat;
      const int64_t fsc = src ? sc2 : sc, fsn = src ? sn2 : sn;
      const int Cs = src ? C2 : C1, c_off = src ? C1 : 0;
      if (Cs == 0) continue;
      if (src == 1 && stem.pts) {
        // three pointwise layers on the tile's points; layer outputs live in shared memory [channel][point]
        float *s_p = s_stem + stem_nw, *s_h1 = s_p + 4 * (TP + 1), *s_h2 = s_h1 + 32 * (TP + 1);
        const float *W1 = s_stem, *b1 = W1 + stem.c1 * stem.cin, *W2 = b1 + stem.c1, *b2 = W2 + stem.c2 * stem.c1,
                    *W3 = b2 + stem.c2, *b3 = W3 + stem.c3 * stem.c2;
        for (int idx = threadIdx.x; idx < stem.cin * TP; idx += blockDim.x) {
          const int a = idx / TP, pt = idx % TP;
          s_p[a * (TP + 1) + pt] = pt < np ? __ldg(stem.pts + a * stem.pts_ld + n0 + pt) : 0.f;
        }
        __syncthreads();
        // thread = (point, group of 4 outputs): per input channel one activation read + one 16-byte weight read
        auto layer = [&](const float *Wt, const float *b, int cin_l, int cout_l, const float *in, float *out, int out_ld) {
          for (int idx = threadIdx.x; idx < (cout_l >> 2) * TP; idx += blockDim.x) {
            const int og = idx / TP, pt = idx - og * TP;
            float4 acc = *reinterpret_cast<const float4 *>(b + 4 * og);
            for (int k = 0; k < cin_l; ++k) {
              const float x = in[k * (TP + 1) + pt];
              const float4 wv = *reinterpret_cast<const float4 *>(Wt + k * cout_l + 4 * og);
              acc.x = fmaf(wv.x, x, acc.x); acc.y = fmaf(wv.y, x, acc.y); acc.z = fmaf(wv.z, x, acc.z); acc.w = fmaf(wv.w, x, acc.w);
            }
            float *o = out + (4 * og) * out_ld + pt;
            o[0] = acc.x > 0.f ? acc.x : stem.slope * acc.x;
            o[out_ld] = acc.y > 0.f ? acc.y : stem.slope * acc.y;
            o[2 * out_ld] = acc.z > 0.f ? acc.z : stem.slope * acc.z;
            o[3 * out_ld] = acc.w > 0.f ? acc.w : stem.slope * acc.w;
          }
          __syncthreads();
        };
        layer(W1, b1, stem.cin, stem.c1, s_p, s_h1, TP + 1);
        layer(W2, b2, stem.c1, stem.c2, s_h1, s_h2, TP + 1);
        layer(W3, b3, stem.c2, stem.c3, s_h2, tile + (size_t)c_off * (TP + 1), TP + 1);
        continue;
      }
      if (fsn == 1) {  // channel-major (C,N): coalesce along points
        for (int idx = threadIdx.x; idx < Cs * TP; idx += blockDim.x) {
          int c = idx / TP, p = idx % TP;
          tile[(c_off + c) * (TP + 1) + p] = p < np ? __ldg(f + c * fsc + (n0 + p)) : 0.f;
        }
      } else {         // point-major: coalesce along channels
        for (int idx = threadIdx.x; idx < Cs * TP; idx += blockDim.x) {
          int p = idx / Cs, c = idx % Cs;
          tile[(c_off + c) * (TP + 1) + p] = p < np ? __ldg(f + c * fsc + (int64_t)(n0 + p) * fsn) : 0.f;
        }
      }
    }
    for (int idx = threadIdx.x; idx < 4 * TP; idx += blockDim.x) {
      int r = idx / TP, p = idx % TP;
      bool ok = p < np;
      s_w[idx] = ok ? __ldg(w + r * w_ld + n0 + p) : 0.f;
      s_row[idx] = ok ? load_idx<IdxT>(off, r * off_ld + n0 + p) + shift : -1;
    }
    __syncthreads();
    if (vec) {
      const int C4 = C / 4;
      for (int item = threadIdx.x; item < np * 4 * C4; item += blockDim.x) {
        const int c4 = item % C4, pr = item / C4, r = pr & 3, p = pr >> 2;
        const int row = s_row[r * TP + p];
        if (row < 0) continue;
        const float wt = s_w[r * TP + p];
        float4 v;
        v.x = tile[(4 * c4 + 0) * (TP + 1) + p] * wt;
        v.y = tile[(4 * c4 + 1) * (TP + 1) + p] * wt;
        v.z = tile[(4 * c4 + 2) * (TP + 1) + p] * wt;
        v.w = tile[(4 * c4 + 3) * (TP + 1) + p] * wt;
        atomicAdd(reinterpret_cast<float4 *>(S + (int64_t)row * ldS) + c4, v);
        if (wsum && c4 == 0) atomicAdd(wsum + row, wt);
      }
    } else {
      for (int item = threadIdx.x; item < np * 4 * C; item += blockDim.x) {
        const int c = item % C, pr = item / C, r = pr & 3, p = pr >> 2;
        const int row = s_row[r * TP + p];
        if (row < 0) continue;
        const float wt = s_w[r * TP + p];
        atomicAdd(S + (int64_t)row * ldS + c, tile[c * (TP + 1) + p] * wt);
        if (wsum && c == 0) atomicAdd(wsum + row, wt);
      }
    }
    __syncthreads();
  }
}

// ---- stem as a kernel of its own --------------------------------------------------------------------------------------
// The gather-form splat (below) wants its per-point features as point-major rows.  For level 0 those are E-Net's stem
// features (reference nets/enet.py:24-28,111): one thread evaluates the three pointwise layers of ONE point entirely
// in registers - weights are read from shared memory as broadcast 16-byte loads (transposed [k][o], zero-padded to
// 32 x 32 so that every loop has a compile-time trip count and the activations stay in registers) - and writes the
// point's 32 features as one 128-byte row.  Same accumulation order as the fused variant in k_scatter (bias first,
// input channels ascending, fmaf).
__global__ void __launch_bounds__(128)
k_stem_rows(Stem stem, int n_host, const int32_t *n_dev, float *__restrict__ out, int64_t out_ld) {
  __shared__ __align__(16) float s_w[3][32 * 32 + 32];          // per layer: Wt[k][o] (32 x 32, zero padded), then bias[32]
  for (int i = threadIdx.x; i < 3 * (32 * 32 + 32); i += blockDim.x) (&s_w[0][0])[i] = 0.f;
  __syncthreads();
  {
    int base = 0;
    const int cins[3] = {stem.cin, stem.c1, stem.c2}, couts[3] = {stem.c1, stem.c2, stem.c3};
    for (int l = 0; l < 3; ++l) {
      const int ci = cins[l], co = couts[l];
      for (int i = threadIdx.x; i < ci * co; i += blockDim.x) {
        const int o = i / ci, k = i - o * ci;
        s_w[l][k * 32 + o] = __ldg(stem.w + base + i);
      }
      for (int i = threadIdx.x; i < co; i += blockDim.x) s_w[l][32 * 32 + i] = __ldg(stem.w + base + ci * co + i);
      base += ci * co + co;
    }
  }
  __syncthreads();
  const int n = n_dev ? min(*n_dev, n_host) : n_host;
  const float slope = stem.slope;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float a[32], b[32];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = k < stem.cin ? __ldg(stem.pts + k * stem.pts_ld + i) : 0.f;
    // layer 1 (cin <= 4 inputs)
#pragma unroll
    for (int o4 = 0; o4 < 8; ++o4) {
      float4 acc = *reinterpret_cast<const float4 *>(&s_w[0][32 * 32 + 4 * o4]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 w = *reinterpret_cast<const float4 *>(&s_w[0][k * 32 + 4 * o4]);
        acc.x = fmaf(w.x, a[k], acc.x); acc.y = fmaf(w.y, a[k], acc.y); acc.z = fmaf(w.z, a[k], acc.z); acc.w = fmaf(w.w, a[k], acc.w);
      }
      b[4 * o4] = acc.x > 0.f ? acc.x : slope * acc.x; b[4 * o4 + 1] = acc.y > 0.f ? acc.y : slope * acc.y;
      b[4 * o4 + 2] = acc.z > 0.f ? acc.z : slope * acc.z; b[4 * o4 + 3] = acc.w > 0.f ? acc.w : slope * acc.w;
    }
    // layers 2 and 3 (<= 32 inputs each): b -> a -> b
#pragma unroll
    for (int l = 1; l < 3; ++l) {
      float *in = l == 1 ? b : a, *o = l == 1 ? a : b;
#pragma unroll
      for (int o4 = 0; o4 < 8; ++o4) {
        float4 acc = *reinterpret_cast<const float4 *>(&s_w[l][32 * 32 + 4 * o4]);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float4 w = *reinterpret_cast<const float4 *>(&s_w[l][k * 32 + 4 * o4]);
          acc.x = fmaf(w.x, in[k], acc.x); acc.y = fmaf(w.y, in[k], acc.y); acc.z = fmaf(w.z, in[k], acc.z); acc.w = fmaf(w.w, in[k], acc.w);
        }
        o[4 * o4] = acc.x > 0.f ? acc.x : slope * acc.x; o[4 * o4 + 1] = acc.y > 0.f ? acc.y : slope * acc.y;
        o[4 * o4 + 2] = acc.z > 0.f ? acc.z : slope * acc.z; o[4 * o4 + 3] = acc.w > 0.f ? acc.w : slope * acc.w;
      }
    }
    float4 *row = reinterpret_cast<float4 *>(out + (int64_t)i * out_ld);
    const int c34 = stem.c3 >> 2;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q < c34) row[q] = make_float4(b[4 * q], b[4 * q + 1], b[4 * q + 2], b[4 * q + 3]);
  }
}

// ---- warp-private splat (level 0) -----------------------------------------------------------------------------------
// The tile kernel above spends its time between CTA barriers: load tile -> __syncthreads -> atomics -> __syncthreads, with
// four scalar shared-memory loads and a division per vector atomic; it reaches ~125-150 G red.v4/s where the hardware
// sustains ~320 G/s into a destination that is L2-resident (tools/bulk_reduce_probe.cu: 16 scans' 230 MB matrix at
// random 126 G/s, <= 115 MB 313 G/s - and the splat walks the batch scan by scan, so one scan's 14.5 MB is what is hot).
// Here every WARP owns its 32-point tile: lane = point for the coalesced channel-major loads (or for the stem, computed
// in registers), the point's C floats go to a warp-private POINT-major shared-memory row with 16-byte stores, and after
// one __syncwarp the warp issues the tile's 4 * 32 * C/4 vector atomics - one 16-byte shared load, four multiplies, one
// red.v4 each; nine consecutive lanes cover one 144-byte row of S.  No CTA-wide barrier anywhere in the loop.
// C4T: C / 4 at compile time (9 for E-Net's level 0; 0 = run-time value).
__device__ __forceinline__ void stem_eval_point(const float (*s_w)[32 * 32 + 32], const Stem &stem, int i, float *b) {
  float a[32];
  const float slope = stem.slope;
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = k < stem.cin ? __ldg(stem.pts + k * stem.pts_ld + i) : 0.f;
#pragma unroll
  for (int o4 = 0; o4 < 8; ++o4) {
    float4 acc = *reinterpret_cast<const float4 *>(&s_w[0][32 * 32 + 4 * o4]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 w = *reinterpret_cast<const float4 *>(&s_w[0][k * 32 + 4 * o4]);
      acc.x = fmaf(w.x, a[k], acc.x); acc.y = fmaf(w.y, a[k], acc.y); acc.z = fmaf(w.z, a[k], acc.z); acc.w = fmaf(w.w, a[k], acc.w);
    }
    b[4 * o4] = acc.x > 0.f ? acc.x : slope * acc.x; b[4 * o4 + 1] = acc.y > 0.f ? acc.y : slope * acc.y;
    b[4 * o4 + 2] = acc.z > 0.f ? acc.z : slope * acc.z; b[4 * o4 + 3] = acc.w > 0.f ? acc.w : slope * acc.w;
  }
#pragma unroll
  for (int l = 1; l < 3; ++l) {
    float *in = l == 1 ? b : a, *o = l == 1 ? a : b;
#pragma unroll
    for (int o4 = 0; o4 < 8; ++o4) {
      float4 acc = *reinterpret_cast<const float4 *>(&s_w[l][32 * 32 + 4 * o4]);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float4 w = *reinterpret_cast<const float4 *>(&s_w[l][k * 32 + 4 * o4]);
        acc.x = fmaf(w.x, in[k], acc.x); acc.y = fmaf(w.y, in[k], acc.y); acc.z = fmaf(w.z, in[k], acc.z); acc.w = fmaf(w.w, in[k], acc.w);
      }
      o[4 * o4] = acc.x > 0.f ? acc.x : slope * acc.x; o[4 * o4 + 1] = acc.y > 0.f ? acc.y : slope * acc.y;
      o[4 * o4 + 2] = acc.z > 0.f ? acc.z : slope * acc.z; o[4 * o4 + 3] = acc.w > 0.f ? acc.w : slope * acc.w;
    }
  }
}

__device__ __forceinline__ void stem_load_weights(float (*s_w)[32 * 32 + 32], const Stem &stem) {
  for (int i = threadIdx.x; i < 3 * (32 * 32 + 32); i += blockDim.x) (&s_w[0][0])[i] = 0.f;
  __syncthreads();
  int base = 0;
  const int cins[3] = {stem.cin, stem.c1, stem.c2}, couts[3] = {stem.c1, stem.c2, stem.c3};
  for (int l = 0; l < 3; ++l) {
    const int ci = cins[l], co = couts[l];
    for (int i = threadIdx.x; i < ci * co; i += blockDim.x) {
      const int o = i / ci, k = i - o * ci;
      s_w[l][k * 32 + o] = __ldg(stem.w + base + i);
    }
    for (int i = threadIdx.x; i < co; i += blockDim.x) s_w[l][32 * 32 + i] = __ldg(stem.w + base + ci * co + i);
    base += ci * co + co;
  }
  __syncthreads();
}

template <typename IdxT, int C4T, bool STEM>
__global__ void __launch_bounds__(STEM ? 128 : 256)
k_scatter_warp(const float *__restrict__ feat, int64_t sc, int C1, const float *__restrict__ feat2, int64_t sc2, int C2, int n_host,
               const int32_t *n_dev, const float *__restrict__ w, int64_t w_ld, const void *__restrict__ off, int64_t off_ld,
               int shift, float *S, int64_t ldS, float *wsum, Stem stem) {
  extern __shared__ __align__(16) float smem[];
  const int C = C1 + C2, C4 = C4T ? C4T : C >> 2;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // per warp: tile [32 points][C] | weights [4][32] | rows [4][32]
  const int warp_floats = 32 * C + 256;
  float *tile = smem + (size_t)wid * warp_floats;
  float *s_w = tile + 32 * C;
  int *s_row = reinterpret_cast<int *>(s_w + 128);
  float (*s_stem)[32 * 32 + 32] = reinterpret_cast<float (*)[32 * 32 + 32]>(smem + (size_t)nw * warp_floats);
  if (STEM) stem_load_weights(s_stem, stem);
  const int n = n_dev ? min(*n_dev, n_host) : n_host;
  const int n_tiles = (n + 31) >> 5;
  const int warps = gridDim.x * nw;
  for (int t = blockIdx.x * nw + wid; t < n_tiles; t += warps) {
    const int i = t * 32 + lane;
    const bool ok = i < n;
    const int np = min(32, n - t * 32);
    float4 *my = reinterpret_cast<float4 *>(tile + lane * C);
    // ---- the point's C channels -> its shared-memory row (16-byte stores: conflict-free at a 144-byte pitch)
    for (int q = 0; q < (C1 >> 2); ++q) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) { v.x = __ldg(feat + (4 * q) * sc + i); v.y = __ldg(feat + (4 * q + 1) * sc + i); v.z = __ldg(feat + (4 * q + 2) * sc + i); v.w = __ldg(feat + (4 * q + 3) * sc + i); }
      my[q] = v;
    }
    if (STEM) {
      float b[32];
      stem_eval_point(s_stem, stem, ok ? i : 0, b);
      const int c34 = stem.c3 >> 2;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < c34) my[(C1 >> 2) + q] = make_float4(b[4 * q], b[4 * q + 1], b[4 * q + 2], b[4 * q + 3]);
    } else {
      for (int q0 = 0; q0 < (C2 >> 2); q0 += 4) {              // 16 independent loads in flight per lane
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q0 + u;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && q < (C2 >> 2)) {
            v[u].x = __ldg(feat2 + (4 * q) * sc2 + i); v[u].y = __ldg(feat2 + (4 * q + 1) * sc2 + i);
            v[u].z = __ldg(feat2 + (4 * q + 2) * sc2 + i); v[u].w = __ldg(feat2 + (4 * q + 3) * sc2 + i);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (q0 + u < (C2 >> 2)) my[(C1 >> 2) + q0 + u] = v[u];
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      s_w[r * 32 + lane] = ok ? __ldg(w + r * w_ld + i) : 0.f;
      s_row[r * 32 + lane] = ok ? load_idx<IdxT>(off, r * off_ld + i) + shift : -1;
    }
    __syncwarp();
    // ---- the tile's vector atomics: item = (point, remainder, 16-byte piece), pieces of one row on consecutive lanes
    const int items = np * 4 * C4;
    for (int it = lane; it < items; it += 32) {
      const int pr = it / C4, c4 = it - pr * C4;
      const int r = pr & 3, pt = pr >> 2;
      const int row = s_row[r * 32 + pt];
      if (row < 0) continue;
      const float wt = s_w[r * 32 + pt];
      float4 v = *reinterpret_cast<const float4 *>(tile + pt * C + 4 * c4);
      v.x *= wt; v.y *= wt; v.z *= wt; v.w *= wt;
      atomicAdd(reinterpret_cast<float4 *>(S + (int64_t)row * ldS) + c4, v);
      if (wsum && c4 == 0) atomicAdd(wsum + row, wt);
    }
    __syncwarp();
  }
}

// ---- gather-form splat ---------------------------------------------------------------------------------------------
// The atomic splat above is bound by the L2's atomic units (~150 G red.v4/s measured) and every lattice vertex
// receives 5 - 20 contributions.  With the vertex -> contributions lists the lattice build can emit
// (efgh_lattice_*_batch: vertex_offsets / contributions / point_rows) the same sums are formed in registers: one warp
// per vertex, LPR lanes x float4 span the previous level's point-major row, 32 / LPR contributions are processed
// side by side, the partial sums meet through shuffles, and the row is written ONCE - already multiplied by the
// density normalisation 1 / (sum of weights + 1e-5) (reference nets/bilateralNN.py:193-211).  That replaces the
// zero-fill of S, the atomics and the normalisation pass.  The level's own 4 el_minus_gr channels and the 4
// barycentric weights of a point come from ONE 32-byte sector (point_rows).  Vertices with more than kHeavy
// contributions (coincident points) are listed by the lattice build and get a whole CTA.
constexpr int kSplatHeavy = 64;        // == lattice.cu kHeavy
constexpr int kSplatMaxHeavy = 16384;  // == lattice.cu kMaxHeavy

template <int NV, int LPR>
struct SplatAcc {
  float4 acc[NV];
  float acc1, wacc;
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc1 = 0.f; wacc = 0.f;
  }
  // contributions [base, min(base + 32, e1)) of one vertex
  __device__ __forceinline__ void batch(const float *__restrict__ point_rows, const float *__restrict__ feat2, int64_t sn2, int C24,
                                        const int32_t *__restrict__ contrib, int base, int e1, int lane) {
    constexpr int G = 32 / LPR;
    const int g = lane / LPR, l = lane - g * LPR;
    const int mine = base + lane < e1 ? __ldg(contrib + base + lane) : 0;      // 32 list entries per coalesced load
    const float wmine = base + lane < e1 ? __ldg(point_rows + (int64_t)(mine >> 2) * 8 + 4 + (mine & 3)) : 0.f;
    const int cnt = min(32, e1 - base);
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += G) {                                      // warp-uniform; fully unrolled: the loads of all
      if (j0 >= cnt) break;                                                    // iterations are independent and issue back to back
      const int j = j0 + g;
      const int p = __shfl_sync(0xffffffffu, mine, j & 31);
      const float wt = __shfl_sync(0xffffffffu, wmine, j & 31);              // 0 beyond the list: contributes nothing
      const int i = p >> 2;
      if (l < 4) acc1 = fmaf(wt, __ldg(point_rows + (int64_t)i * 8 + l), acc1);
      if (l == 0) wacc += wt;
      const float4 *row = reinterpret_cast<const float4 *>(feat2 + (int64_t)i * sn2);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = l + k * LPR;
        if (c4 < C24) {
          const float4 x = __ldg(row + c4);
          acc[k].x = fmaf(wt, x.x, acc[k].x); acc[k].y = fmaf(wt, x.y, acc[k].y);
          acc[k].z = fmaf(wt, x.z, acc[k].z); acc[k].w = fmaf(wt, x.w, acc[k].w);
        }
      }
    }
  }
  __device__ __forceinline__ void combine_groups() {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
      wacc += __shfl_xor_sync(0xffffffffu, wacc, o);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o); acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
        acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o); acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
      }
    }
  }
  // lanes of group 0 write the row (wacc valid on lane 0)
  __device__ __forceinline__ void store(float *__restrict__ S, int64_t ldS, int v, int C24, int normalize, float *inv_out, int lane) {
    const float wsum = __shfl_sync(0xffffffffu, wacc, 0);
    const float inv = __fdiv_rn(1.0f, __fadd_rn(wsum, 1e-5f));
    const float scl = normalize ? inv : 1.0f;
    if (lane < LPR) {
      float *out = S + (int64_t)(v + 1) * ldS;
      if (lane < 4) out[lane] = acc1 * scl;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = lane + k * LPR;
        if (c4 < C24) {
          float4 o4 = acc[k];
          o4.x *= scl; o4.y *= scl; o4.z *= scl; o4.w *= scl;
          *reinterpret_cast<float4 *>(out + 4 + 4 * c4) = o4;
        }
      }
      if (inv_out && lane == 0) inv_out[v + 1] = inv;
    }
  }
};

template <int NV, int LPR>
__global__ void __launch_bounds__(256)
k_splat_gather(const float *__restrict__ point_rows, const float *__restrict__ feat2, int64_t sn2, int C2,
               const int32_t *__restrict__ voff, const int32_t *__restrict__ contrib, int rows_host,
               const int32_t *rows_dev, int normalize, float *__restrict__ S, int64_t ldS, float *__restrict__ inv_out) {
  const int H = rows_dev ? min(*rows_dev, rows_host) : rows_host;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int C = 4 + C2, C24 = C2 >> 2;
  if (blockIdx.x == 0 && threadIdx.x < C) S[threadIdx.x] = 0.f;                 // sink row ("-1 neighbour")
  if (blockIdx.x == 0 && threadIdx.x == 0 && inv_out) inv_out[0] = __fdiv_rn(1.0f, 1e-5f);
  SplatAcc<NV, LPR> A;

  // heavy vertices: the 8 warps of a CTA share one vertex's list, partial sums meet in shared memory
  const int n_heavy = voff[rows_host + 1];
  const bool listed = n_heavy <= kSplatMaxHeavy;          // (an overflowing list is ignored: the warps below take all)
  if (listed && n_heavy > 0) {
    __shared__ float s_part[8][4 * NV * LPR + 8];
    for (int hi = blockIdx.x; hi < n_heavy; hi += gridDim.x) {
      const int v = voff[rows_host + 2 + hi];
      if (v >= H) continue;
      const int e0 = __ldg(voff + v), e1 = __ldg(voff + v + 1);
      A.clear();
      for (int base = e0 + 32 * wid; base < e1; base += 32 * 8) A.batch(point_rows, feat2, sn2, C24, contrib, base, e1, lane);
      A.combine_groups();
      if (lane < LPR) {
#pragma unroll
        for (int k = 0; k < NV; ++k) *reinterpret_cast<float4 *>(&s_part[wid][4 * (lane + k * LPR)]) = A.acc[k];
        if (lane < 4) s_part[wid][4 * NV * LPR + lane] = A.acc1;
        if (lane == 0) s_part[wid][4 * NV * LPR + 4] = A.wacc;
      }
      __syncthreads();
      if (wid == 0) {
        if (lane < LPR) {
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int w8 = 0; w8 < 8; ++w8) {
              const float4 q = *reinterpret_cast<const float4 *>(&s_part[w8][4 * (lane + k * LPR)]);
              t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
            }
            A.acc[k] = t;
          }
          float t1 = 0.f, tw = 0.f;
          for (int w8 = 0; w8 < 8; ++w8) { t1 += s_part[w8][4 * NV * LPR + (lane & 3)]; tw += s_part[w8][4 * NV * LPR + 4]; }
          A.acc1 = t1; A.wacc = tw;
        }
        A.store(S, ldS, v, C24, normalize, inv_out, lane);
      }
      __syncthreads();
    }
  }

  if constexpr (LPR < 32 && NV == 1) {
    // Narrow rows (C2 <= 64): one vertex per GROUP of LPR lanes, 32 / LPR vertices per warp side by side.  A lattice
    // vertex has only 5 - 20 contributions, so a whole warp on one vertex keeps few row reads in flight; with a vertex
    // per group every lane owns its 4 channels for all contributions (no cross-group reduction) and the warp has
    // 32 / LPR times as many independent loads outstanding.  Groups diverge freely: the shuffles are group-masked.
    constexpr int G = 32 / LPR;
    const int g = lane / LPR, l = lane - g * LPR;
    const unsigned gmask = ((1u << LPR) - 1u) << (g * LPR);
    const int n_groups = warps * G;
    for (int v = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * G + g; v < H; v += n_groups) {
      const int e0 = __ldg(voff + v), e1 = __ldg(voff + v + 1);
      if (listed && e1 - e0 > kSplatHeavy) continue;      // done above
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float acc1 = 0.f, wsum = 0.f;
      for (int base = e0; base < e1; base += LPR) {
        const bool have = base + l < e1;
        const int mine = have ? __ldg(contrib + base + l) : 0;
        const float wmine = have ? __ldg(point_rows + (int64_t)(mine >> 2) * 8 + 4 + (mine & 3)) : 0.f;
        const int nb = min(LPR, e1 - base);
#pragma unroll
        for (int j = 0; j < LPR; ++j) {
          if (j >= nb) break;                              // group-uniform
          const int pj = __shfl_sync(gmask, mine, j, LPR);
          const float wt = __shfl_sync(gmask, wmine, j, LPR);
          const int i = pj >> 2;
          if (l < 4) acc1 = fmaf(wt, __ldg(point_rows + (int64_t)i * 8 + l), acc1);
          wsum += wt;
          if (l < C24) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(feat2 + (int64_t)i * sn2) + l);
            acc.x = fmaf(wt, x.x, acc.x); acc.y = fmaf(wt, x.y, acc.y); acc.z = fmaf(wt, x.z, acc.z); acc.w = fmaf(wt, x.w, acc.w);
          }
        }
      }
      const float inv = __fdiv_rn(1.0f, __fadd_rn(wsum, 1e-5f));
      const float scl = normalize ? inv : 1.0f;
      float *out = S + (int64_t)(v + 1) * ldS;
      if (l < 4) out[l] = acc1 * scl;
      if (l < C24) *reinterpret_cast<float4 *>(out + 4 + 4 * l) = make_float4(acc.x * scl, acc.y * scl, acc.z * scl, acc.w * scl);
      if (inv_out && l == 0) inv_out[v + 1] = inv;
    }
  } else {
    for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < H; v += warps) {
      const int e0 = __ldg(voff + v), e1 = __ldg(voff + v + 1);
      if (listed && e1 - e0 > kSplatHeavy) continue;        // done above
      A.clear();
      for (int base = e0; base < e1; base += 32) A.batch(point_rows, feat2, sn2, C24, contrib, base, e1, lane);
      A.combine_groups();
      A.store(S, ldS, v, C24, normalize, inv_out, lane);
    }
  }
}

__device__ __forceinline__ void zero_rows(float *__restrict__ M, int64_t ld, int C, int rows, int64_t tid, int64_t stride) {
  if (ld == C && C % 4 == 0 && (reinterpret_cast<uintptr_t>(M) & 15) == 0) {
    float4 *M4 = reinterpret_cast<float4 *>(M);
    const int64_t total = (int64_t)rows * C / 4;
    for (int64_t i = tid; i < total; i += stride) M4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const int64_t total = (int64_t)rows * C;
    for (int64_t i = tid; i < total; i += stride) M[(i / C) * ld + (i % C)] = 0.f;
  }
}

// One launch zero-fills everything a BCL forward accumulates into: the splat matrix S and density sums (rows + extra)
// and, optionally, the convolution's split-K accumulator Y2 (rows).
__global__ void k_zero(float *__restrict__ S, int64_t ldS, int C, float *__restrict__ wsum, float *__restrict__ Y2,
                       int64_t ldY2, int C2, int rows_host, const int32_t *rows_dev, int rows_extra) {
  const int base = rows_dev ? min(*rows_dev, rows_host) : rows_host;
  const int rows = rows_dev ? min(base + rows_extra, rows_host + rows_extra) : rows_host + rows_extra;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (S) zero_rows(S, ldS, C, rows, tid, stride);
  if (wsum)
    for (int64_t i = tid; i < rows; i += stride) wsum[i] = 0.f;
  if (Y2) zero_rows(Y2, ldY2, C2, base, tid, stride);
}

__global__ void k_normalize(float *__restrict__ S, int64_t ldS, int C, const float *__restrict__ wsum,
                            float *__restrict__ inv_out, int rows_host, const int32_t *rows_dev, int rows_extra) {
  const int rows = rows_dev ? min(*rows_dev + rows_extra, rows_host) : rows_host;
  const int C4 = C / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (C % 4 == 0 && ldS % 4 == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0) {
    const int64_t total = (int64_t)rows * C4;
    for (int64_t i = tid; i < total; i += stride) {
      const int64_t row = i / C4;
      const int c4 = (int)(i - row * C4);
      const float inv = __fdiv_rn(1.0f, __fadd_rn(__ldg(wsum + row), 1e-5f));
      float4 *q = reinterpret_cast<float4 *>(S + row * ldS) + c4;
      float4 v = *q;
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      *q = v;
      if (inv_out && c4 == 0) inv_out[row] = inv;
    }
  } else {
    const int64_t total = (int64_t)rows * C;
    for (int64_t i = tid; i < total; i += stride) {
      const int64_t row = i / C;
      const int c = (int)(i - row * C);
      const float inv = __fdiv_rn(1.0f, __fadd_rn(__ldg(wsum + row), 1e-5f));
      S[row * ldS + c] *= inv;
      if (inv_out && c == 0) inv_out[row] = inv;
    }
  }
}

__global__ void k_inv_norm(const float *__restrict__ wsum, float *__restrict__ inv, int rows_host,
                           const int32_t *rows_dev, int rows_extra) {
  const int rows = rows_dev ? min(*rows_dev + rows_extra, rows_host) : rows_host;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x)
    inv[i] = __fdiv_rn(1.0f, __fadd_rn(wsum[i], 1e-5f));
}

// ---------------------------------------------------------------------------------------------
// gather: out[:, n] = sum_r w[r,n] * Z[off[r,n]+shift, :] * scale[row] + bias
// ---------------------------------------------------------------------------------------------
template <typename IdxT>
__global__ void __launch_bounds__(256)
k_gather(const float *__restrict__ Z, int64_t ldZ, int C, const float *__restrict__ row_scale, int n_host,
         const int32_t *n_dev, const float *__restrict__ w, int64_t w_ld, const void *__restrict__ off,
         int64_t off_ld, int shift, const float *__restrict__ bias, float *__restrict__ out, int64_t sc,
         int64_t sn) {
  extern __shared__ float smem[];
  float *tile = smem;                                   // [C][kTP+1]
  float *s_w = smem + (size_t)C * (kTP + 1);            // [4][kTP]
  int *s_row = reinterpret_cast<int *>(s_w + 4 * kTP);  // [4][kTP]
  const int n = n_dev ? min(*n_dev, n_host) : n_host;
  const int n_tiles = (n + kTP - 1) / kTP;
  const bool vec = (C % 4 == 0) && (ldZ % 4 == 0) && ((reinterpret_cast<uintptr_t>(Z) & 15) == 0);
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int n0 = t * kTP;
    const int np = min(kTP, n - n0);
    for (int idx = threadIdx.x; idx < 4 * kTP; idx += blockDim.x) {
      int r = idx / kTP, p = idx % kTP;
      bool ok = p < np;
      int row = ok ? load_idx<IdxT>(off, r * off_ld + n0 + p) + shift : -1;
      float wt = ok ? __ldg(w + r * w_ld + n0 + p) : 0.f;
      if (row >= 0 && row_scale) wt *= __ldg(row_scale + row);
      s_w[idx] = wt;
      s_row[idx] = row;
    }
    __syncthreads();
    if (vec) {
      const int C4 = C / 4;
      for (int item = threadIdx.x; item < np * C4; item += blockDim.x) {
        const int c4 = item % C4, p = item / C4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = s_row[r * kTP + p];
          if (row < 0) continue;
          const float wt = s_w[r * kTP + p];
          const float4 z = __ldg(reinterpret_cast<const float4 *>(Z + (int64_t)row * ldZ) + c4);
          acc.x = fmaf(wt, z.x, acc.x); acc.y = fmaf(wt, z.y, acc.y);
          acc.z = fmaf(wt, z.z, acc.z); acc.w = fmaf(wt, z.w, acc.w);
        }
        tile[(4 * c4 + 0) * (kTP + 1) + p] = acc.x;
        tile[(4 * c4 + 1) * (kTP + 1) + p] = acc.y;
        tile[(4 * c4 + 2) * (kTP + 1) + p] = acc.z;
        tile[(4 * c4 + 3) * (kTP + 1) + p] = acc.w;
      }
    } else {
      for (int item = threadIdx.x; item < np * C; item += blockDim.x) {
        const int c = item % C, p = item / C;
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = s_row[r * kTP + p];
          if (row < 0) continue;
          acc = fmaf(s_w[r * kTP + p], __ldg(Z + (int64_t)row * ldZ + c), acc);
        }
        tile[c * (kTP + 1) + p] = acc;
      }
    }
    __syncthreads();
    if (sn == 1) {
      for (int idx = threadIdx.x; idx < C * kTP; idx += blockDim.x) {
        int c = idx / kTP, p = idx % kTP;
        if (p < np) out[c * sc + n0 + p] = tile[c * (kTP + 1) + p] + (bias ? __ldg(bias + c) : 0.f);
      }
    } else {
      for (int idx = threadIdx.x; idx < C * kTP; idx += blockDim.x) {
        int p = idx / C, c = idx % C;
        if (p < np) out[c * sc + (int64_t)(n0 + p) * sn] = tile[c * (kTP + 1) + p] + (bias ? __ldg(bias + c) : 0.f);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Tiled fp32 GEMM machinery shared by conv / dgrad / wgrad.
//   block tile BM x BN, K chunk BK = 16, 256 threads, each thread a TM x TN register tile.
//   As[BK][BM+4], Bs[BK][BN+4] (k-major so the inner product reads are float4 and conflict-free).
// ---------------------------------------------------------------------------------------------
constexpr int BK = 16;

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.1f * v;
  return v;
}
__device__ __forceinline__ float act_bwd(float out, int act) {
  if (act == 1) return out > 0.f ? 1.f : 0.f;
  if (act == 2) return out > 0.f ? 1.f : 0.1f;
  return 1.f;
}

template <int BM, int BN, int TM, int TN>
struct Tile {
  static constexpr int TX = BN / TN, TY = BM / TM;
  static_assert(TX * TY == 256, "256 threads");
  float acc[TM][TN];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  }
  __device__ __forceinline__ void mma(const float (*As)[BM + 4], const float (*Bs)[BN + 4]) {
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
};

// Y[h, m] = act(bias[m] + sum_k A[h, k] * Wt[k, m]),  A[h, f*C + c] = X[nbr[f,h]+1, c] * scale[row]
template <typename IdxT, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
k_conv(const float *__restrict__ X, int64_t ldX, int C, const float *__restrict__ row_scale,
       const void *__restrict__ nbr, int64_t nbr_ld, int F, int h_host, const int32_t *h_dev,
       const float *__restrict__ Wt, const float *__restrict__ bias, int M, int act, float *__restrict__ Y,
       int64_t ldY) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  extern __shared__ int s_rows[];  // [F][BM]
  const int H = h_dev ? min(*h_dev, h_host) : h_host;
  const int K = F * C;
  const int m0 = blockIdx.y * BN;
  const bool vecA = (C % 4 == 0) && (ldX % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const bool vecB = (M % 4 == 0) && ((reinterpret_cast<uintptr_t>(Wt) & 15) == 0);
  Tile<BM, BN, TM, TN> T;
  for (int h0 = blockIdx.x * BM; h0 < H; h0 += gridDim.x * BM) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < F * BM; idx += 256) {
      int f = idx / BM, hh = idx % BM;
      int row = -1;
      if (h0 + hh < H) row = nbr ? load_idx<IdxT>(nbr, f * nbr_ld + h0 + hh) + 1 : h0 + hh;
      s_rows[idx] = row;
    }
    __syncthreads();
    T.zero();
    for (int k0 = 0; k0 < K; k0 += BK) {
      // A chunk: BM rows x 16 k  (4 float4 units per row)
      if (vecA) {
        for (int u = threadIdx.x; u < BM * (BK / 4); u += 256) {
          const int hh = u / (BK / 4), q = u % (BK / 4);
          const int k = k0 + 4 * q;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < K) {
            const int f = k / C, c = k - f * C;
            const int row = s_rows[f * BM + hh];
            if (row > 0 || (row == 0 && !nbr)) {
              v = __ldg(reinterpret_cast<const float4 *>(X + (int64_t)row * ldX + c));
              if (row_scale) { const float sc = __ldg(row_scale + row); v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
            }
          }
          As[4 * q + 0][hh] = v.x; As[4 * q + 1][hh] = v.y; As[4 * q + 2][hh] = v.z; As[4 * q + 3][hh] = v.w;
        }
      } else {
        for (int u = threadIdx.x; u < BM * BK; u += 256) {
          const int hh = u / BK, kk = u % BK;
          const int k = k0 + kk;
          float v = 0.f;
          if (k < K) {
            const int f = k / C, c = k - f * C;
            const int row = s_rows[f * BM + hh];
            if (row > 0 || (row == 0 && !nbr)) {
              v = __ldg(X + (int64_t)row * ldX + c);
              if (row_scale) v *= __ldg(row_scale + row);
            }
          }
          As[kk][hh] = v;
        }
      }
      // B chunk: 16 k x BN m
      if (vecB) {
        for (int u = threadIdx.x; u < BK * (BN / 4); u += 256) {
          const int kk = u / (BN / 4), q = u % (BN / 4);
          const int k = k0 + kk, m = m0 + 4 * q;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < K && m < M) v = __ldg(reinterpret_cast<const float4 *>(Wt + (int64_t)k * M + m));
          *reinterpret_cast<float4 *>(&Bs[kk][4 * q]) = v;
        }
      } else {
        for (int u = threadIdx.x; u < BK * BN; u += 256) {
          const int kk = u / BN, mm = u % BN;
          const int k = k0 + kk, m = m0 + mm;
          Bs[kk][mm] = (k < K && m < M) ? __ldg(Wt + (int64_t)k * M + m) : 0.f;
        }
      }
      __syncthreads();
      T.mma(As, Bs);
      __syncthreads();
    }
    const int tx = threadIdx.x % T.TX, ty = threadIdx.x / T.TX;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int h = h0 + ty * TM + i;
      if (h >= H) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int m = m0 + tx * TN + j;
        if (m < M) Y[(int64_t)h * ldY + m] = act_fwd(T.acc[i][j] + (bias ? __ldg(bias + m) : 0.f), act);
      }
    }
  }
}

// dX[nbr[f,h]+1, c] += sum_m dYm[h, m] * Wt[f*C+c, m]     (GEMM (h x M) . (M x F*C), scattered epilogue)
template <typename IdxT, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
k_dgrad(const float *__restrict__ dY, int64_t ldY, const float *__restrict__ act_out, int64_t ldA, int act, int M,
        const void *__restrict__ nbr, int64_t nbr_ld, int F, int h_host, const int32_t *h_dev,
        const float *__restrict__ Wt, int C, float *dX, int64_t ldX) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int H = h_dev ? min(*h_dev, h_host) : h_host;
  const int J = F * C;
  const int j0 = blockIdx.y * BN;
  Tile<BM, BN, TM, TN> T;
  for (int h0 = blockIdx.x * BM; h0 < H; h0 += gridDim.x * BM) {
    T.zero();
    for (int k0 = 0; k0 < M; k0 += BK) {
      for (int u = threadIdx.x; u < BM * BK; u += 256) {
        const int hh = u / BK, kk = u % BK;
        const int h = h0 + hh, m = k0 + kk;
        float v = 0.f;
        if (h < H && m < M) {
          v = __ldg(dY + (int64_t)h * ldY + m);
          if (act_out) v *= act_bwd(__ldg(act_out + (int64_t)h * ldA + m), act);
        }
        As[kk][hh] = v;
      }
      for (int u = threadIdx.x; u < BK * BN; u += 256) {
        const int jj = u / BK, kk = u % BK;
        const int j = j0 + jj, m = k0 + kk;
        Bs[kk][jj] = (j < J && m < M) ? __ldg(Wt + (int64_t)j * M + m) : 0.f;
      }
      __syncthreads();
      T.mma(As, Bs);
      __syncthreads();
    }
    const int tx = threadIdx.x % T.TX, ty = threadIdx.x / T.TX;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int h = h0 + ty * TM + i;
      if (h >= H) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int col = j0 + tx * TN + j;
        if (col >= J) continue;
        if (nbr) {
          const int f = col / C, c = col - f * C;
          const int row = load_idx<IdxT>(nbr, f * nbr_ld + h) + 1;
          if (row > 0) atomicAdd(dX + (int64_t)row * ldX + c, T.acc[i][j]);
        } else {
          dX[(int64_t)h * ldX + col] = T.acc[i][j];
        }
      }
    }
  }
}

// dWt[f*C+c, m] += sum_h A[h, f*C+c] * dYm[h, m];  dbias[m] += sum_h dYm[h, m]
// grid: x = split over vertices (chunks of kWgradChunk), y = (filter tap f, tile of input channels), z = tiles of m.
// A K-tile never straddles a tap, so the neighbour row of each of the 16 vertices of a step is looked up once and
// the gathered rows are read with coalesced 16-byte loads along the channels.
constexpr int kWgradChunk = 1024;
template <typename IdxT, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
k_wgrad(const float *__restrict__ X, int64_t ldX, int C, const float *__restrict__ row_scale,
        const void *__restrict__ nbr, int64_t nbr_ld, int F, int h_host, const int32_t *h_dev,
        const float *__restrict__ dY, int64_t ldY, const float *__restrict__ act_out, int64_t ldA, int act,
        int M, float *dWt, float *dbias) {
  __shared__ float As[BK][BM + 4];  // [vertex of the step][input channel]
  __shared__ float Bs[BK][BN + 4];  // [vertex of the step][m]
  __shared__ int s_row[BK];
  __shared__ float s_scl[BK];
  const int H = h_dev ? min(*h_dev, h_host) : h_host;
  const int ctiles = (C + BM - 1) / BM;
  const int f = blockIdx.y / ctiles, c0 = (blockIdx.y - f * ctiles) * BM, m0 = blockIdx.z * BN;
  const bool vecA = (C % 4 == 0) && (ldX % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const bool vecB = (M % 4 == 0) && (ldY % 4 == 0) && ((reinterpret_cast<uintptr_t>(dY) & 15) == 0) &&
                    (!act_out || (ldA % 4 == 0 && (reinterpret_cast<uintptr_t>(act_out) & 15) == 0));
  Tile<BM, BN, TM, TN> T;
  for (int hc = blockIdx.x * kWgradChunk; hc < H; hc += gridDim.x * kWgradChunk) {
    T.zero();
    float bsum = 0.f;
    const int hend = min(H, hc + kWgradChunk);
    for (int h0 = hc; h0 < hend; h0 += BK) {
      if (threadIdx.x < BK) {
        const int h = h0 + threadIdx.x;
        int row = -1;
        if (h < hend) {
          row = nbr ? load_idx<IdxT>(nbr, f * nbr_ld + h) + 1 : h;
          if (nbr && row == 0) row = -1;                              // sink row: zeros
        }
        s_row[threadIdx.x] = row;
        s_scl[threadIdx.x] = (row >= 0 && row_scale) ? __ldg(row_scale + row) : 1.0f;
      }
      __syncthreads();
      if (vecA) {
        for (int u = threadIdx.x; u < BK * (BM / 4); u += 256) {
          const int hh = u / (BM / 4), q = u % (BM / 4);
          const int c = c0 + 4 * q, row = s_row[hh];
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row >= 0 && c < C) {
            v = __ldg(reinterpret_cast<const float4 *>(X + (int64_t)row * ldX + c));
            const float sc = s_scl[hh];
            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
          }
          *reinterpret_cast<float4 *>(&As[hh][4 * q]) = v;
        }
      } else {
        for (int u = threadIdx.x; u < BK * BM; u += 256) {
          const int hh = u / BM, cc = u % BM;
          const int c = c0 + cc, row = s_row[hh];
          As[hh][cc] = (row >= 0 && c < C) ? __ldg(X + (int64_t)row * ldX + c) * s_scl[hh] : 0.f;
        }
      }
      if (vecB) {
        for (int u = threadIdx.x; u < BK * (BN / 4); u += 256) {
          const int hh = u / (BN / 4), q = u % (BN / 4);
          const int h = h0 + hh, m = m0 + 4 * q;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (h < hend && m < M) {
            v = __ldg(reinterpret_cast<const float4 *>(dY + (int64_t)h * ldY + m));
            if (act_out) {
              const float4 a = __ldg(reinterpret_cast<const float4 *>(act_out + (int64_t)h * ldA + m));
              v.x *= act_bwd(a.x, act); v.y *= act_bwd(a.y, act); v.z *= act_bwd(a.z, act); v.w *= act_bwd(a.w, act);
            }
          }
          *reinterpret_cast<float4 *>(&Bs[hh][4 * q]) = v;
        }
      } else {
        for (int u = threadIdx.x; u < BK * BN; u += 256) {
          const int hh = u / BN, mm = u % BN;
          const int h = h0 + hh, m = m0 + mm;
          float v = 0.f;
          if (h < hend && m < M) {
            v = __ldg(dY + (int64_t)h * ldY + m);
            if (act_out) v *= act_bwd(__ldg(act_out + (int64_t)h * ldA + m), act);
          }
          Bs[hh][mm] = v;
        }
      }
      __syncthreads();
      T.mma(As, Bs);
      if (dbias && blockIdx.y == 0 && threadIdx.x < BN) {
#pragma unroll
        for (int hh = 0; hh < BK; ++hh) bsum += Bs[hh][threadIdx.x];
      }
      __syncthreads();
    }
    const int tx = threadIdx.x % T.TX, ty = threadIdx.x / T.TX;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int c = c0 + ty * TM + i;
      if (c >= C) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int m = m0 + tx * TN + j;
        if (m < M) atomicAdd(dWt + (int64_t)(f * C + c) * M + m, T.acc[i][j]);
      }
    }
    if (dbias && blockIdx.y == 0 && threadIdx.x < BN && m0 + threadIdx.x < M) atomicAdd(dbias + m0 + threadIdx.x, bsum);
  }
}

// dX[h, :] *= act'(Y[h, :]) in place, rows from the device-side count (backward of the ReLU between the two
// convolutions; the activation's derivative is read off its output)
__global__ void k_act_bwd(float *__restrict__ dX, int64_t ldX, const float *__restrict__ Yact, int64_t ldY, int C, int act,
                          int rows_host, const int32_t *rows_dev) {
  const int rows = rows_dev ? min(*rows_dev, rows_host) : rows_host;
  const int C4 = C >> 2;
  const int64_t total = (int64_t)rows * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t h = i / C4;
    const int c4 = (int)(i - h * C4);
    float4 *d = reinterpret_cast<float4 *>(dX + h * ldX) + c4;
    const float4 y = __ldg(reinterpret_cast<const float4 *>(Yact + h * ldY) + c4);
    float4 v = *d;
    v.x *= act_bwd(y.x, act); v.y *= act_bwd(y.y, act); v.z *= act_bwd(y.z, act); v.w *= act_bwd(y.w, act);
    *d = v;
  }
}

// Training loss on the last level's output of a batch of scans (rows [start[b], start[b+1]) belong to scan b):
//   loss = (1/B) sum_b 0.5 * mean(Z_b^2),   dZ[h, :] = Z[h, :] / (B * rows_b * C)
// One pass: the gradient is written and the loss accumulated (block reduction + one atomic per CTA).
__global__ void __launch_bounds__(256)
k_loss_hms(const float *__restrict__ Z, int64_t ldZ, int C, const int32_t *__restrict__ start, int B, float *__restrict__ dZ,
           int64_t ldD, float *__restrict__ loss, int rows_host) {
  __shared__ int s_start[65];
  __shared__ float s_red[8];
  for (int b = threadIdx.x; b <= B; b += blockDim.x) s_start[b] = start[b];
  __syncthreads();
  const int rows = min(s_start[B], rows_host);
  const int C4 = C >> 2;
  const int64_t total = (int64_t)rows * C4;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i / C4);
    const int c4 = (int)(i - (int64_t)h * C4);
    int lo = 0, hi = B;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_start[mid] <= h) lo = mid; else hi = mid; }
    const float scale = 1.0f / ((float)B * (float)(s_start[lo + 1] - s_start[lo]) * (float)C);
    const float4 z = __ldg(reinterpret_cast<const float4 *>(Z + (int64_t)h * ldZ) + c4);
    reinterpret_cast<float4 *>(dZ + (int64_t)h * ldD)[c4] = make_float4(z.x * scale, z.y * scale, z.z * scale, z.w * scale);
    acc += 0.5f * scale * (z.x * z.x + z.y * z.y + z.z * z.z + z.w * z.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    if (t != 0.f) atomicAdd(loss, t);
  }
}

template <typename F>
int dispatch_idx(int idx_bits, F &&f) {
  if (idx_bits == 64) return f((int64_t)0);
  if (idx_bits == 32) return f((int32_t)0);
  set_error("idx_bits must be 32 or 64, got %d", idx_bits);
  return EFGH_EINVAL;
}

}  // namespace
}  // namespace efgh

using namespace efgh;

namespace {
int scatter_impl(const char *who, const float *feat, int64_t stride_c, int64_t stride_n, int C, const float *feat2,
                 int64_t stride_c2, int64_t stride_n2, int C2, int64_t n, const int32_t *n_dev, const float *w, int64_t w_ld,
                 const void *off, int idx_bits, int64_t off_ld, int row_shift, float *S, int64_t ldS, float *wsum, Stem stem,
                 void *stream) {
  if (!feat2 && !stem.pts) C2 = 0;
  EFGH_REQUIRE(C > 0 && C2 >= 0 && n >= 0 && n < (1ll << 30), "%s: bad sizes C=%d C2=%d n=%lld", who, C, C2, (long long)n);
  if (n == 0) return EFGH_OK;
  EFGH_REQUIRE(feat && w && off && S, "%s: null pointer", who);
  // 32-point tiles when a channel-major source needs 128-byte coalescing along points; 8-point tiles otherwise
  // (four times as many CTAs - the deep levels have few points)
  const bool wide = stem.pts ? true : (C >= C2 ? stride_n == 1 : stride_n2 == 1);   // layout of the wider source decides
  const int TP = wide ? 32 : 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Channel-major sources, 16-byte rows: the warp-private kernel (no CTA barriers in the loop).  EFGH_SCATTER=tile
  // keeps the tile kernel (timing studies).
  static const bool tile_only = getenv("EFGH_SCATTER") && !strcmp(getenv("EFGH_SCATTER"), "tile");
  const int Ct = C + C2;
  static const bool warp_always = getenv("EFGH_SCATTER") && !strcmp(getenv("EFGH_SCATTER"), "warp");
  // (measured, 16 scans per launch: with the stem fused the warp-private kernel makes the whole sequence 3.4 % faster; with
  //  the features read from memory its level-0 stage is 10 % faster stand-alone but the CUDA-graph replay, where the next
  //  level's lattice kernels run beside it, 7 % slower than with the tile kernel - so it is the default for the stem only)
  if (!tile_only && (stem.pts || warp_always) && stride_n == 1 && (stem.pts || C2 == 0 || stride_n2 == 1) && C % 4 == 0 && C2 % 4 == 0 && Ct <= 128 && ldS % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(S) & 15) == 0 && (!stem.pts || stem.c3 == C2)) {
    const int threads = stem.pts ? 128 : 256, nw = threads / 32;
    const size_t smem_w = sizeof(float) * ((size_t)nw * (32 * Ct + 256) + (stem.pts ? 3 * (32 * 32 + 32) : 0));
    int per_sm = (int)std::min<size_t>(stem.pts ? 4 : 8, (200 * 1024) / smem_w);
    if (getenv("EFGH_SCATTER_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(getenv("EFGH_SCATTER_CTAS"))));   // timing studies
    EFGH_REQUIRE(per_sm >= 1, "%s: C=%d too large", who, Ct);
    return dispatch_idx(idx_bits, [&](auto tag) -> int {
      using IdxT = decltype(tag);
      auto launch = [&](auto kern) -> int {
        if (smem_w > 48 * 1024) EFGH_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
        kern<<<grid_for((n + 31) / 32, nw, per_sm), threads, smem_w, s>>>(feat, stride_c, C, feat2, stride_c2, C2, (int)n, n_dev, w, w_ld, off, off_ld,
                                                                          row_shift, S, ldS, wsum, stem);
        EFGH_LAUNCH_CHECK();
        return EFGH_OK;
      };
      if (stem.pts) return Ct == 36 ? launch(k_scatter_warp<IdxT, 9, true>) : launch(k_scatter_warp<IdxT, 0, true>);
      return Ct == 36 ? launch(k_scatter_warp<IdxT, 9, false>) : launch(k_scatter_warp<IdxT, 0, false>);
    });
  }
  size_t smem = sizeof(float) * ((size_t)(C + C2) * (TP + 1) + 4 * TP) + sizeof(int) * 4 * TP;
  if (stem.pts) smem += sizeof(float) * (stem_weight_floats(stem.cin, stem.c1, stem.c2, stem.c3) + (4 + 64) * (TP + 1));
  EFGH_REQUIRE(smem <= 200 * 1024, "%s: C=%d too large", who, C + C2);
  return dispatch_idx(idx_bits, [&](auto tag) -> int {
    using IdxT = decltype(tag);
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024) EFGH_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid_for((n + TP - 1) / TP, 1, 8), 256, smem, s>>>(feat, stride_c, stride_n, C, feat2, stride_c2, stride_n2, C2, (int)n,
                                                                 n_dev, w, w_ld, off, off_ld, row_shift, S, ldS, wsum, stem);
      EFGH_LAUNCH_CHECK();
      return EFGH_OK;
    };
    return wide ? launch(k_scatter<IdxT, 32>) : launch(k_scatter<IdxT, 8>);
  });
}
}  // namespace

extern "C" int efgh_bcl_scatter(const float *feat, int64_t stride_c, int64_t stride_n, int C, const float *feat2,
                                int64_t stride_c2, int64_t stride_n2, int C2, int64_t n, const int32_t *n_dev,
                                const float *w, int64_t w_ld, const void *off, int idx_bits, int64_t off_ld, int row_shift,
                                float *S, int64_t ldS, float *wsum, void *stream) {
  Stem none = {nullptr, 0, nullptr, 0, 0, 0, 0, 0.f};
  return scatter_impl("efgh_bcl_scatter", feat, stride_c, stride_n, C, feat2, stride_c2, stride_n2, C2, n, n_dev, w, w_ld, off,
                      idx_bits, off_ld, row_shift, S, ldS, wsum, none, stream);
}

extern "C" int64_t efgh_bcl_stem_weight_floats(int c_in, int c1, int c2, int c3) { return stem_weight_floats(c_in, c1, c2, c3); }

extern "C" int efgh_bcl_scatter_stem(const float *feat, int64_t stride_c, int64_t stride_n, int C, const float *pts,
                                     int64_t pts_ld, int c_in, int c1, int c2, int c3, const float *stem_weights,
                                     float leaky_slope, int64_t n, const int32_t *n_dev, const float *w, int64_t w_ld,
                                     const void *off, int idx_bits, int64_t off_ld, int row_shift, float *S, int64_t ldS,
                                     float *wsum, void *stream) {
  EFGH_REQUIRE(pts && stem_weights && pts_ld >= n, "efgh_bcl_scatter_stem: null pts / weights or pts_ld < n");
  EFGH_REQUIRE(c_in >= 1 && c_in <= 4 && c1 >= 4 && c1 <= 32 && c2 >= 4 && c2 <= 32 && c3 >= 4 && c3 <= 32 &&
                   c1 % 4 == 0 && c2 % 4 == 0 && c3 % 4 == 0,
               "efgh_bcl_scatter_stem: stem widths %d -> %d -> %d -> %d (inputs <= 4, layers multiples of 4 up to 32)", c_in, c1, c2, c3);
  EFGH_REQUIRE(C % 4 == 0, "efgh_bcl_scatter_stem: C=%d must be a multiple of 4 (shared-memory alignment of the stem weights)", C);
  Stem st = {pts, pts_ld, stem_weights, c_in, c1, c2, c3, leaky_slope};
  return scatter_impl("efgh_bcl_scatter_stem", feat, stride_c, stride_n, C, nullptr, 0, 0, c3, n, n_dev, w, w_ld, off, idx_bits,
                      off_ld, row_shift, S, ldS, wsum, st, stream);
}

extern "C" int efgh_bcl_stem_rows(const float *pts, int64_t pts_ld, int c_in, int c1, int c2, int c3, const float *stem_weights,
                                  float leaky_slope, int64_t n, const int32_t *n_dev, float *out, int64_t out_ld, void *stream) {
  EFGH_REQUIRE(pts && stem_weights && out && pts_ld >= n, "efgh_bcl_stem_rows: null pointer or pts_ld < n");
  EFGH_REQUIRE(c_in >= 1 && c_in <= 4 && c1 >= 4 && c1 <= 32 && c2 >= 4 && c2 <= 32 && c3 >= 4 && c3 <= 32 && c1 % 4 == 0 &&
                   c2 % 4 == 0 && c3 % 4 == 0,
               "efgh_bcl_stem_rows: stem widths %d -> %d -> %d -> %d (inputs <= 4, layers multiples of 4 up to 32)", c_in, c1, c2, c3);
  EFGH_REQUIRE(n >= 0 && n < (1ll << 30) && out_ld >= c3 && out_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "efgh_bcl_stem_rows: bad sizes or unaligned output");
  if (n == 0) return EFGH_OK;
  Stem st = {pts, pts_ld, stem_weights, c_in, c1, c2, c3, leaky_slope};
  k_stem_rows<<<grid_for(n, 128, 16), 128, 0, static_cast<cudaStream_t>(stream)>>>(st, (int)n, n_dev, out, out_ld);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

extern "C" int efgh_bcl_splat_gather(const float *point_rows, const float *feat2, int64_t stride_n2, int C2,
                                     const int32_t *vertex_offsets, const int32_t *contributions, int64_t rows,
                                     const int32_t *rows_dev, int normalize, float *S, int64_t ldS, float *inv_out,
                                     void *stream) {
  EFGH_REQUIRE(C2 > 0 && C2 % 4 == 0 && C2 <= 512, "efgh_bcl_splat_gather: C2=%d must be a multiple of 4, at most 512", C2);
  EFGH_REQUIRE(rows >= 0 && rows < (1ll << 30), "efgh_bcl_splat_gather: bad rows");
  EFGH_REQUIRE(point_rows && feat2 && vertex_offsets && contributions && S, "efgh_bcl_splat_gather: null pointer");
  EFGH_REQUIRE(stride_n2 % 4 == 0 && ldS % 4 == 0 && ldS >= 4 + C2 && (reinterpret_cast<uintptr_t>(feat2) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(S) & 15) == 0 && (reinterpret_cast<uintptr_t>(point_rows) & 15) == 0,
               "efgh_bcl_splat_gather: point_rows / feat2 / S rows must be 16-byte aligned (strides multiples of 4)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for((rows + 1) * 32, 256, 8);
  const int c24 = C2 / 4;
#define EFGH_SPLAT(NV, LPR)                                                                                              \
  k_splat_gather<NV, LPR><<<grid, 256, 0, s>>>(point_rows, feat2, stride_n2, C2, vertex_offsets, contributions, (int)rows, \
                                               rows_dev, normalize, S, ldS, inv_out)
  if (c24 <= 8) EFGH_SPLAT(1, 8);
  else if (c24 <= 16) EFGH_SPLAT(1, 16);
  else if (c24 <= 32) EFGH_SPLAT(1, 32);
  else if (c24 <= 64) EFGH_SPLAT(2, 32);
  else EFGH_SPLAT(4, 32);
#undef EFGH_SPLAT
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

extern "C" int efgh_bcl_zero(float *S, int64_t ldS, int C, float *wsum, float *Y2, int64_t ldY2, int C2, int64_t rows,
                             const int32_t *rows_dev, int rows_extra, void *stream) {
  EFGH_REQUIRE(rows >= 0 && rows < (1ll << 31) && rows_extra >= 0, "efgh_bcl_zero: bad sizes");
  EFGH_REQUIRE(!S || (C > 0 && ldS >= C), "efgh_bcl_zero: bad S shape");
  EFGH_REQUIRE(!Y2 || (C2 > 0 && ldY2 >= C2), "efgh_bcl_zero: bad Y2 shape");
  if (rows + rows_extra == 0 || (!S && !wsum && !Y2)) return EFGH_OK;
  const int64_t work = (rows + rows_extra) * (int64_t)((S ? C : 1) + (Y2 ? C2 : 0)) / 4;
  k_zero<<<grid_for(work, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(S, ldS, C, wsum, Y2, ldY2, C2, (int)rows, rows_dev,
                                                                               rows_extra);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

extern "C" int efgh_bcl_normalize(float *S, int64_t ldS, int C, const float *wsum, float *inv_out, int64_t rows,
                                  const int32_t *rows_dev, int rows_extra, void *stream) {
  EFGH_REQUIRE(C > 0 && ldS >= C && rows >= 0 && rows < (1ll << 31), "efgh_bcl_normalize: bad sizes");
  if (rows == 0) return EFGH_OK;
  EFGH_REQUIRE(S && wsum, "efgh_bcl_normalize: null pointer");
  k_normalize<<<grid_for(rows * C / 4, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(S, ldS, C, wsum, inv_out, (int)rows,
                                                                                          rows_dev, rows_extra);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

extern "C" int efgh_bcl_inv_norm(const float *wsum, float *inv, int64_t rows, const int32_t *rows_dev, int rows_extra,
                                 void *stream) {
  EFGH_REQUIRE(rows >= 0 && rows < (1ll << 31), "efgh_bcl_inv_norm: bad rows");
  if (rows == 0) return EFGH_OK;
  EFGH_REQUIRE(wsum && inv, "efgh_bcl_inv_norm: null pointer");
  k_inv_norm<<<grid_for(rows, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(wsum, inv, (int)rows, rows_dev,
                                                                                     rows_extra);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

extern "C" int efgh_bcl_gather(const float *Z, int64_t ldZ, int C, const float *row_scale, int64_t n,
                               const int32_t *n_dev, const float *w, int64_t w_ld, const void *off, int idx_bits,
                               int64_t off_ld, int row_shift, const float *bias, float *out, int64_t stride_c,
                               int64_t stride_n, void *stream) {
  EFGH_REQUIRE(C > 0 && n >= 0 && n < (1ll << 30), "efgh_bcl_gather: bad sizes C=%d n=%lld", C, (long long)n);
  if (n == 0) return EFGH_OK;
  EFGH_REQUIRE(Z && w && off && out, "efgh_bcl_gather: null pointer");
  const size_t smem = sizeof(float) * ((size_t)C * (kTP + 1) + 4 * kTP) + sizeof(int) * 4 * kTP;
  EFGH_REQUIRE(smem <= 200 * 1024, "efgh_bcl_gather: C=%d too large", C);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dispatch_idx(idx_bits, [&](auto tag) -> int {
    using IdxT = decltype(tag);
    auto kern = k_gather<IdxT>;
    if (smem > 48 * 1024) EFGH_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_for((n + kTP - 1) / kTP, 1, 8), 256, smem, s>>>(Z, ldZ, C, row_scale, (int)n, n_dev, w, w_ld, off, off_ld,
                                                                 row_shift, bias, out, stride_c, stride_n);
    EFGH_LAUNCH_CHECK();
    return EFGH_OK;
  });
}

extern "C" int efgh_bcl_conv(const float *X, int64_t ldX, int C, const float *row_scale, const void *nbr, int idx_bits,
                             int64_t nbr_ld, int F, int64_t h, const int32_t *h_dev, const float *Wt,
                             const float *bias, int M, int act, float *Y, int64_t ldY, int precision, void *stream) {
  EFGH_REQUIRE(C > 0 && M > 0 && h >= 0 && h < (1ll << 30), "efgh_bcl_conv: bad sizes");
  EFGH_REQUIRE(precision == 0, "efgh_bcl_conv: precision %d not available in this build", precision);
  if (!nbr) F = 1;
  EFGH_REQUIRE(F >= 1 && F <= 1024, "efgh_bcl_conv: bad filter size %d", F);
  if (h == 0) return EFGH_OK;
  EFGH_REQUIRE(X && Wt && Y, "efgh_bcl_conv: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dispatch_idx(idx_bits, [&](auto tag) -> int {
    using IdxT = decltype(tag);
    // dynamic shared memory: the tile's neighbour rows, F x BM ints - beyond 48 KB the kernel attribute must be raised
    auto launch = [&](auto kern, int BM, int BN) -> int {
      const size_t smem = sizeof(int) * (size_t)F * BM;
      EFGH_REQUIRE(smem <= 160 * 1024, "efgh_bcl_conv: filter size %d too large for the fp32 kernel (row table %zu bytes)", F, smem);
      if (smem > 48 * 1024) EFGH_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      dim3 grid(grid_for((h + BM - 1) / BM, 1, 4), (M + BN - 1) / BN);
      kern<<<grid, 256, smem, s>>>(X, ldX, C, row_scale, nbr, nbr_ld, F, (int)h, h_dev, Wt, bias, M, act, Y, ldY);
      EFGH_LAUNCH_CHECK();
      return EFGH_OK;
    };
    if (M <= 32) return launch(k_conv<IdxT, 128, 32, 4, 4>, 128, 32);
    return launch(k_conv<IdxT, 64, 64, 4, 4>, 64, 64);
  });
}

extern "C" int efgh_bcl_conv_dgrad(const float *dY, int64_t ldY, const float *act_out, int64_t ldA, int act, int M,
                                   const void *nbr, int idx_bits, int64_t nbr_ld, int F, int64_t h,
                                   const int32_t *h_dev, const float *Wt, int C, float *dX, int64_t ldX, void *stream) {
  EFGH_REQUIRE(C > 0 && M > 0 && h >= 0 && h < (1ll << 30), "efgh_bcl_conv_dgrad: bad sizes");
  if (!nbr) F = 1;
  EFGH_REQUIRE(F >= 1, "efgh_bcl_conv_dgrad: bad filter size %d", F);
  if (h == 0) return EFGH_OK;
  EFGH_REQUIRE(dY && Wt && dX, "efgh_bcl_conv_dgrad: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dispatch_idx(idx_bits, [&](auto tag) -> int {
    using IdxT = decltype(tag);
    constexpr int BM = 64, BN = 64;
    dim3 grid(grid_for((h + BM - 1) / BM, 1, 4), (F * C + BN - 1) / BN);
    k_dgrad<IdxT, BM, BN, 4, 4><<<grid, 256, 0, s>>>(dY, ldY, act_out, ldA, act, M, nbr, nbr_ld, F, (int)h, h_dev, Wt, C, dX,
                                                     ldX);
    EFGH_LAUNCH_CHECK();
    return EFGH_OK;
  });
}

extern "C" int efgh_bcl_conv_wgrad(const float *X, int64_t ldX, int C, const float *row_scale, const void *nbr,
                                   int idx_bits, int64_t nbr_ld, int F, int64_t h, const int32_t *h_dev,
                                   const float *dY, int64_t ldY, const float *act_out, int64_t ldA, int act, int M,
                                   float *dWt, float *dbias, void *stream) {
  EFGH_REQUIRE(C > 0 && M > 0 && h >= 0 && h < (1ll << 30), "efgh_bcl_conv_wgrad: bad sizes");
  if (!nbr) F = 1;
  EFGH_REQUIRE(F >= 1, "efgh_bcl_conv_wgrad: bad filter size %d", F);
  if (h == 0) return EFGH_OK;
  EFGH_REQUIRE(X && dY && dWt, "efgh_bcl_conv_wgrad: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dispatch_idx(idx_bits, [&](auto tag) -> int {
    using IdxT = decltype(tag);
    constexpr int BM = 64, BN = 64;
    int gx = (int)((h + kWgradChunk - 1) / kWgradChunk);
    if (gx > 1024) gx = 1024;
    dim3 grid(gx, F * ((C + BM - 1) / BM), (M + BN - 1) / BN);
    k_wgrad<IdxT, BM, BN, 4, 4><<<grid, 256, 0, s>>>(X, ldX, C, row_scale, nbr, nbr_ld, F, (int)h, h_dev, dY, ldY, act_out,
                                                     ldA, act, M, dWt, dbias);
    EFGH_LAUNCH_CHECK();
    return EFGH_OK;
  });
}

extern "C" int efgh_bcl_act_bwd(float *dX, int64_t ldX, const float *act_out, int64_t ldA, int C, int act, int64_t rows,
                                const int32_t *rows_dev, void *stream) {
  EFGH_REQUIRE(C > 0 && C % 4 == 0 && ldX % 4 == 0 && ldA % 4 == 0 && rows >= 0 && rows < (1ll << 31), "efgh_bcl_act_bwd: bad sizes");
  if (rows == 0 || act == 0) return EFGH_OK;
  EFGH_REQUIRE(dX && act_out && (reinterpret_cast<uintptr_t>(dX) & 15) == 0 && (reinterpret_cast<uintptr_t>(act_out) & 15) == 0,
               "efgh_bcl_act_bwd: null or unaligned pointer");
  k_act_bwd<<<grid_for(rows * (C / 4), 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(dX, ldX, act_out, ldA, C, act, (int)rows,
                                                                                          rows_dev);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

extern "C" int efgh_bcl_loss_half_mean_square(const float *Z, int64_t ldZ, int C, const int32_t *scan_start, int B, float *dZ,
                                              int64_t ldD, float *loss, int64_t rows_cap, void *stream) {
  EFGH_REQUIRE(C > 0 && C % 4 == 0 && ldZ % 4 == 0 && ldD % 4 == 0 && B >= 1 && B <= 64 && rows_cap >= 0 && rows_cap < (1ll << 31),
               "efgh_bcl_loss_half_mean_square: bad sizes");
  EFGH_REQUIRE(Z && dZ && loss && scan_start, "efgh_bcl_loss_half_mean_square: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  EFGH_CUDA_CHECK(cudaMemsetAsync(loss, 0, sizeof(float), s));
  if (rows_cap == 0) return EFGH_OK;
  k_loss_hms<<<grid_for(rows_cap * (C / 4), 256, 4), 256, 0, s>>>(Z, ldZ, C, scan_start, B, dZ, ldD, loss, (int)rows_cap);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}
