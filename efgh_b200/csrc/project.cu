// Point -> image scatter projections and cloud pre-processing: the per-point work that
// surrounds the lattice path in an EFGH forward.
//
//   range image   reference common/torch_utils.py:11-59   (F-Net input): spherical projection of the cloud
//   depth image   reference common/torch_utils.py:61-103  (G-Net input, G loss): pinhole projection with a 3x4 matrix
//   preprocess    reference data_loader/loader_utils.py:163-202: crop to the +-radius box in x / y (order kept),
//                 subsample with a caller-supplied index set (numpy's RNG stays on the host) or zero-pad, rigid
//                 transform in float64
//
// The reference builds both images with `img[u.tolist(), v.tolist()] = values`: a Python-list round trip per sample
// and, for duplicate pixels, "the last point in cloud order wins" (sequential index_put).  Here that rule is explicit
// and deterministic: pass 1 keeps, per pixel, the LARGEST point index that lands on it (atomicMax), pass 2 lets
// exactly that point write its 4 values.
//
// Float semantics (compiled with -fmad=false, every operation an explicit round-to-nearest intrinsic): the float32
// operation order of the reference's torch expressions; asin / atan2 are evaluated in float64 and rounded once, i.e.
// the correctly rounded float32 value (torch's CPU kernels use SLEEF, CUDA's libdevice differs again - DESIGN.md §4
// "projection parity").
#include "common.cuh"

namespace efgh {
namespace {

__global__ void k_image_clear(float *__restrict__ img, int64_t img_elems, int32_t *__restrict__ winner, int64_t pixels) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4 *img4 = reinterpret_cast<float4 *>(img);
  for (int64_t i = tid; i < img_elems / 4; i += stride) img4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (img_elems / 4) * 4 + tid; i < img_elems; i += stride) img[i] = 0.f;
  for (int64_t i = tid; i < pixels; i += stride) winner[i] = -1;
}

struct RangeParams {
  float fov_up, fov_down, fov_span, pi_f, two_pi_f, hm1, wm1;
  int H, W;
};

// pixel of point (x, y, z) or -1
__device__ __forceinline__ int range_pixel(const RangeParams &p, float x, float y, float z, float &r) {
  r = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));   // torch.sqrt(torch.sum(torch.pow(xyz, 2), 1))
  const float pitch = (float)asin((double)__fdiv_rn(z, r));                                   // torch.asin(z_ / r_)
  const float yaw = (float)atan2((double)y, (double)x);                                       // torch.atan2(y_, x_)
  if (!(pitch < p.fov_up && pitch > p.fov_down)) return -1;                                   // :38 (NaN fails both)
  const float u = __fmul_rn(__fdiv_rn(__fsub_rn(p.fov_up, pitch), p.fov_span), p.hm1);       // :48
  const float v = __fmul_rn(__fdiv_rn(__fadd_rn(-yaw, p.pi_f), p.two_pi_f), p.wm1);          // :49
  const int iu = (int)u, iv = (int)v;                                                         // .long(): truncation
  if (iu < 0 || iu >= p.H || iv < 0 || iv >= p.W) return -1;                                  // (cannot happen for finite input)
  return iu * p.W + iv;
}

template <int PASS>
__global__ void __launch_bounds__(256)
k_range_image(const float *__restrict__ pc, int64_t ld, int64_t batch_stride, int n, int B, RangeParams p,
              int32_t *__restrict__ winner, float *__restrict__ img) {
  const int64_t total = (int64_t)B * n;
  const int64_t pixels = (int64_t)p.H * p.W;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / n), i = (int)(t - (int64_t)b * n);
    const float *q = pc + b * batch_stride + i;
    const float x = __ldg(q), y = __ldg(q + ld), z = __ldg(q + 2 * ld);
    float r;
    const int pix = range_pixel(p, x, y, z, r);
    if (pix < 0) continue;
    int32_t *w = winner + b * pixels + pix;
    if (PASS == 0) {
      atomicMax(w, i);
    } else if (*w == i) {
      float *o = img + (int64_t)b * 4 * pixels + pix;
      o[0] = x; o[pixels] = y; o[2 * pixels] = z; o[3 * pixels] = r;
    }
  }
}

struct DepthParams { int H, W; float hf, wf; };

template <int PASS>
__global__ void __launch_bounds__(256)
k_depth_image(const float *__restrict__ pc, int64_t ld, int64_t batch_stride, int n, int B, const float *__restrict__ T,
              DepthParams p, int32_t *__restrict__ winner, float *__restrict__ img) {
  const int64_t total = (int64_t)B * n;
  const int64_t pixels = (int64_t)p.H * p.W;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / n), i = (int)(t - (int64_t)b * n);
    const float *q = pc + b * batch_stride + i;
    const float x = __ldg(q), y = __ldg(q + ld), z = __ldg(q + 2 * ld);
    const float *M = T + b * 12;
    float row[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {                                    // torch.mm(cam_T_velo[b], [pc; 1]): k-ascending FMA chain (sgemm)
      float acc = __fmul_rn(__ldg(M + 4 * r), x);
      acc = __fmaf_rn(__ldg(M + 4 * r + 1), y, acc);
      acc = __fmaf_rn(__ldg(M + 4 * r + 2), z, acc);
      row[r] = __fmaf_rn(__ldg(M + 4 * r + 3), 1.0f, acc);
    }
    const float w = row[2];
    const float px = __fdiv_rn(row[0], w), py = __fdiv_rn(row[1], w);                 // :76-77
    if (!(px < p.wf && px > 0.f && py < p.hf && py > 0.f && w > 0.f)) continue;     // :79
    const int ix = (int)px, iy = (int)py;
    if (ix >= p.W || iy >= p.H) continue;
    const int pix = iy * p.W + ix;
    int32_t *wn = winner + b * pixels + pix;
    if (PASS == 0) {
      atomicMax(wn, i);
    } else if (*wn == i) {
      float *o = img + (int64_t)b * 4 * pixels + pix;
      o[0] = x; o[pixels] = y; o[2 * pixels] = z; o[3 * pixels] = w;                 // :91-96
    }
  }
}

// ---- pre-processing ---------------------------------------------------------------------------------------------------
// pass 1: flags of the crop (loader_utils.py:182-187) -> per-tile counts; pass 2 (single CTA): exclusive scan of the
// tile counts; pass 3: order-preserving compaction into `kept` (indices of the surviving points).  A cloud is at most
// a few hundred thousand points, so a three-launch scan is launch-latency, not bandwidth.
constexpr int kPreTile = 1024;

__device__ __forceinline__ bool crop_keep(const float *__restrict__ p, float radius, int use_radius) {
  if (!use_radius) return true;
  const float x = p[0], y = p[1];
  return x >= -radius && x < radius && y >= -radius && y < radius;
}

__global__ void __launch_bounds__(256)
k_crop_count(const float *__restrict__ pts, int n, float radius, int use_radius, int32_t *__restrict__ tile_cnt) {
  __shared__ int s_cnt[8];
  const int tile = blockIdx.x;
  int c = 0;
  for (int k = threadIdx.x; k < kPreTile; k += 256) {
    const int i = tile * kPreTile + k;
    if (i < n && crop_keep(pts + (int64_t)i * 4, radius, use_radius)) ++c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_cnt[w];
    tile_cnt[tile] = t;
  }
}

__global__ void k_crop_scan(int32_t *__restrict__ tile_cnt, int n_tiles, int32_t *__restrict__ total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < n_tiles; ++t) { const int c = tile_cnt[t]; tile_cnt[t] = run; run += c; }
    *total = run;
  }
}

__global__ void __launch_bounds__(256)
k_crop_compact(const float *__restrict__ pts, int n, float radius, int use_radius, const int32_t *__restrict__ tile_off,
               int32_t *__restrict__ kept) {
  // one warp walks the tile in order, 32 points per step (ballot + popcount keeps the original order)
  const int tile = blockIdx.x;
  if (threadIdx.x >= 32) return;
  int base = tile_off[tile];
  for (int k = 0; k < kPreTile; k += 32) {
    const int i = tile * kPreTile + k + threadIdx.x;
    const bool keep = i < n && crop_keep(pts + (int64_t)i * 4, radius, use_radius);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) kept[base + __popc(m & ((1u << threadIdx.x) - 1u))] = i;
    base += __popc(m);
  }
}

// out[:, j] = T (4x4, float64) @ [x y z 1] of source point j (or of the zero padding), float64 accumulation in the order
// of numpy's matmul for a (4,4) @ (4,N) product (OpenBLAS dgemm: k-ascending fused multiply-adds)
__global__ void __launch_bounds__(256)
k_preproc_gather(const float *__restrict__ pts, const int32_t *__restrict__ kept, const int32_t *__restrict__ kept_total,
                 const int64_t *__restrict__ sample, int n_sample, int num_points, const double *__restrict__ T,
                 double *__restrict__ out64, float *__restrict__ out32, int32_t *__restrict__ status) {
  const int m = *kept_total;
  double t[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) t[k] = T[k];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < num_points; j += gridDim.x * blockDim.x) {
    double v[4] = {0.0, 0.0, 0.0, 1.0};                                 // padding columns: zeros, homogeneous 1 (:195-199)
    if (num_points < m) {                                               // subsample (:189-193): sample[] indexes the CROPPED cloud
      long long sidx = j < n_sample ? sample[j] : -1;
      if (sidx < 0 || sidx >= m) { atomicOr(status, 1); sidx = 0; }
      const float *p = pts + (int64_t)kept[sidx] * 4;
      v[0] = (double)p[0]; v[1] = (double)p[1]; v[2] = (double)p[2];
    } else if (j < m) {
      const float *p = pts + (int64_t)kept[j] * 4;
      v[0] = (double)p[0]; v[1] = (double)p[1]; v[2] = (double)p[2];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double acc = __dmul_rn(t[4 * r], v[0]);                         // dgemm micro-kernel: k-ascending FMA chain
      acc = __fma_rn(t[4 * r + 1], v[1], acc);
      acc = __fma_rn(t[4 * r + 2], v[2], acc);
      acc = __fma_rn(t[4 * r + 3], v[3], acc);
      if (out64) out64[(int64_t)r * num_points + j] = acc;
      if (out32 && r < 3) out32[(int64_t)r * num_points + j] = (float)acc;
    }
  }
}

}  // namespace
}  // namespace efgh

using namespace efgh;

extern "C" int efgh_project_range_image(const float *pc, int64_t pc_ld, int64_t batch_stride, int64_t n, int batch, int height,
                                        int width, double fov_up, double fov_down, int32_t *winner, float *img, void *stream) {
  EFGH_REQUIRE(n >= 0 && n < (1ll << 30) && batch >= 1 && height >= 2 && width >= 2 && (int64_t)height * width < (1ll << 30),
               "efgh_project_range_image: bad sizes");
  EFGH_REQUIRE(winner && img && (n == 0 || pc) && (reinterpret_cast<uintptr_t>(img) & 15) == 0, "efgh_project_range_image: null or unaligned pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t pixels = (int64_t)height * width * batch;
  k_image_clear<<<grid_for(pixels, 256, 8), 256, 0, s>>>(img, pixels * 4, winner, pixels);
  EFGH_LAUNCH_CHECK();
  if (n == 0) return EFGH_OK;
  RangeParams p;
  p.fov_up = (float)fov_up; p.fov_down = (float)fov_down; p.fov_span = (float)(fov_up - fov_down);
  p.pi_f = (float)3.141592653589793; p.two_pi_f = (float)(2.0 * 3.141592653589793);
  p.hm1 = (float)(height - 1); p.wm1 = (float)(width - 1); p.H = height; p.W = width;
  const int grid = grid_for(n * batch, 256, 8);
  k_range_image<0><<<grid, 256, 0, s>>>(pc, pc_ld, batch_stride, (int)n, batch, p, winner, img);
  EFGH_LAUNCH_CHECK();
  k_range_image<1><<<grid, 256, 0, s>>>(pc, pc_ld, batch_stride, (int)n, batch, p, winner, img);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

extern "C" int efgh_project_depth_image(const float *pc, int64_t pc_ld, int64_t batch_stride, int64_t n, int batch,
                                        const float *cam_T_velo, int height, int width, int32_t *winner, float *img, void *stream) {
  EFGH_REQUIRE(n >= 0 && n < (1ll << 30) && batch >= 1 && height >= 1 && width >= 1 && (int64_t)height * width < (1ll << 30),
               "efgh_project_depth_image: bad sizes");
  EFGH_REQUIRE(winner && img && cam_T_velo && (n == 0 || pc) && (reinterpret_cast<uintptr_t>(img) & 15) == 0,
               "efgh_project_depth_image: null or unaligned pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t pixels = (int64_t)height * width * batch;
  k_image_clear<<<grid_for(pixels, 256, 8), 256, 0, s>>>(img, pixels * 4, winner, pixels);
  EFGH_LAUNCH_CHECK();
  if (n == 0) return EFGH_OK;
  DepthParams p = {height, width, (float)height, (float)width};
  const int grid = grid_for(n * batch, 256, 8);
  k_depth_image<0><<<grid, 256, 0, s>>>(pc, pc_ld, batch_stride, (int)n, batch, cam_T_velo, p, winner, img);
  EFGH_LAUNCH_CHECK();
  k_depth_image<1><<<grid, 256, 0, s>>>(pc, pc_ld, batch_stride, (int)n, batch, cam_T_velo, p, winner, img);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}

extern "C" size_t efgh_preproc_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return sizeof(int32_t) * (size_t)(n + (n + kPreTile - 1) / kPreTile + 8);
}

extern "C" int efgh_preproc_cloud(const float *xyzi, int64_t n, int use_radius, float radius, const int64_t *sample,
                                  int64_t n_sample, int64_t num_points, const double *transform, double *out64, float *out32,
                                  int32_t *kept_count, void *workspace, size_t workspace_bytes, void *stream) {
  EFGH_REQUIRE(n >= 0 && n < (1ll << 30) && num_points >= 1 && num_points < (1ll << 30) && n_sample >= 0, "efgh_preproc_cloud: bad sizes");
  EFGH_REQUIRE(transform && kept_count && workspace && (out64 || out32) && (n == 0 || xyzi), "efgh_preproc_cloud: null pointer");
  EFGH_REQUIRE(workspace_bytes >= efgh_preproc_workspace_bytes(n), "efgh_preproc_cloud: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n_tiles = (int)((n + kPreTile - 1) / kPreTile);
  int32_t *kept = static_cast<int32_t *>(workspace);
  int32_t *tile_cnt = kept + (n > 0 ? n : 1);
  int32_t *status = tile_cnt + n_tiles + 1;
  EFGH_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
  if (n_tiles > 0) {
    k_crop_count<<<n_tiles, 256, 0, s>>>(xyzi, (int)n, radius, use_radius, tile_cnt);
    EFGH_LAUNCH_CHECK();
  }
  k_crop_scan<<<1, 32, 0, s>>>(tile_cnt, n_tiles, kept_count);
  EFGH_LAUNCH_CHECK();
  if (n_tiles > 0) {
    k_crop_compact<<<n_tiles, 32, 0, s>>>(xyzi, (int)n, radius, use_radius, tile_cnt, kept);
    EFGH_LAUNCH_CHECK();
  }
  k_preproc_gather<<<grid_for(num_points, 256, 8), 256, 0, s>>>(xyzi, kept, kept_count, sample, (int)n_sample, (int)num_points, transform,
                                                               out64, out32, status);
  EFGH_LAUNCH_CHECK();
  return EFGH_OK;
}
