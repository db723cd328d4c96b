"""numpy restatement of the point -> image projections and of the cloud pre-processing - TEST INFRASTRUCTURE ONLY.

  range_image   reference common/torch_utils.py:11-59   (range_img_from_cartesian_pc_torch)
  depth_image   reference common/torch_utils.py:61-103  (depth_img_from_cartesian_pc_torch)
  preproc_pcd   reference data_loader/loader_utils.py:163-202 (without reduce_lidar_line; the subsample index set is an argument)

Every float32 operation is written in the order torch evaluates the reference's expressions (Python scalars are cast to
float32 first, `/` is a true division, `sum(pow(xyz, 2), 1)` adds x^2 + y^2 + z^2 left to right - all checked against
torch 2.11 CPU).  Two deliberate definitions where the reference itself is not well defined across platforms:
  * asin / atan2 are evaluated in float64 and rounded once (the correctly rounded float32 result).  torch-CPU (SLEEF)
    and CUDA libdevice each differ from that in the last bit for a few per cent of the inputs, so a pixel index that
    sits within one ulp of an integer can move by one pixel between ANY two of the three; the square root is the IEEE
    one (np.sqrt, CUDA sqrt.rn) - torch's CPU sqrt (SLEEF u05) is off by one ulp for 0.7 % of the values;
  * duplicate pixels: the last point in cloud order wins (numpy's and torch-CPU's sequential assignment; the
    reference's CUDA index_put_ is non-deterministic there).
Parity status: PINNED against the live reference functions run on CPU in the build container
(tests/golden/proj_*.npz, preproc_*.npz, produced by oracle/make_golden.py): depth image and pre-processing
bit-exact, range image bit-exact except pixels within one float32 ulp of a pixel border (counted in the test).
"""
import math

import numpy as np

f32 = np.float32


def range_pixels(pc, range_img_size, lidar_fov_rad):
    """pc (3, N) float32 -> (mask (N,) bool, u (N,) int64, v (N,) int64, r (N,) float32)"""
    fov_up, fov_down = lidar_fov_rad[0] * math.pi, lidar_fov_rad[1] * math.pi
    x, y, z = (np.asarray(pc[i], dtype=f32) for i in range(3))
    with np.errstate(all="ignore"):
        r = np.sqrt((x * x + y * y) + z * z, dtype=f32)
        pitch = np.arcsin((z / r).astype(np.float64)).astype(f32)
        yaw = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(f32)
        mask = (pitch < f32(fov_up)) & (pitch > f32(fov_down))
        u = ((f32(fov_up) - pitch) / f32(fov_up - fov_down)) * f32(range_img_size[0] - 1)
        v = ((-yaw + f32(math.pi)) / f32(2 * math.pi)) * f32(range_img_size[1] - 1)
        u = np.where(mask, u, 0).astype(np.int64)
        v = np.where(mask, v, 0).astype(np.int64)
    return mask, u, v, r


def range_image(pc, range_img_size, lidar_fov_rad):
    """pc (B, 3, N) float32 -> (B, 4, H, W) float32"""
    pc = np.asarray(pc, dtype=f32)
    H, W = range_img_size
    out = np.zeros((pc.shape[0], 4, H, W), f32)
    for b in range(pc.shape[0]):
        mask, u, v, r = range_pixels(pc[b], range_img_size, lidar_fov_rad)
        vals = np.stack([pc[b, 0], pc[b, 1], pc[b, 2], r], 1)[mask]
        img = np.zeros((H, W, 4), f32)
        img[u[mask], v[mask]] = vals                       # repeated indices: the last assignment stays
        out[b] = img.transpose(2, 0, 1)
    return out


def _fma32(a, b, c):
    """float32 fused multiply-add (one rounding): exact in float64 up to the final rounding for float32 inputs
    whenever no double rounding occurs - a float32 product is exact in float64, the sum rounds once to float64 and
    once to float32; inputs here never hit the 2^-29 double-rounding window in the golden fixtures (checked there)."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(f32)


def depth_image(pc, cam_T_velo, cam_img_size):
    """pc (B, 3, N) float32, cam_T_velo (B, 3, 4) float32 -> (B, 4, H, W) float32"""
    pc = np.asarray(pc, dtype=f32)
    T = np.asarray(cam_T_velo, dtype=f32)
    H, W = cam_img_size
    out = np.zeros((pc.shape[0], 4, H, W), f32)
    for b in range(pc.shape[0]):
        x, y, z = pc[b, 0], pc[b, 1], pc[b, 2]
        rows = []
        for r in range(3):                                   # sgemm: k-ascending fused multiply-add chain
            acc = (T[b, r, 0] * x).astype(f32)
            acc = _fma32(y, T[b, r, 1], acc)
            acc = _fma32(z, T[b, r, 2], acc)
            acc = _fma32(np.ones_like(x), T[b, r, 3], acc)
            rows.append(acc)
        with np.errstate(all="ignore"):
            w = rows[2]
            px, py = rows[0] / w, rows[1] / w
            mask = (px < f32(W)) & (px > 0) & (py < f32(H)) & (py > 0) & (w > 0)
        iy, ix = py[mask].astype(np.int64), px[mask].astype(np.int64)
        img = np.zeros((H, W, 4), f32)
        img[iy, ix] = np.stack([x, y, z, w], 1)[mask]
        out[b] = img.transpose(2, 0, 1)
    return out


def preproc_pcd(pcd, transform, num_points, radius=50.0, sample=None):
    """pcd (n, 4) float32; transform (4, 4) float64; sample: indices into the cropped cloud (required when it has
    more than num_points points).  Returns ((4, num_points) float64, cropped size)."""
    pcd = np.asarray(pcd, dtype=f32)
    if radius is not None:
        keep = (pcd[:, 0] >= -radius) & (pcd[:, 0] < radius) & (pcd[:, 1] >= -radius) & (pcd[:, 1] < radius)
        pcd = pcd[np.where(keep)[0]]
    m = pcd.shape[0]
    if num_points < m:
        pcd_ = pcd[np.asarray(sample)].T
    else:
        pcd_ = np.zeros((3, num_points))
        pcd_[:3, :m] = pcd[:, :3].T
    pc = np.ones((4, pcd_.shape[1]))
    pc[:3, :] = pcd_[:3, :]
    return np.asarray(transform, dtype=np.float64) @ pc, m
