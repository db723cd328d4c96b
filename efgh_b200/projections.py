"""Point -> image scatter projections on the GPU - drop-ins for reference common/torch_utils.py:11-103.

Same names, arguments and return values as the reference functions, so reference nets/fnet.py:45,
nets/gnet.py:136 and losses/loss_utils.py:183 can import them from here unchanged:

    range_img_from_cartesian_pc_torch(pc, range_img_size, lidar_fov_rad, device) -> (B, 4, H, W)
    depth_img_from_cartesian_pc_torch(pc, cam_T_velo, cam_img_size, device)      -> (B, 4, H, W)

The reference loops over the batch in Python and scatters through `indices.tolist()` (two Python lists of N ints per
sample, a host round trip each); here one launch triple handles the whole batch on the device.  Duplicate pixels are
resolved deterministically: the point with the largest index wins (what the reference's sequential assignment does).
No CPU fallback: CPU tensors are moved to the device named by `device` (as the reference's `.to(device)` does).
"""
import math

import torch

from . import _capi


def _prep(pc, device):
    dev = torch.device("cuda" if device in (None, "cuda") else device)
    if dev.type != "cuda":
        raise _capi.EfghError("efgh_b200.projections runs on CUDA only; there is no CPU path (got device=%r)" % (device,))
    if pc.is_cuda:
        dev = pc.device
    pc = pc.to(device=dev, dtype=torch.float32)
    if pc.dim() != 3 or pc.shape[1] < 3:
        raise ValueError("pc must be (B, 3, N), got %s" % (tuple(pc.shape),))
    if pc.shape[-1] > 1 and pc.stride(-1) != 1:
        pc = pc.contiguous()
    return pc, dev


def range_img_from_cartesian_pc_torch(pc, range_img_size, lidar_fov_rad, device="cuda", return_winner=False):
    """reference common/torch_utils.py:11-59.  pc (B, 3, N); range_img_size (H, W); lidar_fov_rad (up, down) in
    radians / pi.  Returns (B, 4, H, W) float32: x, y, z, range of the point that owns each pixel, 0 elsewhere."""
    pc, dev = _prep(pc, device)
    B, _, n = pc.shape
    H, W = int(range_img_size[0]), int(range_img_size[1])
    fov_up, fov_down = lidar_fov_rad[0] * math.pi, lidar_fov_rad[1] * math.pi          # torch_utils.py:19-20
    with torch.cuda.device(dev):
        img = torch.empty((B, 4, H, W), dtype=torch.float32, device=dev)
        winner = torch.empty((B, H, W), dtype=torch.int32, device=dev)
        _capi.check(_capi.lib().efgh_project_range_image(pc.data_ptr(), pc.stride(1) if n > 1 else max(n, 1), pc.stride(0), n, B, H, W,
                                                         fov_up, fov_down, winner.data_ptr(), img.data_ptr(), _capi.stream_ptr()),
                    "efgh_project_range_image")
    return (img, winner) if return_winner else img


def depth_img_from_cartesian_pc_torch(pc, cam_T_velo, cam_img_size, device="cuda", return_winner=False):
    """reference common/torch_utils.py:61-103.  pc (B, 3, N); cam_T_velo (B, 3, 4); cam_img_size (H, W).
    Returns (B, 4, H, W) float32: x, y, z of the point that owns each pixel and its projective depth w."""
    pc, dev = _prep(pc, device)
    B, _, n = pc.shape
    H, W = int(cam_img_size[0]), int(cam_img_size[1])
    T = cam_T_velo.to(device=dev, dtype=torch.float32).contiguous()
    if tuple(T.shape) != (B, 3, 4):
        raise ValueError("cam_T_velo must be (B, 3, 4), got %s" % (tuple(T.shape),))
    with torch.cuda.device(dev):
        img = torch.empty((B, 4, H, W), dtype=torch.float32, device=dev)
        winner = torch.empty((B, H, W), dtype=torch.int32, device=dev)
        _capi.check(_capi.lib().efgh_project_depth_image(pc.data_ptr(), pc.stride(1) if n > 1 else max(n, 1), pc.stride(0), n, B,
                                                         T.data_ptr(), H, W, winner.data_ptr(), img.data_ptr(), _capi.stream_ptr()),
                    "efgh_project_depth_image")
    return (img, winner) if return_winner else img
