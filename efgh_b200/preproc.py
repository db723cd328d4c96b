"""Cloud pre-processing on the GPU - drop-in for reference data_loader/loader_utils.py:59-61,163-202.

    pcd_read(filename)                       -> (n, 4) float32 numpy array (the `.bin` wire format: x, y, z, intensity)
    preproc_pcd(pcd, gts, num_points, lidar_line=None, radius=50., sample=None, device="cuda")
                                             -> (4, num_points) float64 CUDA tensor (the reference returns numpy)

`preproc_pcd` crops to the +-radius box in x / y (order kept), subsamples to `num_points` or zero-pads, and applies
the rigid perturbation gts['rand_init_l'] (4x4 float64).  The subsample is the one non-deterministic step of the
reference (`np.random.choice(range(m), num_points, replace=False)`); here the index set is drawn on the host with the
same call unless the caller passes `sample`, so results are reproducible against the reference under the same
numpy RNG state.  The reference's `reduce_lidar_line` (a Python loop that re-slices KITTI's 64 rings) is NOT covered:
`lidar_line` other than None raises.
"""
import numpy as np
import torch

from . import _capi


def pcd_read(filename):
    """reference data_loader/loader_utils.py:59-61"""
    scan = np.fromfile(filename, dtype=np.float32)
    return scan.reshape((-1, 4))


def preproc_pcd(pcd, gts, num_points, lidar_line=None, radius=50., sample=None, device="cuda", return_float32=False):
    if lidar_line is not None:
        raise NotImplementedError("efgh_b200.preproc: reduce_lidar_line (loader_utils.py:165-180) is host-side ring slicing; "
                                  "apply it before calling (or pass lidar_line=None)")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _capi.EfghError("efgh_b200.preproc runs on CUDA only; there is no CPU path")
    L = _capi.lib()
    with torch.cuda.device(dev):
        pts = torch.as_tensor(pcd, dtype=torch.float32).to(dev).contiguous()
        if pts.dim() != 2 or pts.shape[1] != 4:
            raise ValueError("pcd must be (n, 4) float32 x, y, z, intensity records, got %s" % (tuple(pts.shape),))
        n = pts.shape[0]
        T = torch.as_tensor(np.asarray(gts['rand_init_l'], dtype=np.float64)).to(dev).contiguous()
        ws = torch.empty(max(int(L.efgh_preproc_workspace_bytes(n)), 16), dtype=torch.uint8, device=dev)
        kept = torch.zeros((1,), dtype=torch.int32, device=dev)
        out64 = torch.empty((4, num_points), dtype=torch.float64, device=dev)
        out32 = torch.empty((3, num_points), dtype=torch.float32, device=dev) if return_float32 else None
        stream = _capi.stream_ptr()

        def run(smp):
            _capi.check(L.efgh_preproc_cloud(pts.data_ptr(), n, 0 if radius is None else 1, float(radius or 0.0), _capi.ptr(smp),
                                             0 if smp is None else smp.numel(), num_points, T.data_ptr(), out64.data_ptr(),
                                             _capi.ptr(out32), kept.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                        "efgh_preproc_cloud")
        if sample is None and num_points >= n:
            run(None)                                          # cannot need a subsample: one pass, no host sync
        else:
            if sample is None:
                # the cropped size decides whether a subsample is needed, and np.random.choice needs it: one small D2H read
                run(torch.zeros((num_points,), dtype=torch.int64, device=dev))
                m = int(kept.item())
                if num_points < m:
                    sample = np.random.choice(range(m), size=num_points, replace=False, p=None)      # loader_utils.py:190-192
            if sample is not None:
                run(torch.as_tensor(np.asarray(sample, dtype=np.int64)).to(dev))
    return (out64, out32) if return_float32 else out64
