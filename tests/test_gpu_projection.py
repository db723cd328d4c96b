"""GPU (-m gpu): point -> image scatter projections and cloud pre-processing (row f4), through
the C ABI, against the numpy oracle and the fixtures generated from the live reference."""
import math
import os

import numpy as np
import pytest
import torch

from efgh_b200 import synth

pytestmark = pytest.mark.gpu

FOV = (0.125, -0.125)          # reference configs/train_rellis.yaml:21


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from efgh_b200 import _capi
    _capi.lib()
    return torch.device("cuda:0")


def _bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def test_range_image_vs_oracle_and_reference_golden(dev, golden_dir):
    from efgh_b200.projections import range_img_from_cartesian_pc_torch
    from oracle import projection as op
    z = np.load(os.path.join(golden_dir, "proj_range.npz"))
    size, fov = tuple(int(v) for v in z["size"]), tuple(float(v) for v in z["fov"])
    got = range_img_from_cartesian_pc_torch(torch.from_numpy(z["pc"]).to(dev), size, fov, "cuda").cpu().numpy()
    want = op.range_image(z["pc"], size, fov)
    assert got.shape == want.shape == z["img"].shape
    assert np.array_equal(_bits(got), _bits(want))                    # bit-exact against the oracle
    # against the live reference (torch CPU): same owners except border points, range within one ulp (test_oracle_golden)
    owner_differs = (_bits(got[:, :3]) != _bits(z["img"][:, :3])).any(axis=1)
    assert owner_differs.sum() <= 4


@pytest.mark.parametrize("sensor,size", [("os1-64", (600, 3840)), ("hdl-64", (64, 2048))])
def test_range_image_fullsize_vs_oracle(sensor, size, dev):
    """BASELINE configs[2] shape: a full 131k-point sweep into the (600, 3840) range image of reference nets/fnet.py:45
    (range_img_size = (raw_cam_img_size[0] / 2, raw_cam_img_size[1] * 2), reference common/numpy_utils.py:17)."""
    from efgh_b200.projections import range_img_from_cartesian_pc_torch
    from oracle import projection as op
    pc = np.stack([synth.synth_scan(s, sensor) for s in (3, 4)], 0)
    img, winner = range_img_from_cartesian_pc_torch(torch.from_numpy(pc).to(dev), size, FOV, "cuda", return_winner=True)
    want = op.range_image(pc, size, FOV)
    got = img.cpu().numpy()
    assert np.array_equal(_bits(got), _bits(want))
    # properties: every set pixel holds a point of the cloud, its own range, and the winner is the LAST point on that pixel
    w = winner.cpu().numpy()
    for b in range(2):
        mask, u, v, r = op.range_pixels(pc[b], size, FOV)
        idx = np.nonzero(mask)[0]
        last = np.full(size, -1, np.int64)
        last[u[idx], v[idx]] = idx                                   # sequential: the last assignment stays
        assert np.array_equal(w[b], last)
        assert (got[b, 3] > 0).sum() == (last >= 0).sum()


def test_depth_image_vs_reference_golden(dev, golden_dir):
    from efgh_b200.projections import depth_img_from_cartesian_pc_torch
    z = np.load(os.path.join(golden_dir, "proj_depth.npz"))
    size = tuple(int(v) for v in z["size"])
    got = depth_img_from_cartesian_pc_torch(torch.from_numpy(z["pc"]).to(dev), torch.from_numpy(z["T"]).to(dev), size, "cuda").cpu().numpy()
    assert got.shape == z["img"].shape
    assert np.array_equal(_bits(got), _bits(z["img"]))               # bit-exact against the LIVE reference's output


def test_depth_image_fullsize_vs_oracle(dev):
    """131k-point sweep into the 1200 x 1920 camera image (reference configs/train_rellis.yaml raw_cam_img_size)."""
    from efgh_b200.projections import depth_img_from_cartesian_pc_torch
    from oracle import projection as op
    pc = synth.synth_scan(8, "os1-64")[None]
    H, W = 1200, 1920
    K = np.array([[2813.6, 0, 969.3], [0, 2808.3, 624.0], [0, 0, 1.0]])          # RELLIS-3D-like intrinsics
    R0 = np.array([[0, -1.0, 0], [0, 0, -1.0], [1.0, 0, 0]])
    T = (K @ np.concatenate([R0, np.array([[0.03], [-0.1], [-0.12]])], 1)).astype(np.float32)[None]
    got = depth_img_from_cartesian_pc_torch(torch.from_numpy(pc).to(dev), torch.from_numpy(T).to(dev), (H, W), "cuda").cpu().numpy()
    want = op.depth_image(pc, T, (H, W))
    assert (want[0, 3] != 0).sum() > 3000
    assert np.array_equal(_bits(got), _bits(want))


def test_projection_duplicates_and_degenerate_inputs(dev):
    """Points on ONE ray at different ranges share a pixel: the last one in cloud order owns it.  An empty cloud, a
    zero point (r = 0 -> NaN pitch) and points outside the FoV / behind the camera leave the image untouched."""
    from efgh_b200.projections import range_img_from_cartesian_pc_torch, depth_img_from_cartesian_pc_torch
    ray = np.array([math.cos(0.1) * math.cos(0.7), math.cos(0.1) * math.sin(0.7), math.sin(0.1)], np.float32)
    pc = (ray[:, None] * np.array([5.0, 9.0, 7.0, 3.0], np.float32)[None, :])[None]           # 4 points, same direction
    img, win = range_img_from_cartesian_pc_torch(torch.from_numpy(pc).to(dev), (16, 64), FOV, "cuda", return_winner=True)
    assert int((win >= 0).sum()) == 1 and int(win.max()) == 3
    assert abs(float(img[0, 3].max()) - 3.0) < 1e-5
    empty = range_img_from_cartesian_pc_torch(torch.zeros(2, 3, 0, device=dev), (16, 64), FOV, "cuda")
    assert tuple(empty.shape) == (2, 4, 16, 64) and float(empty.abs().sum()) == 0
    bad = torch.tensor([[[0.0, 0.0, 1.0], [0.0, 0.0, 0.0], [0.0, 5.0, -5.0]]], device=dev)     # zero point, straight up, straight down
    assert float(range_img_from_cartesian_pc_torch(bad, (16, 64), FOV, "cuda").abs().sum()) == 0
    T = torch.tensor([[[100.0, 0, 32, 0], [0, 100.0, 16, 0], [0, 0, 1.0, 0]]], device=dev)     # camera looks along +z
    behind = torch.tensor([[[0.1, 0.1], [0.1, 0.1], [-2.0, 2.0]]], device=dev)                  # first point behind the camera
    d = depth_img_from_cartesian_pc_torch(behind, T, (32, 64), "cuda")
    assert int((d[0, 3] != 0).sum()) == 1 and float(d[0, 3].max()) == 2.0


@pytest.mark.parametrize("name", ["subsample", "pad"])
def test_preproc_vs_reference_golden(name, dev, golden_dir):
    from efgh_b200.preproc import preproc_pcd
    z = np.load(os.path.join(golden_dir, "preproc_%s.npz" % name))
    sample = z["sample"] if z["sample"].size else None
    out64, out32 = preproc_pcd(z["scan"], {"rand_init_l": z["T"]}, int(z["num_points"]), sample=sample, return_float32=True)
    got = out64.cpu().numpy()
    assert got.shape == z["out"].shape
    # float64: the fused multiply-add chain of the dgemm micro-kernel reproduces numpy to the last bit here
    assert np.array_equal(got, z["out"])
    assert np.array_equal(out32.cpu().numpy(), z["out"][:3].astype(np.float32))     # what the network consumes after .float()


def test_preproc_bin_roundtrip_and_rng_replay(dev, tmp_path):
    """`.bin` wire format (reference loader_utils.py:59-61) -> preproc_pcd with the subsample drawn by numpy's global RNG
    exactly as the reference draws it (same seed => same cloud as the oracle fed with the replayed index set)."""
    from efgh_b200.preproc import preproc_pcd, pcd_read
    from oracle import projection as op
    scan = np.concatenate([synth.synth_scan(17, "os1-64").T * 1.3, np.random.default_rng(1).uniform(0, 1, (131072, 1)).astype(np.float32)], 1)
    scan = np.ascontiguousarray(scan.astype(np.float32))
    path = os.path.join(tmp_path, "000000.bin")
    scan.tofile(path)
    pcd = pcd_read(path)
    assert pcd.shape == (131072, 4) and np.array_equal(pcd, scan)
    Tl = np.eye(4)
    Tl[:3, 3] = [0.5, -1.0, 0.25]
    np.random.seed(99)
    state = np.random.get_state()
    got = preproc_pcd(pcd, {"rand_init_l": Tl}, 65536).cpu().numpy()          # reference configs/train_rellis.yaml:19 num_points
    keep = (scan[:, 0] >= -50.) & (scan[:, 0] < 50.) & (scan[:, 1] >= -50.) & (scan[:, 1] < 50.)
    m = int(keep.sum())
    assert 65536 < m < 131072
    np.random.set_state(state)
    sample = np.random.choice(range(m), size=65536, replace=False, p=None)
    want, m2 = op.preproc_pcd(scan, Tl, 65536, sample=sample)
    assert m2 == m and np.array_equal(got, want)
    # no crop, padding branch
    got2 = preproc_pcd(pcd[:1000], {"rand_init_l": Tl}, 4096, radius=None).cpu().numpy()
    want2, _ = op.preproc_pcd(scan[:1000], Tl, 4096, radius=None)
    assert np.array_equal(got2, want2)
