"""Synthetic LiDAR sweeps shaped like the sensors EFGHNet is trained on.

There are no datasets in the build/bench environment, so every test and bench line uses these clouds.
The generator mirrors what the reference's loaders hand to the network: a float32 (3, N) cloud, cropped
to the 50 m box (reference data_loader/loader_utils.py:163,182-187) and with the points randomly
permuted (reference data_loader/rellis3d_loader.py:251-252) - the permutation matters because lattice
indices are assigned in first-occurrence order.
"""
import numpy as np

SENSORS = {
    # name: (beams, azimuth steps, fov_up_deg, fov_down_deg)
    "os1-64": (64, 2048, 22.5, -22.5),      # RELLIS-3D Ouster OS1-64, 131 072 pts
    "os1-64-16k": (64, 256, 22.5, -22.5),   # config 1 of BASELINE.json, 16 384 pts
    "os1-64-64k": (64, 1024, 22.5, -22.5),  # 65 536 pts (the shipped num_points)
    "hdl-64": (64, 1920, 2.0, -24.8),       # KITTI HDL-64-like, 122 880 pts
    "nusc-32": (32, 1088, 10.0, -30.0),     # nuScenes-32-like, 34 816 pts
}


def synth_scan(seed=0, sensor="os1-64", max_range=50.0, sensor_height=1.5, noise=0.02,
               rpy_deg=None, trans=None):
    """Returns a float32 (3, N) cloud. seed = scan index (numpy default_rng)."""
    beams, steps, up, down = SENSORS[sensor] if isinstance(sensor, str) else sensor
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(up, down, beams))[:, None]
    az = np.linspace(-np.pi, np.pi, steps, endpoint=False)[None, :]
    el = np.broadcast_to(el, (beams, steps))
    az = np.broadcast_to(az, (beams, steps))
    r_obst = rng.uniform(3.0, max_range, size=(beams, steps))
    with np.errstate(divide="ignore"):
        r_ground = np.where(el < 0, -sensor_height / np.sin(np.minimum(el, -1e-9)), np.inf)
    r = np.minimum(r_obst, r_ground) + rng.normal(0.0, noise, size=(beams, steps))
    r = np.clip(r, 0.5, max_range)
    xyz = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], 0)
    xyz = xyz.reshape(3, -1)
    if rpy_deg is not None or trans is not None:
        rr, pp, yy = np.deg2rad(rpy_deg if rpy_deg is not None else (0, 0, 0))
        rx = np.array([[1, 0, 0], [0, np.cos(rr), -np.sin(rr)], [0, np.sin(rr), np.cos(rr)]])
        ry = np.array([[np.cos(pp), 0, np.sin(pp)], [0, 1, 0], [-np.sin(pp), 0, np.cos(pp)]])
        rz = np.array([[np.cos(yy), -np.sin(yy), 0], [np.sin(yy), np.cos(yy), 0], [0, 0, 1]])
        xyz = rz @ ry @ rx @ xyz
        if trans is not None:
            xyz = xyz + np.asarray(trans, dtype=np.float64)[:, None]
    xyz = np.clip(xyz, -max_range, max_range)
    perm = rng.permutation(xyz.shape[1])
    return np.ascontiguousarray(xyz[:, perm]).astype(np.float32)


# The shipped RELLIS-3D config (reference configs/train_rellis.yaml:28-35): dim 3, five scales, radius 1.
SCALE_MAP = [[1.0, 1], [0.75, 1], [0.5, 1], [0.25, 1], [0.125, 1]]
# E-Net BCL channel plan (reference nets/enet.py:30-83): C_in, [C_mid, C_out] per level.
ENET_BCL = [(36, [32, 32]), (36, [64, 64]), (68, [128, 128]), (132, [256, 256]), (260, [256, 256])]
