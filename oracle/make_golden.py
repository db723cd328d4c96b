"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python -m oracle.make_golden

Lattice goldens come from reference GenerateData.__call__ (nets/generate_data.py:117-193) and
get_keys_and_barycentric (:56-112); BCL goldens from reference BilateralConvFlex (nets/bilateralNN.py)
forward + autograd backward on CPU.  Small cases are stored in full; full-size clouds are stored as
per-array SHA-256 digests plus the vertex counts (a checksum of checksums), so the committed fixtures
stay small while still pinning the 16k / 131k behaviour.
"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness  # noqa: E402
from efgh_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def lattice_cases():
    rng = np.random.default_rng(1234)
    full = synth.synth_scan(7, "os1-64-16k")
    cases = {
        "sub2048": (full[:, :2048], synth.SCALE_MAP),
        "ragged257": (full[:, 5000:5257], synth.SCALE_MAP),
        "single": (full[:, :1], synth.SCALE_MAP[:2]),
        "two": (full[:, :2], synth.SCALE_MAP[:3]),
        # exact ties in el_minus_gr / points on lattice vertices / half-way rounding
        "ties": (np.concatenate([np.zeros((3, 4), np.float32),
                                 rng.integers(-6, 7, size=(3, 120)).astype(np.float32),
                                 (rng.integers(-40, 41, size=(3, 132)) * 0.25).astype(np.float32)], 1),
                 synth.SCALE_MAP[:3]),
        "gauss": ((rng.standard_normal((3, 700)) * 8).astype(np.float32), [[1.0, 1], [0.5, 2], [0.25, 1]]),
        "noblur": (full[:, 100:400], [[1.0, 1], [0.5, -1], [0.25, 1]]),
        "upscale": (full[:, 300:600], [[0.5, 1], [1.0, 1], [2.0, 1]]),
        "negquad": (-np.abs(full[:, 600:900]), synth.SCALE_MAP[:2]),
    }
    return cases


def run_lattice(g, pc, scale_map):
    gd = g.GenerateData(3, scale_map, "cpu")
    keys, bary, elmgr = gd.get_keys_and_barycentric(torch.from_numpy(pc.copy()))
    _, data = gd(torch.from_numpy(pc.copy()))
    return keys, data


def main():
    os.makedirs(OUT, exist_ok=True)
    t, g, b = ref_harness.load()
    # ---- lattice, small cases stored in full
    for name, (pc, smap) in lattice_cases().items():
        keys, data = run_lattice(g, pc, smap)
        blob = {"pc": pc, "scale_map": np.asarray(smap, np.float64), "keys0": keys.astype(np.int32)}
        for li, d in enumerate(data):
            blob["L%d_bary" % li] = d["pc1_barycentric"].numpy()
            blob["L%d_elmgr" % li] = d["pc1_el_minus_gr"].numpy()
            blob["L%d_off" % li] = d["pc1_lattice_offset"].numpy().astype(np.int32)
            blob["L%d_nbr" % li] = d["pc1_blur_neighbors"].numpy().astype(np.int32)
            blob["L%d_cnt" % li] = np.int64(d["pc1_hash_cnt"])
        np.savez_compressed(os.path.join(OUT, "lattice_%s.npz" % name), **blob)
        print("lattice", name, pc.shape, [d["pc1_hash_cnt"] for d in data])
    # ---- lattice, full-size clouds as digests (arrays hashed in the reference's dtypes/layouts)
    dig = {}
    for sensor, seed in (("os1-64-16k", 0), ("os1-64-16k", 3), ("os1-64-64k", 0), ("os1-64", 0),
                         ("os1-64", 1), ("hdl-64", 0), ("nusc-32", 0)):
        pc = synth.synth_scan(seed, sensor)
        _, data = run_lattice(g, pc, synth.SCALE_MAP)
        key = "%s/seed%d" % (sensor, seed)
        dig[key + "/pc"] = digest(pc)
        dig[key + "/cnt"] = ",".join(str(d["pc1_hash_cnt"]) for d in data)
        for li, d in enumerate(data):
            for k in ("pc1_barycentric", "pc1_el_minus_gr", "pc1_lattice_offset", "pc1_blur_neighbors"):
                dig["%s/L%d/%s" % (key, li, k)] = digest(d[k].numpy())
        print("digest", key, dig[key + "/cnt"])
    with open(os.path.join(OUT, "lattice_digests.txt"), "w") as f:
        for k in sorted(dig):
            f.write("%s %s\n" % (k, dig[k]))
    # ---- BCL forward/backward on small lattices
    torch.manual_seed(0)
    pc = synth.synth_scan(11, "os1-64-16k")[:, :1500]
    gd = g.GenerateData(3, [[1.0, 1], [0.5, 1]], "cpu")
    _, data = gd(torch.from_numpy(pc.copy()))
    d0, d1 = data
    N, H = pc.shape[1], d0["pc1_hash_cnt"]
    for name, kw in {
        "splat_noslice": dict(num_input=12, num_output=[10, 7], do_splat=True, do_slice=False, use_norm=True, last_relu=False),
        "splat_slice_bias": dict(num_input=8, num_output=[16, 9], do_splat=True, do_slice=True, use_norm=True, last_relu=True),
        "nonorm_single": dict(num_input=5, num_output=[6], do_splat=True, do_slice=False, use_norm=False, last_relu=False),
        "nosplat_three": dict(num_input=6, num_output=[8, 8, 4], do_splat=False, do_slice=True, use_norm=True, last_relu=True),
        "enet_l0": dict(num_input=36, num_output=[32, 32], do_splat=True, do_slice=False, use_norm=True, last_relu=False),
    }.items():
        use_leaky = name != "nosplat_three"
        m = b.BilateralConvFlex(3, 1, kw["num_input"], kw["num_output"], "cpu", use_bias=True,
                                use_leaky=use_leaky, use_norm=kw["use_norm"], do_splat=kw["do_splat"],
                                do_slice=kw["do_slice"], last_relu=kw["last_relu"], chunk_size=-1)
        for p in m.parameters():
            torch.nn.init.normal_(p, 0, 0.2)
        n_in = N if kw["do_splat"] else H
        feat = torch.randn(1, kw["num_input"], n_in, requires_grad=True)
        out = m(feat, d0["pc1_barycentric"], d0["pc1_lattice_offset"], d0["pc1_blur_neighbors"],
                d0["pc1_barycentric"] if kw["do_slice"] else None,
                d0["pc1_lattice_offset"] if kw["do_slice"] else None)
        gout = torch.randn_like(out)
        out.backward(gout)
        blob = {"pc": pc, "feat": feat.detach().numpy(), "out": out.detach().numpy(), "gout": gout.numpy(),
                "gfeat": feat.grad.numpy(), "bary": d0["pc1_barycentric"].numpy(),
                "off": d0["pc1_lattice_offset"].numpy().astype(np.int32),
                "nbr": d0["pc1_blur_neighbors"].numpy().astype(np.int32),
                "cfg": np.array([kw["num_input"], int(kw["do_splat"]), int(kw["do_slice"]), int(kw["use_norm"]),
                                 int(kw["last_relu"]), int(use_leaky)] + kw["num_output"], np.int64)}
        for k, p in m.named_parameters():
            blob["p_" + k] = p.detach().numpy()
            blob["g_" + k] = p.grad.numpy()
        np.savez_compressed(os.path.join(OUT, "bcl_%s.npz" % name), **blob)
        print("bcl", name, tuple(out.shape), float(out.abs().mean()))


def stem_golden():
    """E-Net stem (reference nets/enet.py:24-28,111: three conv_1x1 of nets/net_utils.py:35-43 on the unscaled xyz)
    -> tests/golden/stem_{leaky,relu}.npz: weights, input cloud, output of the LIVE reference modules on CPU."""
    import torch.nn as nn
    nu = ref_harness.load_net_utils()
    os.makedirs(OUT, exist_ok=True)
    pc = synth.synth_scan(12, "os1-64-16k")[:, :700]
    for name, leaky in (("leaky", True), ("relu", False)):
        torch.manual_seed(7 if leaky else 8)
        stem = nn.Sequential(nu.conv_1x1(3, 32, use_leaky=leaky), nu.conv_1x1(32, 32, use_leaky=leaky),
                             nu.conv_1x1(32, 32, use_leaky=leaky))
        for prm in stem.parameters():                    # the reference's N(0, 1e-3) init gives ~0 outputs
            torch.nn.init.normal_(prm, 0, 0.5)
        with torch.no_grad():
            out = stem(torch.from_numpy(pc.copy())[None])[0]
        blob = {"pc": pc, "out": out.numpy(), "leaky": np.int64(leaky)}
        for li in range(3):
            blob["W%d" % li] = stem[li][0].weight.detach().numpy()
            blob["b%d" % li] = stem[li][0].bias.detach().numpy()
        np.savez_compressed(os.path.join(OUT, "stem_%s.npz" % name), **blob)
        print("stem", name, tuple(out.shape), float(out.abs().mean()))


def projection_cases():
    """Clouds / calibrations of the f3 / f4 fixtures (shared with the tests through the npz files)."""
    rng = np.random.default_rng(77)
    full = synth.synth_scan(13, "os1-64-16k")
    pc = np.stack([full[:, :6000], synth.synth_scan(14, "os1-64-16k")[:, :6000]], 0)          # (2, 3, 6000)
    pc[1, :, :40] = pc[1, :, 40:80]                     # exact duplicates: same pixel, the later point must win
    pc[0, :, 100] = 0.0                                 # r = 0 -> NaN pitch -> masked out
    # a pinhole camera looking along +x (RELLIS-like intrinsics scaled to the test image), small rotation + offset
    Hc, Wc = 150, 240
    K = np.array([[170.0, 0, Wc / 2.0], [0, 170.0, Hc / 2.0], [0, 0, 1.0]])
    R0 = np.array([[0, -1.0, 0], [0, 0, -1.0], [1.0, 0, 0]])                                   # velo (x fwd, y left, z up) -> cam (x right, y down, z fwd)
    Ts = []
    for b in range(2):
        from scipy.spatial.transform import Rotation
        R = Rotation.from_euler("xyz", rng.uniform(-0.05, 0.05, 3)).as_matrix() @ R0
        t = rng.uniform(-0.2, 0.2, 3)
        Ts.append(K @ np.concatenate([R, t[:, None]], 1))
    return pc.astype(np.float32), np.stack(Ts, 0).astype(np.float32), (Hc, Wc)


def projection_golden():
    """f3: reference common/torch_utils.py:11-103 run LIVE on CPU (single thread: sequential duplicate resolution)
    -> tests/golden/proj_range.npz, proj_depth.npz; f4: data_loader/loader_utils.py:163-202 -> preproc_*.npz."""
    c = ref_harness.load_common()
    tu, lu = c["torch_utils"], c["loader_utils"]
    os.makedirs(OUT, exist_ok=True)
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        pc, T, (Hc, Wc) = projection_cases()
        size, fov = (32, 512), (0.125, -0.125)          # reference configs/train_rellis.yaml:21 lidar_fov_rad
        rimg = tu.range_img_from_cartesian_pc_torch(torch.from_numpy(pc.copy()), size, fov, "cpu").numpy()
        np.savez_compressed(os.path.join(OUT, "proj_range.npz"), pc=pc, size=np.asarray(size), fov=np.asarray(fov), img=rimg)
        print("range image", rimg.shape, int((rimg[:, 3] > 0).sum()), "pixels set")
        dimg = tu.depth_img_from_cartesian_pc_torch(torch.from_numpy(pc.copy()), torch.from_numpy(T.copy()), (Hc, Wc), "cpu").numpy()
        np.savez_compressed(os.path.join(OUT, "proj_depth.npz"), pc=pc, T=T, size=np.asarray((Hc, Wc)), img=dimg)
        print("depth image", dimg.shape, int((dimg[:, 3] != 0).sum()), "pixels set")
    finally:
        torch.set_num_threads(nthreads)
    # ---- pre-processing: a `.bin`-format scan (n, 4) float32 with points outside the 50 m box; subsample / pad cases
    rng = np.random.default_rng(5)
    scan = np.concatenate([synth.synth_scan(15, "os1-64-16k")[:, :5000].T, rng.uniform(0, 1, (5000, 1)).astype(np.float32)], 1)
    scan[::7, 0] *= 1.6                                  # push every 7th point out of the +-50 m box in x
    scan[3::11, 1] = -50.0                               # on the closed lower edge (kept)
    scan[5::13, 1] = 50.0                                # on the open upper edge (dropped)
    scan = np.ascontiguousarray(scan.astype(np.float32))
    from scipy.spatial.transform import Rotation
    Tl = np.eye(4)
    Tl[:3, :3] = Rotation.from_euler("xyz", [0.31, -0.22, 0.47]).as_matrix()
    Tl[:3, 3] = [1.3, -0.8, 0.25]
    for name, num_points in (("subsample", 2048), ("pad", 6000)):
        import random
        np.random.seed(1234)
        state = np.random.get_state()
        out = lu.preproc_pcd(scan.copy(), {"rand_init_l": Tl}, num_points)
        # the index set the reference drew: replay the RNG on the cropped size
        keep = (scan[:, 0] >= -50.) & (scan[:, 0] < 50.) & (scan[:, 1] >= -50.) & (scan[:, 1] < 50.)
        m = int(keep.sum())
        sample = np.zeros((0,), np.int64)
        if num_points < m:
            np.random.set_state(state)
            sample = np.random.choice(range(m), size=num_points, replace=False, p=None).astype(np.int64)
        np.savez_compressed(os.path.join(OUT, "preproc_%s.npz" % name), scan=scan, T=Tl, num_points=np.int64(num_points),
                            sample=sample, m=np.int64(m), out=np.asarray(out, dtype=np.float64))
        print("preproc", name, out.shape, "cropped", m)


if __name__ == "__main__":
    if "--only-stem" in sys.argv:
        stem_golden()
    elif "--only-projection" in sys.argv:
        projection_golden()
    else:
        main()
        stem_golden()
        projection_golden()
