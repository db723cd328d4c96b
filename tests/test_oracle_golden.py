"""CPU: the oracle (oracle/) against the golden vectors generated from the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import bcl as obcl
from oracle import lattice as ol
from efgh_b200 import synth
from tests import helpers as H


def test_constants():
    E = ol.elevate_matrix()
    # bit patterns of the float32 elevation matrix and std
    assert [hex(v) for v in E.view(np.uint32)[0]] == ["0x3f3504f3", "0x3ed105eb", "0x3e93cd3a"]
    assert hex(int(E.view(np.uint32)[3, 2])) == "0xbf5db3d7"
    assert hex(int(np.float32(ol.expected_std()).view(np.uint32))) == "0x405105ec"
    offs = ol.blur_offsets(1)
    assert offs.shape == (15, 4) and offs[1].tolist() == [-1, -1, -1, 3] and offs[14].tolist() == [1, 1, 1, -3]
    assert np.array_equal(offs[1:], -offs[1:][::-1])  # mirror symmetry f <-> 15-f
    assert ol.blur_offsets(2).shape == (65, 4)


@pytest.mark.parametrize("variant", ["port", "ref"])
@pytest.mark.parametrize("name", H.lattice_golden_names())
def test_lattice_small_golden(name, variant):
    if variant == "ref" and not ol.has_ref():
        pytest.skip("oracle/_ref not built (reference absent)")
    pc, smap, keys0, levels = H.load_lattice_golden(name)
    keys, _, _ = ol.keys_and_barycentric(pc, variant)  # golden keys0: unscaled cloud
    assert np.array_equal(keys, keys0)
    got = ol.generate(pc, smap, variant)
    assert len(got) == len(levels)
    for li, (g, w) in enumerate(zip(got, levels)):
        H.assert_level_equal(g, w, "%s L%d" % (name, li))


@pytest.mark.parametrize("case", H.digest_cases())
def test_lattice_fullsize_digests(case):
    dig = H.load_digests()
    sensor, seed = case.split("/")
    pc = synth.synth_scan(int(seed[4:]), sensor)
    assert H.digest(pc) == dig[case + "/pc"], "synthetic cloud generator drifted"
    got = ol.generate(pc, synth.SCALE_MAP)
    assert ",".join(str(g["pc1_hash_cnt"]) for g in got) == dig[case + "/cnt"]
    for li, g in enumerate(got):
        for k in ("pc1_barycentric", "pc1_el_minus_gr", "pc1_lattice_offset", "pc1_blur_neighbors"):
            assert H.digest(g[k]) == dig["%s/L%d/%s" % (case, li, k)], (case, li, k)


def _bcl_from_golden(z, dtype):
    cfg = z["cfg"].tolist()
    num_input, do_splat, do_slice, use_norm, last_relu, use_leaky = cfg[:6]
    nconv = len(cfg) - 6
    t = lambda a: torch.from_numpy(np.asarray(a))
    convs = []
    keys = sorted(k for k in z.files if k.startswith("p_blur_conv") and k.endswith("weight"))
    for wk in keys:
        convs.append((t(z[wk]).clone().requires_grad_(True), t(z[wk.replace("weight", "bias")]).clone().requires_grad_(True)))
    assert len(convs) == nconv
    feat = t(z["feat"]).clone().requires_grad_(True)
    bias = t(z["p_bias"]).clone().requires_grad_(True) if "p_bias" in z.files else None
    bary, off, nbr = t(z["bary"]), t(z["off"].astype(np.int64)), t(z["nbr"].astype(np.int64))
    out = obcl.bcl_forward(feat, bary, off, nbr, convs, use_norm=bool(use_norm), do_splat=bool(do_splat),
                           do_slice=bool(do_slice), out_bary=bary if do_slice else None,
                           out_off=off if do_slice else None, slice_bias=bias, last_relu=bool(last_relu),
                           use_leaky=bool(use_leaky), dtype=dtype)
    return out, feat, convs, bias, keys


@pytest.mark.parametrize("name", H.bcl_golden_names())
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_bcl_oracle_vs_reference_golden(name, dtype, golden_dir):
    z = np.load("%s/bcl_%s.npz" % (golden_dir, name))
    out, feat, convs, bias, keys = _bcl_from_golden(z, dtype)
    assert tuple(out.shape) == z["out"].shape
    assert H.rel_err(out.detach().numpy(), z["out"]) < 2e-5
    out.backward(torch.from_numpy(z["gout"]).to(dtype))
    assert H.rel_err(feat.grad.numpy(), z["gfeat"]) < 2e-5
    for wk, (W, b) in zip(keys, convs):
        assert H.rel_err(W.grad.numpy(), z["g_" + wk[2:]]) < 5e-5
        assert H.rel_err(b.grad.numpy(), z["g_" + wk[2:].replace("weight", "bias")]) < 5e-5
    if bias is not None:
        assert H.rel_err(bias.grad.numpy(), z["g_bias"]) < 5e-5


@pytest.mark.reference
def test_oracle_vs_live_reference_random():
    """Only in the build container: fresh random clouds through the unmodified reference."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("live reference not present")
    _, g, _ = ref_harness.load()
    rng = np.random.default_rng(99)
    for n, spread in ((1, 5.0), (3, 0.1), (777, 20.0), (4096, 40.0)):
        pc = (rng.standard_normal((3, n)) * spread).astype(np.float32)
        smap = [[1.0, 1], [0.5, 1], [0.25, 1]]
        _, ref = g.GenerateData(3, smap, "cpu")(torch.from_numpy(pc.copy()))
        got = ol.generate(pc, smap)
        for li, (gg, w) in enumerate(zip(got, ref)):
            w = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in w.items()}
            H.assert_level_equal(gg, w, "n=%d L%d" % (n, li))


@pytest.mark.parametrize("name", ["leaky", "relu"])
def test_stem_oracle_vs_reference_golden(name, golden_dir):
    """oracle.bcl.stem_forward (E-Net conv_in, reference nets/enet.py:24-28,111) against the output of the live
    reference's nets/net_utils.py conv_1x1 stack stored by oracle/make_golden.py."""
    import numpy as np
    import torch
    from oracle import bcl as obcl
    g = np.load(os.path.join(golden_dir, "stem_%s.npz" % name))
    layers = [(g["W%d" % i], g["b%d" % i]) for i in range(3)]
    got = obcl.stem_forward(g["pc"], layers, leaky=bool(g["leaky"]), dtype=torch.float32).numpy()
    assert got.shape == g["out"].shape
    assert np.abs(got - g["out"]).max() <= 2e-6 * np.abs(g["out"]).max()


# ------------------------------------------------------------------------------------------------
# f3 / f4: projections and pre-processing oracle against fixtures of the live reference (oracle/make_golden.py)
# ------------------------------------------------------------------------------------------------
def test_depth_image_oracle_matches_reference_golden(golden_dir):
    """reference common/torch_utils.py:61-103 run live on CPU: bit-exact (float32 sgemm chain, true divisions,
    truncating index conversion, last point wins on duplicate pixels)."""
    import numpy as np
    from oracle import projection as op
    z = np.load(os.path.join(golden_dir, "proj_depth.npz"))
    got = op.depth_image(z["pc"], z["T"], tuple(int(v) for v in z["size"]))
    assert got.shape == z["img"].shape
    assert np.array_equal(got.view(np.int32), z["img"].view(np.int32))
    assert int((z["img"][:, 3] != 0).sum()) > 1000


def test_range_image_oracle_matches_reference_golden(golden_dir):
    """reference common/torch_utils.py:11-59 run live on CPU.  The oracle rounds asin / atan2 correctly (float64, one
    rounding); torch's CPU kernels (SLEEF) differ from that in the last bit for a few per cent of the inputs, which
    moves a pixel (or the FoV mask) only when the point sits within an ulp of a pixel border or of a FoV limit - the
    synthetic OS1-64 puts its first and last beam EXACTLY on the configured +-22.5 degree limits, so those two rings
    are such points.  Every pixel of the two images is either identical or explained by a border point."""
    import numpy as np
    from oracle import projection as op
    z = np.load(os.path.join(golden_dir, "proj_range.npz"))
    size, fov = tuple(int(v) for v in z["size"]), tuple(float(v) for v in z["fov"])
    got = op.range_image(z["pc"], size, fov)
    ref = z["img"]
    assert got.shape == ref.shape
    differ = (got[:, :3].view(np.int32) != ref[:, :3].view(np.int32)).any(axis=1)   # (B, H, W) pixels owned by another point
    # the range channel: torch's CPU sqrt (SLEEF u05) is not the IEEE square root - 0.7 % of the values differ by one ulp
    # from np.sqrt / CUDA's sqrt.rn, which the oracle and the kernel use
    same = ~differ
    r_got, r_ref = got[:, 3][same], ref[:, 3][same]
    assert np.all(np.abs(r_got.view(np.int32).astype(np.int64) - r_ref.view(np.int32).astype(np.int64)) <= 1)
    assert np.mean(r_got != r_ref) < 0.02
    # border points: exact (float64) pitch within 2e-6 rad of a FoV limit, or exact u / v within a few float32 ulps of an integer
    import math
    Hh, Ww = size
    fu, fd = fov[0] * math.pi, fov[1] * math.pi
    allowed = np.zeros_like(differ)
    n_border = 0
    for b in range(z["pc"].shape[0]):
        x, y, zz = (z["pc"][b, i].astype(np.float64) for i in range(3))
        with np.errstate(all="ignore"):
            pitch = np.arcsin(zz / np.sqrt(x * x + y * y + zz * zz))
            u = (fu - pitch) / (fu - fd) * (Hh - 1)
            v = (-np.arctan2(y, x) + math.pi) / (2 * math.pi) * (Ww - 1)
        border = (np.abs(pitch - fu) < 2e-6) | (np.abs(pitch - fd) < 2e-6) | (np.abs(u - np.rint(u)) < 1e-4) | (np.abs(v - np.rint(v)) < 1e-3)
        border &= np.isfinite(pitch)
        n_border += int(border.sum())
        for du in (-1, 0, 1):
            for dv in (-1, 0, 1):
                uu = np.clip(np.floor(u[border]).astype(int) + du, 0, Hh - 1)
                vv = np.clip(np.floor(v[border]).astype(int) + dv, 0, Ww - 1)
                allowed[b, uu, vv] = True
    n_points = z["pc"].shape[0] * z["pc"].shape[2]
    print("range image: %d of %d pixels differ from the SLEEF-based live reference; %d of %d points sit on a pixel / FoV border"
          % (int(differ.sum()), differ.size, n_border, n_points))
    assert not (differ & ~allowed).any(), "a pixel differs that no border point can explain"
    assert n_border < 0.05 * n_points and differ.sum() <= n_border
    assert int((ref[:, 3] > 0).sum()) > 5000


@pytest.mark.parametrize("name", ["subsample", "pad"])
def test_preproc_oracle_matches_reference_golden(name, golden_dir):
    """reference data_loader/loader_utils.py:163-202 run live with a seeded numpy RNG: bit-exact float64."""
    import numpy as np
    from oracle import projection as op
    z = np.load(os.path.join(golden_dir, "preproc_%s.npz" % name))
    got, m = op.preproc_pcd(z["scan"], z["T"], int(z["num_points"]), sample=z["sample"] if z["sample"].size else None)
    assert m == int(z["m"])
    assert got.shape == z["out"].shape and np.array_equal(got, z["out"])
