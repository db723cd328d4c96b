"""E-Net consumer (reference nets/enet.py): CPU checks of the module surface, GPU forward/backward/optimiser step."""
import math

import numpy as np
import pytest
import torch

from efgh_b200 import synth

ARGS = {"dim": 3, "scale_map": synth.SCALE_MAP, "DEVICE": "cuda", "use_leaky": True, "bcn_use_bias": True,
        "bcn_use_norm": True, "last_relu": False}


def test_enet_state_dict_matches_reference_layout():
    """Key names / shapes a reference checkpoint carries (reference nets/enet.py:24-97, bilateralNN.py:99-143)."""
    from efgh_b200.enet import Enet
    sd = Enet(ARGS).state_dict()
    keys = list(sd)
    for i in (0, 1, 2):
        assert "conv_in.%d.0.weight" % i in keys and "conv_in.%d.0.bias" % i in keys
    plan = {"bcn1": (36, 32), "bcn2": (36, 64), "bcn3": (68, 128), "bcn4": (132, 256), "bcn5": (260, 256)}
    for name, (cin, cout) in plan.items():
        assert tuple(sd[name + ".feat_indices"].shape) == (cin,)
        assert tuple(sd[name + ".blur_conv.0.weight"].shape) == (cout, cin, 15, 1)
        assert tuple(sd[name + ".blur_conv.2.weight"].shape) == (cout, cout, 1, 1)
        assert name + ".out_indices" not in keys and name + ".bias" not in keys      # do_slice=False in E-Net
    for k in ("conv_gn_1.weight", "bn_gn_3.running_var", "lin_gn_abs.weight", "lin_gn_sgn.bias"):
        assert k in keys
    n_bcl = sum(v.numel() for k, v in sd.items() if k.startswith("bcn") and "blur_conv" in k)
    assert n_bcl == 1841728


def test_geometry_helpers_match_reference_formulas():
    from efgh_b200.enet import normal_vector_3d_from_abs_sign, rotation_matrix_between_two_vectors
    torch.manual_seed(0)
    B = 16
    a = torch.rand(B, 3, 1)
    a = a / a.norm(dim=1, keepdim=True)
    sign = torch.randn(B, 8)
    n = normal_vector_3d_from_abs_sign(a, sign)
    for b in range(B):                          # reference common/torch_utils.py:135-144, restated as the loop it is
        code = int(torch.argmax(torch.softmax(sign[b], 0)))
        code, z = divmod(code, 2)
        code, y = divmod(code, 2)
        code, x = divmod(code, 2)
        sg = torch.tensor([x, y, z], dtype=torch.float32)
        sg = torch.where(sg == 0, -torch.ones_like(sg), sg)
        assert torch.allclose(n[b, :, 0], a[b, :, 0] * sg)
    e3 = torch.tensor([0., 0., 1.])[None, :, None]
    R = rotation_matrix_between_two_vectors(n, e3)
    assert tuple(R.shape) == (B, 4, 4)
    mapped = torch.bmm(R[:, :3, :3], n)[:, :, 0]
    assert torch.allclose(mapped, e3[:, :, 0].expand(B, 3), atol=1e-5)
    ident = rotation_matrix_between_two_vectors(e3, e3)
    assert torch.allclose(ident[0], torch.eye(4))
    flip = rotation_matrix_between_two_vectors(-e3, e3)
    assert flip[0, 0, 0] == 1 and flip[0, 2, 2] == -1        # reference special case (1 + c == 0, x components zero)


@pytest.mark.gpu
def test_enet_forward_backward_step():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from efgh_b200.enet import Enet
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    model = Enet(ARGS).to(dev)
    for m in model.modules():                   # weights large enough for a non-degenerate signal
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.normal_(m.weight, 0, 0.05)
    pc = torch.from_numpy(synth.synth_scan(1, "os1-64-16k"))[None].to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    before = model.bcn3.blur_conv[0].weight.detach().clone()
    out = model(pc)
    assert tuple(out["e_gn"].shape) == (1, 3, 1) and tuple(out["e_l"].shape) == (1, 4, 4)
    assert [o.shape[1] for o in out["bcn_outputs"]] == [32, 64, 128, 256, 256]
    assert abs(float(out["e_gn"].norm()) - 1.0) < 1e-4
    target = torch.tensor([[0.0], [0.0], [1.0]], device=dev)[None]
    loss = (1 - (out["e_gn_abs"] * target.abs()).sum()) + out["e_gn_sgn"].logsumexp(1).mean() - out["e_gn_sgn"][:, 7].mean()
    loss.backward()
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    assert len(grads) > 30 and all(torch.isfinite(g).all() for g in grads)
    assert float(model.bcn1.blur_conv[0].weight.grad.abs().max()) > 0     # gradient reached the first BCL through 5 levels
    opt.step()
    assert not torch.equal(before, model.bcn3.blur_conv[0].weight.detach())


@pytest.mark.gpu
def test_enet_batched_inference_matches_module_path():
    """Enet.infer (one batched ScanPipeline launch sequence: stem fused into the level-0 splat, gather-form splat, no
    host sync per level) returns what forward() returns cloud by cloud: BCL features within the chained tolerance,
    same gravity-normal estimate.  Weights are re-drawn (the reference's N(0, 1e-3) init gives ~0 features)."""
    from efgh_b200.enet import Enet
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    net = Enet(ARGS).to(dev).eval()
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.startswith(("conv_in", "bcn")):
                p.normal_(0, 0.1 if p.dim() > 1 else 0.05)
    clouds = [torch.from_numpy(synth.synth_scan(80 + b, "os1-64-16k")).to(dev) for b in range(3)]
    # the module path's stem is a cuDNN Conv1d, which torch lets run in TF32 by default (1e-3 off); the fused stem is fp32
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    request_restore = tf32
    with torch.no_grad():
        want = [net(c[None]) for c in clouds]
        got = net.infer(clouds, vertex_cap_factor=0.5)          # too small on purpose: exercises the capacity retry
    for b in range(3):
        for li, (g, w) in enumerate(zip(got[b]["bcn_outputs"], want[b]["bcn_outputs"])):
            assert tuple(g.shape) == tuple(w.shape)
            err = float((g - w).abs().max() / w.abs().max())
            assert err < 5e-5, "cloud %d level %d rel err %g" % (b, li, err)
        assert float((got[b]["e_gn"] - want[b]["e_gn"]).abs().max()) < 1e-4
    # weights changed in place -> the cached pipeline must be rebuilt
    with torch.no_grad():
        net.bcn1.blur_conv[0].weight.mul_(2.0)
        want2 = net(clouds[0][None])
        got2 = net.infer(clouds[:1])
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = request_restore
    err = float((got2[0]["bcn_outputs"][0] - want2["bcn_outputs"][0]).abs().max() / want2["bcn_outputs"][0].abs().max())
    assert err < 5e-5, "after an in-place weight update: rel err %g" % err
