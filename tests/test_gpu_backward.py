"""GPU (-m gpu): BCL backward at every E-Net shape against the float64 oracle's autograd, and the
NCCL gradient all-reduce of the training configuration (BASELINE configs[3])."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from efgh_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

GRAD_TOL = 2e-5          # max abs error / max abs value of the float64 oracle gradient
CHAIN_GRAD_TOL = 1e-4    # five chained layers: per-layer forward errors compound into the gradients


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from efgh_b200 import _capi
    _capi.lib()
    return torch.device("cuda:0")


_lattices = {}


def _lattice(sensor, dev):
    """Five-level lattice of one synthetic cloud (built once per module): the real neighbour tables / offsets every
    level's BCL sees in E-Net."""
    if sensor not in _lattices:
        from efgh_b200.generate_data import GenerateData
        pc = synth.synth_scan(11, sensor)
        gd = GenerateData(3, synth.SCALE_MAP, "cuda", exact=True)
        _, data = gd(torch.from_numpy(pc).to(dev))
        _lattices[sensor] = data
    return _lattices[sensor]


def _module(level, dev, seed):
    from efgh_b200.bilateralNN import BilateralConvFlex
    cin, nout = synth.ENET_BCL[level]
    torch.manual_seed(seed)
    m = BilateralConvFlex(3, 1, cin, list(nout), "cuda", True, True, True, True, False, False, chunk_size=-1).to(dev)
    for p in m.parameters():
        torch.nn.init.normal_(p, 0, 0.1)
    return m


def _oracle_convs(m):
    return [(c.weight.detach().cpu().double().requires_grad_(True), c.bias.detach().cpu().double().requires_grad_(True))
            for c in m.blur_conv if isinstance(c, torch.nn.Conv2d)]


def _oracle_pre(m, d, feat):
    """Forward only: output and the pre-activation of the ReLU between the two convolutions."""
    from oracle import bcl as obcl
    pre = []
    with torch.no_grad():
        out = obcl.bcl_forward(feat.detach().cpu().double(), d["pc1_barycentric"].cpu(), d["pc1_lattice_offset"].cpu(),
                               d["pc1_blur_neighbors"].cpu(), _oracle_convs(m), dtype=torch.float64, pre_acts=pre)
    return out, pre[0]


def _check_active_set(mask_gpu, pre, tol, ctx):
    """A ReLU's derivative jumps at 0: the kernels' active set may differ from the float64 oracle's only where the
    oracle's pre-activation lies within the FORWARD tolerance of zero.  Returns the number of such elements."""
    differ = mask_gpu != (pre > 0)
    n = int(differ.sum())
    if n:
        worst = float(pre[differ].abs().max() / pre.abs().max())
        assert worst < tol, "%s: ReLU active sets differ at a pre-activation %.2e of the maximum away from zero" % (ctx, worst)
    return n


def _oracle_grads(m, d, feat, gout, mask):
    from oracle import bcl as obcl
    convs = _oracle_convs(m)
    f = feat.detach().cpu().double().requires_grad_(True)
    out = obcl.bcl_forward(f, d["pc1_barycentric"].cpu(), d["pc1_lattice_offset"].cpu(), d["pc1_blur_neighbors"].cpu(), convs,
                           dtype=torch.float64, relu_masks=[mask])
    out.backward(gout.double())
    return out.detach(), f.grad, convs


@pytest.mark.parametrize("sensor", ["os1-64-16k", "os1-64-64k"])
@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_bcl_backward_every_enet_level_vs_oracle(level, sensor, dev, monkeypatch):
    """Every E-Net BCL shape (36->[32,32] ... 260->[256,256]) on its REAL lattice level of a 16k and a 65k cloud:
    forward within 1e-5 and d features / d weights / d biases within 2e-5 of the float64 oracle's autograd - for the
    tensor-core (gather-form) data gradient + tensor-core weight gradient and the scatter-form data gradient + fp32 CUDA-core
    weight gradient, with int64 and int32 lattice indices.  Both sides
    differentiate the inner ReLU with the kernels' active set, after checking that it differs from the oracle's own
    only at pre-activations within the forward tolerance of zero (_check_active_set)."""
    from efgh_b200 import bilateralNN
    d = _lattice(sensor, dev)[level]
    cin, nout = synth.ENET_BCL[level]
    n_in = d["pc1_barycentric"].shape[-1]
    m = _module(level, dev, 20 + level)
    g = torch.Generator().manual_seed(100 + level)
    feat0 = torch.randn(1, cin, n_in, generator=g)
    gout = torch.randn(1, nout[-1], d["pc1_hash_cnt"], generator=g)
    ref_out, pre = _oracle_pre(m, d, feat0)
    mods = [c for c in m.blur_conv if isinstance(c, torch.nn.Conv2d)]
    keep = {}
    monkeypatch.setattr(bilateralNN, "DEBUG_KEEP", keep)
    for dgrad_tc in (True, False):
        for idx_dtype in ((torch.int64, torch.int32) if dgrad_tc else (torch.int64,)):
            monkeypatch.setattr(bilateralNN, "DGRAD_ON_TENSOR_CORES", dgrad_tc)
            monkeypatch.setattr(bilateralNN, "WGRAD_ON_TENSOR_CORES", dgrad_tc)    # (the scatter pass also takes the fp32 CUDA-core weight gradient)
            m.zero_grad(set_to_none=True)
            feat = feat0.to(dev).requires_grad_(True)
            off = d["pc1_lattice_offset"].to(idx_dtype)
            nbr = d["pc1_blur_neighbors"].to(idx_dtype)
            out = m(feat, d["pc1_barycentric"], off, nbr, None, None)
            ctx = "level %d %s dgrad=%s idx=%s" % (level, sensor, "tc" if dgrad_tc else "scatter", idx_dtype)
            assert H.rel_err(out.detach().cpu().numpy(), ref_out.numpy()) < 1e-5, ctx
            mask = (keep["ys"][0] > 0).cpu()                      # the active set of the ReLU the kernels applied
            flips = _check_active_set(mask, pre, 1e-5, ctx)
            _, ref_gfeat, ref_convs = _oracle_grads(m, d, feat0, gout, mask)
            out.backward(gout.to(dev))
            print("%s: %d ReLU elements within the forward tolerance of 0 differ from the oracle's active set" % (ctx, flips))
            assert H.rel_err(feat.grad.cpu().numpy(), ref_gfeat.numpy()) < GRAD_TOL, ctx + " d feat"
            for c, (W, b) in zip(mods, ref_convs):
                assert H.rel_err(c.weight.grad.cpu().numpy(), W.grad.numpy()) < GRAD_TOL, ctx + " d weight"
                assert H.rel_err(c.bias.grad.cpu().numpy(), b.grad.numpy()) < GRAD_TOL, ctx + " d bias"


def test_bcl_chain_backward_vs_oracle_chain(dev, monkeypatch):
    """The five E-Net BCLs chained as reference nets/enet.py:113-141 (level l's input = cat(el_minus_gr_l, level l-1's
    output)), loss on bcn5's output: gradients of the stem features and of all 20 weight / bias tensors against the
    float64 oracle chain's autograd (same ReLU active sets, checked against the chained forward tolerance)."""
    from efgh_b200 import bilateralNN
    from oracle import bcl as obcl
    data = _lattice("os1-64-16k", dev)
    mods = [_module(li, dev, 40 + li) for li in range(5)]
    g = torch.Generator().manual_seed(7)
    feat0 = torch.randn(1, 32, data[0]["pc1_barycentric"].shape[-1], generator=g)
    gout = torch.randn(1, 256, data[4]["pc1_hash_cnt"], generator=g)
    keep = {}
    monkeypatch.setattr(bilateralNN, "DEBUG_KEEP", keep)
    x = feat0.to(dev).requires_grad_(True)
    h = x
    masks = []
    for d, m in zip(data, mods):
        h = m(torch.cat((d["pc1_el_minus_gr"], h), 1), d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
        masks.append((keep["ys"][0] > 0).cpu())
    h.backward(gout.to(dev))
    f = feat0.double().requires_grad_(True)
    ref_w = []
    r = f
    flips = []
    for li, (d, m) in enumerate(zip(data, mods)):
        convs = _oracle_convs(m)
        ref_w.append(convs)
        pre = []
        r = obcl.bcl_forward(torch.cat((d["pc1_el_minus_gr"].cpu().double(), r), 1), d["pc1_barycentric"].cpu(),
                             d["pc1_lattice_offset"].cpu(), d["pc1_blur_neighbors"].cpu(), convs, dtype=torch.float64,
                             relu_masks=[masks[li]], pre_acts=pre)
        flips.append(_check_active_set(masks[li], pre[0], 5e-5, "chain level %d" % li))
    print("ReLU elements at the kink per level:", flips)
    r.backward(gout.double())
    assert H.rel_err(h.detach().cpu().numpy(), r.detach().numpy()) < 5e-5
    assert H.rel_err(x.grad.cpu().numpy(), f.grad.numpy()) < CHAIN_GRAD_TOL, "d stem features"
    for li, (m, convs) in enumerate(zip(mods, ref_w)):
        cs = [c for c in m.blur_conv if isinstance(c, torch.nn.Conv2d)]
        for c, (W, b) in zip(cs, convs):
            assert H.rel_err(c.weight.grad.cpu().numpy(), W.grad.numpy()) < CHAIN_GRAD_TOL, "level %d d weight" % li
            assert H.rel_err(c.bias.grad.cpu().numpy(), b.grad.numpy()) < CHAIN_GRAD_TOL, "level %d d bias" % li


def test_weight_cache_sees_data_edits_and_device_moves(dev):
    """re-laid / packed weights must never go stale - not after `.data` edits recorded by autograd-less
    code paths (explicit invalidation), not after an optimizer step, not across `init_weights`."""
    from efgh_b200 import bilateralNN
    d = _lattice("os1-64-16k", dev)[0]
    m = _module(0, dev, 3)
    feat = torch.randn(1, 36, d["pc1_barycentric"].shape[-1], device=dev)
    args = (d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
    with torch.no_grad():
        y0 = m(feat, *args).clone()
        assert H.rel_err(m(feat, *args).cpu().numpy(), y0.cpu().numpy()) < 1e-6     # cached path (the atomic splat sums in arrival order)
        m.blur_conv[0].weight.mul_(2.0)                              # in-place through the tensor: version bump
        y1 = m(feat, *args).clone()
        assert not torch.allclose(y1, y0)
        m.blur_conv[0].weight.data.mul_(0.5)                         # through .data: invisible to the version counter
        m.invalidate_cache()
        assert H.rel_err(m(feat, *args).cpu().numpy(), y0.cpu().numpy()) < 1e-6
    # training: a recorded forward never trusts the cache
    m.blur_conv[0].weight.data.mul_(2.0)
    y2 = m(feat.requires_grad_(True), *args)
    assert H.rel_err(y2.detach().cpu().numpy(), y1.cpu().numpy()) < 1e-6
    for mod in m.blur_conv.modules():
        bilateralNN.init_weights(mod)
    with torch.no_grad():
        assert float(m(feat, *args).abs().max()) < 1e-3              # N(0, 1e-3) weights: the old images are gone


@pytest.mark.parametrize("world", [2])
def test_gradient_allreduce_nccl(world, dev):
    """sharding.allreduce_gradients over NCCL (one process per GPU, torchrun): replicas with different per-rank
    gradients end up with the mean on every rank, in a few bucketed calls."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(root, "tests", "nccl_allreduce_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "NCCL_ALLREDUCE_OK" in r.stdout, r.stdout[-2000:]


def test_pipeline_batched_backward_vs_oracle(dev):
    """ScanPipeline(train=True): B scans through ONE forward launch sequence and ONE backward (what bench.py --train
    times).  Loss, d feat0 and the 20 weight / bias gradients (summed over the scans) against the float64 oracle chain
    run scan by scan with the kernels' ReLU active sets."""
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    from oracle import lattice as ol, bcl as obcl
    B, n = 2, 16384
    clouds = [synth.synth_scan(90 + b, "os1-64-16k") for b in range(B)]
    weights = make_enet_weights(synth.ENET_BCL, seed=12)
    pipe = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=4.0, batch=B, emit_int64=False, train=True)
    g = torch.Generator().manual_seed(4)
    feat0 = torch.randn(32, B * n, generator=g)
    pc_all = torch.from_numpy(np.concatenate(clouds, 1)).to(dev)
    for _ in range(2):                                              # twice: gradient buffers are reused
        pipe.enqueue(pc_all, feat0.to(dev))
        loss, dZ = pipe.loss_half_mean_square()
        dfeat0 = pipe.backward(dZ)
    torch.cuda.synchronize()
    assert pipe.aliased_levels() == []
    vs = pipe.vertex_starts()
    grads = pipe.weight_grads()
    ref_loss = 0.0
    ref_w = [[(W.double().clone().requires_grad_(True), b.double().clone().requires_grad_(True)) for W, b in lv] for lv in weights]
    for b in range(B):
        want = ol.generate(clouds[b], synth.SCALE_MAP)
        f = feat0[None, :, b * n:(b + 1) * n].double().requires_grad_(True)
        r = f
        for li, w in enumerate(want):
            mask = (pipe.levels[li]["Y"][vs[li][b]:vs[li][b + 1]] > 0).cpu()
            pre = []
            r = obcl.bcl_forward(torch.cat((torch.from_numpy(w["pc1_el_minus_gr"]).double(), r), 1), torch.from_numpy(w["pc1_barycentric"]),
                                 torch.from_numpy(w["pc1_lattice_offset"]), torch.from_numpy(w["pc1_blur_neighbors"]), ref_w[li],
                                 dtype=torch.float64, relu_masks=[mask], pre_acts=pre)
            _check_active_set(mask, pre[0], 5e-5, "scan %d level %d" % (b, li))
        lb = 0.5 * r.square().mean() / B
        lb.backward()
        ref_loss += float(lb.detach())
        assert H.rel_err(dfeat0[:, b * n:(b + 1) * n].cpu().numpy(), f.grad[0].numpy()) < CHAIN_GRAD_TOL, "scan %d d feat0" % b
    assert abs(float(loss) - ref_loss) < 2e-4 * abs(ref_loss)          # a chained quantity: five layers' forward errors
    for li in range(5):
        for k in range(2):
            gw, gb = grads[li][k]
            assert H.rel_err(gw.cpu().numpy(), ref_w[li][k][0].grad.numpy()) < CHAIN_GRAD_TOL, "level %d conv %d d weight" % (li, k)
            assert H.rel_err(gb.cpu().numpy(), ref_w[li][k][1].grad.numpy()) < CHAIN_GRAD_TOL, "level %d conv %d d bias" % (li, k)


def test_batched_trainer_matches_module_path_trainer(dev):
    """One Adam step of BatchedTrainer (one launch sequence forward + backward for all scans) moves the BCL weights like
    per-scan autograd through the drop-in modules does, given the same loss."""
    from efgh_b200 import training
    clouds = [torch.from_numpy(synth.synth_scan(95 + b, "os1-64-16k")).to(dev) for b in range(2)]
    bt = training.BatchedTrainer(clouds, dev, vertex_cap_factor=4.0)
    with torch.no_grad():
        for p in bt.params:
            p.normal_(0, 0.1 if p.dim() > 1 else 0.05)
    before = [p.detach().clone() for p in bt.params]
    # reference step: same weights, loss 0.5 * mean(Z_b^2) / B through the module path
    for p in bt.params:
        p.grad = None
    total = 0.0
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False      # the stem is a cuDNN Conv1d
    for c in clouds:
        _, data = bt_gd(dev)(c)
        h = bt.stem(c[None])
        for d, m in zip(data, bt.bcns):
            h = m(torch.cat((d["pc1_el_minus_gr"], h), 1), d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
        lb = 0.5 * h.square().mean() / len(clouds)
        lb.backward()
        total += float(lb)
    want = [p.grad.detach().clone() for p in bt.params]
    for p in bt.params:
        p.grad = None
    try:
        loss = bt.step()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert abs(float(loss) - total) < 1e-3 * abs(total)
    moved = 0
    for p, b, w in zip(bt.params, before, want):
        assert torch.isfinite(p).all()
        moved += int(not torch.equal(p.detach(), b))
        # the gradients the optimizer consumed (still on .grad) match the module path's.  Loose on purpose: the two paths
        # sum the splat in different orders, and a single pre-activation that lands on the other side of the ReLU kink
        # moves a weight-gradient column of the small deep levels by ~1e-2 (the rigorous check is the oracle test above).
        assert float((p.grad - w).abs().max()) <= 3e-2 * float(w.abs().max()) + 1e-12
    assert moved == len(bt.params)


def bt_gd(dev):
    from efgh_b200.generate_data import GenerateData
    return GenerateData(3, synth.SCALE_MAP, "cuda", exact=False)

