"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): two 2048-point scans through one
batched launch sequence (all lattice kernels, atomic + gather-form splat, tensor-core conv), plus the drop-in
modules forward + backward on one scan.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from efgh_b200 import synth
from efgh_b200.pipeline import ScanPipeline, make_enet_weights
from efgh_b200.generate_data import GenerateData
from efgh_b200.bilateralNN import BilateralConvFlex

dev = torch.device("cuda:0")
n = 2048
clouds = [synth.synth_scan(s, "os1-64-16k")[:, :n] for s in (0, 1)]
weights = make_enet_weights(synth.ENET_BCL, seed=1)
pipe = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=16.0, batch=2)
feats = torch.randn(32, 2 * n)
pipe.enqueue(torch.from_numpy(np.concatenate(clouds, 1)).to(dev), feats.to(dev))
print("batched counts", pipe.counts())
gd = GenerateData(3, synth.SCALE_MAP[:2], "cuda", exact=False)
_, data = gd(torch.from_numpy(clouds[0]).to(dev))
x = torch.randn(1, 32, n, device=dev, requires_grad=True)
y = x
for (cin, nout), d in zip(synth.ENET_BCL[:2], data):
    m = BilateralConvFlex(3, 1, cin, nout, "cuda", True, True, True, True, False, False, chunk_size=-1).to(dev)
    y = m(torch.cat((d["pc1_el_minus_gr"], y), 1), d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
y.square().mean().backward()
torch.cuda.synchronize()
print("modules fwd+bwd ok", float(x.grad.abs().sum()))

# capacity overflow: more vertices than the arrays hold -> status bit, no out-of-bounds access
from efgh_b200.generate_data import VertexCapExceeded
small = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=0.5, batch=2)
small.enqueue(torch.from_numpy(np.concatenate(clouds, 1)).to(dev), feats.to(dev))
try:
    small.counts()
    print("unexpected: no overflow")
except VertexCapExceeded as e:
    print("overflow reported:", e)
torch.cuda.synchronize()

# ---- round-2 kernels: warp-private splat with the fused stem, batched backward (tensor-core weight gradient, k_act_bwd,
#      loss), image projections and pre-processing
gs_ = torch.Generator().manual_seed(5)
stem_layers = [(torch.randn(co, ci, 1, generator=gs_) * 0.3, torch.randn(co, generator=gs_) * 0.1) for ci, co in ((3, 32), (32, 32), (32, 32))]
pstem = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=16.0, batch=2, stem=(stem_layers, True))
pstem.enqueue(torch.from_numpy(np.concatenate(clouds, 1)).to(dev), None)
print("stem-fused counts", pstem.counts())
ptrain = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=16.0, emit_int64=False, batch=2, train=True)
ptrain.enqueue(torch.from_numpy(np.concatenate(clouds, 1)).to(dev), feats.to(dev))
loss, dZ = ptrain.loss_half_mean_square()
dfeat = ptrain.backward(dZ)
torch.cuda.synchronize()
print("batched backward ok", float(loss), float(dfeat.abs().sum()), float(ptrain.weight_grads()[0][0][0].abs().sum()))
from efgh_b200 import projections, preproc
pc3 = torch.from_numpy(clouds[0])[None].to(dev)
ri = projections.range_img_from_cartesian_pc_torch(pc3, (64, 256), (0.4, -0.4), "cuda")
T = torch.eye(4)[:3][None].to(dev) * 100.0
di = projections.depth_img_from_cartesian_pc_torch(pc3, T, (64, 128), "cuda")
pcd = np.concatenate((clouds[0].T, np.ones((n, 1), np.float32)), 1)
out = preproc.preproc_pcd(pcd, {"rand_init_l": np.eye(4)}, 1024, radius=50.0)
torch.cuda.synchronize()
print("projections / preproc ok", tuple(ri.shape), tuple(di.shape), tuple(out.shape))
