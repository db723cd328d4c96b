"""Timing of the DROP-IN module path (GenerateData + 5 x BilateralConvFlex, as reference nets/enet.py uses them),
forward only, one 131k-point scan per call - the path an EFGH user gets by changing two imports."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efgh_b200 import synth
from efgh_b200.generate_data import GenerateData
from efgh_b200.bilateralNN import BilateralConvFlex

dev = torch.device("cuda:0")
torch.manual_seed(0)
pcs = [torch.from_numpy(synth.synth_scan(s, "os1-64")).to(dev) for s in range(4)]
feat = torch.randn(1, 32, pcs[0].shape[1], device=dev)
for exact in (True, False):
    gd = GenerateData(3, synth.SCALE_MAP, "cuda", exact=exact)
    bcls = [BilateralConvFlex(3, 1, cin, nout, "cuda", True, True, True, True, False, False, chunk_size=-1).to(dev) for cin, nout in synth.ENET_BCL]
    def fwd(pc):
        with torch.no_grad():
            _, data = gd(pc)
            x = feat
            for d, m in zip(data, bcls):
                x = m(torch.cat((d["pc1_el_minus_gr"], x), 1), d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
            return x
    for _ in range(3):
        fwd(pcs[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for i in range(n):
        fwd(pcs[i % 4])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("module path, exact=%s: %.3f ms per scan (%.0f scans/s)" % (exact, dt * 1e3, 1.0 / dt))
