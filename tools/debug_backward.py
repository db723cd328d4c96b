"""Per-stage backward errors of one E-Net BCL level against the float64 oracle (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from efgh_b200 import synth, bilateralNN
from efgh_b200.generate_data import GenerateData
from efgh_b200.bilateralNN import BilateralConvFlex, conv_dgrad_tc, conv_dgrad, _wt_first, _wt_point
from oracle import bcl as obcl

dev = torch.device("cuda:0")
level = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sensor = sys.argv[2] if len(sys.argv) > 2 else "os1-64-16k"
pc = synth.synth_scan(11, sensor)
gd = GenerateData(3, synth.SCALE_MAP, "cuda", exact=True)
_, data = gd(torch.from_numpy(pc).to(dev))
d = data[level]
cin, nout = synth.ENET_BCL[level]
torch.manual_seed(20 + level)
m = BilateralConvFlex(3, 1, cin, list(nout), "cuda", True, True, True, True, False, False, chunk_size=-1).to(dev)
for p in m.parameters():
    torch.nn.init.normal_(p, 0, 0.1)
H = d["pc1_hash_cnt"]
g = torch.Generator().manual_seed(100 + level)
n_in = d["pc1_barycentric"].shape[-1]
feat0 = torch.randn(1, cin, n_in, generator=g)
gout = torch.randn(1, nout[-1], H, generator=g)

def rel(a, b):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max())

# oracle with intermediates (float64)
W0 = m.blur_conv[0].weight.detach().cpu().double().requires_grad_(True); b0 = m.blur_conv[0].bias.detach().cpu().double().requires_grad_(True)
W1 = m.blur_conv[2].weight.detach().cpu().double().requires_grad_(True); b1 = m.blur_conv[2].bias.detach().cpu().double().requires_grad_(True)
nb = d["pc1_blur_neighbors"][0].cpu(); off = d["pc1_lattice_offset"][0].cpu(); w = d["pc1_barycentric"][0].cpu().double()
f = feat0[0].double().requires_grad_(True)
rows = (off + 1).reshape(-1)
contrib = (w[:, None, :] * f[None]).permute(0, 2, 1).reshape(-1, cin)
S = torch.zeros(H + 1, cin, dtype=torch.float64).index_add(0, rows, contrib)
Wn = torch.zeros(H + 1, dtype=torch.float64).index_add(0, rows, w.reshape(-1))
S = S * (1.0 / (Wn + 1e-5))[:, None]
S.retain_grad()
X = S[nb + 1]
y1 = torch.einsum("fhc,mcf->hm", X, W0[..., 0]) + b0
y1.retain_grad()
a1 = torch.relu(y1); a1.retain_grad()
y2 = a1 @ W1[:, :, 0, 0].t() + b1
y2.backward(gout[0].t().double())
W0g, b0g, W1g, b1g = W0.grad, b0.grad, W1.grad, b1.grad
print("level %d %s: H=%d n_in=%d cin=%d nout=%s" % (level, sensor, H, n_in, cin, nout))

# GPU pieces
dY2 = gout[0].t().contiguous().to(dev)
Y1 = a1.detach().float().to(dev)
Sg = S.detach().float().to(dev)
mirror = m._mirror
for ns in (3,):
    dA1 = conv_dgrad_tc(dY2, None, 0, None, m.blur_conv[2].weight.detach(), mirror, H, ns)
    print("conv2 dgrad tc   (N=%d K=%d): rel err %.3e" % (nout[0], nout[1], rel(dA1, a1.grad)))
    dA1s = conv_dgrad(dY2, None, 0, None, _wt_point(m.blur_conv[2].weight), nout[0], H)
    print("conv2 dgrad ffma           : rel err %.3e" % rel(dA1s, a1.grad))
    dA1x = a1.grad.float().to(dev)
    for bits in (torch.int64, torch.int32):
        nbr = d["pc1_blur_neighbors"].to(bits)
        dS = conv_dgrad_tc(dA1x, Y1, 1, nbr, m.blur_conv[0].weight.detach(), mirror, H + 1, ns)
        if dS is None:
            print("conv1 dgrad tc: not eligible")
        else:
            print("conv1 dgrad tc %s: all %.3e  first4 %.3e  rest %.3e" % (bits, rel(dS[1:], S.grad[1:]), rel(dS[1:, :4], S.grad[1:, :4]), rel(dS[1:, 4:], S.grad[1:, 4:])))
            e = (dS.cpu().double() - S.grad).abs()
            e[0] = 0
            r, c = np.unravel_index(int(e.argmax()), e.shape)
            print("   worst at row %d col %d: got %.6f want %.6f; rows with err>1e-4*max: %d" % (r, c, float(dS[r, c]), float(S.grad[r, c]), int((e.max(1).values > 1e-4 * float(S.grad.abs().max())).sum())))
    dSs = conv_dgrad(dA1x, Y1, 1, d["pc1_blur_neighbors"], _wt_first(m.blur_conv[0].weight), cin, H + 1)
    print("conv1 dgrad ffma scatter   : rel err %.3e" % rel(dSs[1:], S.grad[1:]))
    # splat adjoint (gather with the normalisation factors) fed with the oracle's dS
    inv = (1.0 / (Wn + 1e-5)).float().to(dev)
    # the oracle's S.grad is the gradient of the NORMALISED matrix; the adjoint multiplies by inv[row]
    dfeat = bilateralNN.gather(S.grad.float().to(dev).contiguous(), inv, d["pc1_barycentric"], d["pc1_lattice_offset"], 1, None, n_in)
    print("splat adjoint (gather)     : rel err %.3e" % rel(dfeat, f.grad))
    # the whole module
    for tc in (True, False):
        bilateralNN.DGRAD_ON_TENSOR_CORES = tc
        m.zero_grad(set_to_none=True)
        x = feat0.to(dev).requires_grad_(True)
        out = m(x, d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
        out.backward(gout.to(dev))
        print("module dgrad_tc=%s: fwd %.3e dfeat %.3e dW0 %.3e db0 %.3e dW1 %.3e db1 %.3e" % (
            tc, rel(out[0].t(), y2), rel(x.grad[0], f.grad), rel(m.blur_conv[0].weight.grad, torch.autograd.grad(y2, [], allow_unused=True) if False else W0g),
            rel(m.blur_conv[0].bias.grad, b0g), rel(m.blur_conv[2].weight.grad, W1g), rel(m.blur_conv[2].bias.grad, b1g)))
