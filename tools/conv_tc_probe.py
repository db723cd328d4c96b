"""Standalone probe (not collected by pytest): tensor-core conv vs the fp32 CUDA-core conv on random data."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efgh_b200 import _capi, bilateralNN as B

torch.manual_seed(0)
dev = torch.device("cuda:0")
L = _capi.lib()
for (H, C, F, M) in [(300, 36, 15, 32), (1000, 36, 15, 64), (5000, 68, 15, 128), (700, 132, 15, 256), (4000, 32, 1, 32), (900, 256, 1, 256), (100654, 36, 15, 32)]:
    X = torch.randn(H + 1, C, device=dev); X[0] = 0
    scale = torch.rand(H + 1, device=dev) + 0.5
    nbr = torch.randint(-1, H, (1, F, H), device=dev, dtype=torch.int32) if F > 1 else None
    Wt = torch.randn(F * C, M, device=dev) * 0.1
    bias = torch.randn(M, device=dev)
    Xin = X if F > 1 else X[1:].contiguous()
    sc = scale if F > 1 else None
    B.CONV_PRECISION = "fp32"
    Xn = (Xin * sc[:, None]).contiguous() if sc is not None else Xin
    ref = B.conv(Xn, nbr, Wt, bias, 1, H)
    for prec in ("3xtf32", "tf32"):
        B.CONV_PRECISION = prec
        got = B.conv(Xn, nbr, Wt, bias, 1, H)
        torch.cuda.synchronize()
        err = float((got - ref).abs().max() / ref.abs().max())
        print("H=%d C=%d F=%d M=%d %s rel err %.3e" % (H, C, F, M, prec, err), flush=True)
# raw accumulate + deferred bias/act chain: conv1 (accumulate) -> conv2 (in_bias, in_act)
H, C, F, M1, M2 = 4191, 132, 15, 256, 256
X = torch.randn(H + 1, C, device=dev); X[0] = 0
scale = torch.rand(H + 1, device=dev) + 0.5
nbr = torch.randint(-1, H, (1, F, H), device=dev, dtype=torch.int32)
Wt0 = torch.randn(F * C, M1, device=dev) * 0.05; b0 = torch.randn(M1, device=dev)
Wt1 = torch.randn(M1, M2, device=dev) * 0.1; b1 = torch.randn(M2, device=dev)
B.CONV_PRECISION = "fp32"
X = (X * scale[:, None]).contiguous()
ref = B.conv(B.conv(X, nbr, Wt0, b0, 1, H), None, Wt1, b1, 0, H)
for ns in (3, 1):
    img0 = torch.empty(L.efgh_bcl_packed_weight_bytes(F * C, M1, ns) // 4, device=dev)
    img1 = torch.empty(L.efgh_bcl_packed_weight_bytes(M1, M2, ns) // 4, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _capi.check(L.efgh_bcl_pack_weights(Wt0.data_ptr(), F * C, M1, ns, img0.data_ptr(), st), "pack")
    _capi.check(L.efgh_bcl_pack_weights(Wt1.data_ptr(), M1, M2, ns, img1.data_ptr(), st), "pack")
    Y = torch.zeros(H, M1, device=dev); Z = torch.empty(H, M2, device=dev)
    _capi.check(L.efgh_bcl_conv_tc(X.data_ptr(), C, C, None, 0, nbr.data_ptr(), 32, H, F, H, None, img0.data_ptr(),
                                   None, M1, 0, Y.data_ptr(), M1, ns, 1, st), "conv1")
    _capi.check(L.efgh_bcl_conv_tc(Y.data_ptr(), M1, M1, b0.data_ptr(), 1, None, 32, 0, 1, H, None, img1.data_ptr(),
                                   b1.data_ptr(), M2, 0, Z.data_ptr(), M2, ns, 0, st), "conv2")
    torch.cuda.synchronize()
    print("chain nsplit=%d groups=%d rel err %.3e" % (ns, L.efgh_bcl_conv_tc_groups(F * C, M), float((Z - ref).abs().max() / ref.abs().max())), flush=True)
print("probe done")
