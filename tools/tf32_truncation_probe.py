"""Does tcgen05.mma kind::tf32 truncate or round the low 13 mantissa bits of an fp32 A operand read from TMEM?
Runs the one-pass TF32 convolution twice - A masked to TF32 by the producers (the product path) and A handed over
as raw fp32 (debug flag 4) - and compares the outputs bit for bit."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efgh_b200 import _capi

dev = torch.device("cuda:0")
L = _capi.lib()
L.efgh_debug_set_conv_flags.argtypes = [ctypes.c_int]
L.efgh_debug_set_conv_flags.restype = None
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
H, C, M = 4096, 32, 32
X = torch.randn(H, C, device=dev)
Wt = torch.randn(C, M, device=dev)
img = torch.empty(L.efgh_bcl_packed_weight_bytes(C, M, 1) // 4, device=dev)
_capi.check(L.efgh_bcl_pack_weights(Wt.data_ptr(), C, M, 1, img.data_ptr(), st), "pack")
outs = []
for flag in (0, 4):
    L.efgh_debug_set_conv_flags(flag)
    Y = torch.empty(H, M, device=dev)
    _capi.check(L.efgh_bcl_conv_tc(X.data_ptr(), C, C, None, 0, None, 32, 0, 1, H, None, img.data_ptr(), None, M, 0,
                                   Y.data_ptr(), M, 1, 0, st), "conv")
    torch.cuda.synchronize()
    outs.append(Y.clone())
L.efgh_debug_set_conv_flags(0)
same = torch.equal(outs[0], outs[1])
Xm = (X.view(torch.int32) & -8192).view(torch.float32)
Wm = (Wt.view(torch.int32) & -8192).view(torch.float32)
ref_trunc = (Xm.double() @ Wm.double()).float()
# round-to-nearest-even to 10 mantissa bits
def rn_tf32(t):
    i = t.view(torch.int32).to(torch.int64)
    i = (i + 0x0fff + ((i >> 13) & 1)) & ~0x1fff
    return i.to(torch.int32).view(torch.float32)
ref_round = (rn_tf32(X).double() @ Wm.double()).float()
print("masked vs raw A bitwise equal:", same, " max |diff| %.3e" % float((outs[0] - outs[1]).abs().max()))
print("raw-A output vs truncation model: %.3e   vs round-to-nearest model: %.3e" %
      (float((outs[1] - ref_trunc).abs().max()), float((outs[1] - ref_round).abs().max())))
