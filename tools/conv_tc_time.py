"""Timing study of k_conv_tc (not a test): E-Net conv1 shapes, 3xTF32 vs one TF32 pass, K-group length override."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efgh_b200 import _capi

dev = torch.device("cuda:0")
L = _capi.lib()
L.efgh_debug_set_conv_flags.argtypes = [ctypes.c_int]
L.efgh_debug_set_conv_flags.restype = None
st = torch.cuda.current_stream().cuda_stream
scans = int(sys.argv[1]) if len(sys.argv) > 1 else 1
FLAGS = [int(v, 0) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0]
NS = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [3]
SHAPES = [(100654, 36, 15, 32), (62551, 36, 15, 64), (23050, 68, 15, 128), (4194, 132, 15, 256), (885, 260, 15, 256), (100654, 32, 1, 32)]
if len(sys.argv) > 4:      # extra shapes: "H,C,F,M;H,C,F,M" (e.g. the aligned-row variants 32 / 64 channels of levels 0-2)
    SHAPES = [tuple(int(v) for v in t.split(',')) for t in sys.argv[4].split(';')]
for (H, C, F, M) in SHAPES:
    H *= scans
    X = torch.randn(H + 1, C, device=dev); X[0] = 0
    nbr = torch.randint(-1, H, (F, H), device=dev, dtype=torch.int32) if F > 1 else None
    Wt = torch.randn(F * C, M, device=dev) * 0.1
    Y = torch.zeros(H, M, device=dev)
    for ns in NS:
        for gc in FLAGS:
            L.efgh_debug_set_conv_flags(gc)
            img = torch.empty(L.efgh_bcl_packed_weight_bytes(F * C, M, ns) // 4, device=dev)
            _capi.check(L.efgh_bcl_pack_weights(Wt.data_ptr(), F * C, M, ns, img.data_ptr(), st), "pack")
            acc = 1 if L.efgh_bcl_conv_tc_groups(F * C, M) > 1 else 0
            Y.zero_()
            def run():
                _capi.check(L.efgh_bcl_conv_tc(X.data_ptr(), C, C, None, 0, nbr.data_ptr() if nbr is not None else None, 32, H, F, H, None,
                                               img.data_ptr(), None, M, 0, Y.data_ptr(), M, ns, acc, st), "conv")
            for _ in range(3):
                run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(20):
                run()
            b.record()
            torch.cuda.synchronize()
            print("H=%d C=%d F=%d M=%d nsplit=%d group_chunks=%s groups=%d: %.1f us" %
                  (H, C, F, M, ns, "flags=0x%x" % gc, L.efgh_bcl_conv_tc_groups(F * C, M), a.elapsed_time(b) / 20 * 1e3), flush=True)
L.efgh_debug_set_conv_flags(0)
