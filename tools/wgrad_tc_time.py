"""Timing study of k_wgrad_tc (not a test): the ten E-Net weight-gradient shapes of a training step (8 scans of 65 536
points), optionally with parts of the kernel switched off (efgh_debug_set_wgrad_flags; results are wrong then)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efgh_b200 import _capi

dev = torch.device("cuda:0")
L = _capi.lib()
L.efgh_debug_set_wgrad_flags.argtypes = [ctypes.c_int]
L.efgh_debug_set_wgrad_flags.restype = None
st = torch.cuda.current_stream().cuda_stream
scans = int(sys.argv[1]) if len(sys.argv) > 1 else 8
FLAGS = [int(v, 0) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0]
# (H per scan, C, F, M): conv1 (gathered) and conv2 (1x1) of every level
SHAPES = [(52000, 36, 15, 32), (52000, 32, 1, 32), (35000, 36, 15, 64), (35000, 64, 1, 64), (14000, 68, 15, 128), (14000, 128, 1, 128),
          (2800, 132, 15, 256), (2800, 256, 1, 256), (620, 260, 15, 256), (620, 256, 1, 256)]
for fl in FLAGS:
    L.efgh_debug_set_wgrad_flags(fl)
    tot = 0.0
    for (H, C, F, M) in SHAPES:
        H *= scans
        X = torch.randn(H + 1, C, device=dev); X[0] = 0
        nbr = torch.randint(-1, H, (F, H), device=dev, dtype=torch.int32) if F > 1 else None
        G = torch.randn(H, M, device=dev)
        dW = torch.zeros(F * C, M, device=dev)
        db = torch.zeros(M, device=dev)
        def run():
            _capi.check(L.efgh_bcl_conv_wgrad_tc(X.data_ptr(), C, C, nbr.data_ptr() if nbr is not None else None, 32, H, F, H, None, G.data_ptr(), M, M,
                                                 dW.data_ptr(), db.data_ptr(), st), "wgrad")
        for _ in range(2):
            run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            run()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 5 * 1e3
        tot += us
        n_kt = (F * C + 127) // 128
        stages = (H + 31) // 32 * n_kt * (2 if M > 128 else 1)
        print("flags=0x%x H=%d C=%d F=%d M=%d: %.1f us  (%.2f us per stage per SM)" % (fl, H, C, F, M, us, us * 148 / stages), flush=True)
    print("flags=0x%x total %.1f us" % (fl, tot), flush=True)
L.efgh_debug_set_wgrad_flags(0)
