"""(needs a library built with the trace hooks compiled in: make -C efgh_b200/csrc clean all EXTRA=-DEFGH_CONV_TRACE)
Debug: timeline of the tensor-core conv roles in CTA 0 (not collected by pytest)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efgh_b200 import _capi
L = _capi.lib()
dev = torch.device("cuda:0")
CAP = 512
def run(H, C, F, M, label, flags=0, show=0):
    L.efgh_debug_set_conv_flags(flags)
    X = torch.randn(H + 1, C, device=dev)
    nbr = torch.randint(-1, H, (F, H), device=dev, dtype=torch.int32) if F > 1 else None
    Wt = torch.randn(F * C, M, device=dev) * 0.1
    img = torch.empty(L.efgh_bcl_packed_weight_bytes(F * C, M, 3) // 4, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _capi.check(L.efgh_bcl_pack_weights(Wt.data_ptr(), F * C, M, 3, img.data_ptr(), st), "pack")
    Y = torch.zeros(H, M, device=dev)
    trace = torch.zeros(5 * CAP * 2, dtype=torch.int64, device=dev)
    groups = L.efgh_bcl_conv_tc_groups(F * C, M)
    for rep in range(3):
        trace.zero_()
        L.efgh_debug_set_conv_trace.argtypes = [ctypes.c_void_p]
        L.efgh_debug_set_conv_trace(trace.data_ptr() if rep == 2 else None)
        _capi.check(L.efgh_bcl_conv_tc(X.data_ptr(), C, C, None, 0, nbr.data_ptr() if nbr is not None else None, 32, H, F, H, None,
                                       img.data_ptr(), None, M, 0, Y.data_ptr(), M, 3, 1 if groups > 1 else 0, st), "conv")
        torch.cuda.synchronize()
    L.efgh_debug_set_conv_trace(None)
    t = trace.cpu().view(5, CAP, 2)
    t0 = min(int(t[r, 0, 1]) for r in range(5) if int(t[r, 0, 1]) > 0)
    print("==== %s H=%d C=%d F=%d M=%d groups=%d" % (label, H, C, F, M, groups))
    names = ["P0", "P1", "TMA", "MMA", "EPI"]
    evs = []
    for r in range(5):
        for i in range(CAP):
            if int(t[r, i, 1]) == 0: break
            evs.append((int(t[r, i, 1]) - t0, names[r], int(t[r, i, 0])))
    evs.sort()
    for tt, nm, ev in evs[:show]:
        print("%8.2f us  %-4s %d" % (tt / 1000.0, nm, ev))
    print("last event at %.2f us, %d events" % (evs[-1][0] / 1000.0, len(evs)))
run(100654, 36, 15, 32, "L0 conv1 default", show=150)
run(62551, 36, 15, 64, "L1 conv1")
run(23050, 68, 15, 128, "L2 conv1")
run(4194, 132, 15, 256, "L3 conv1", show=60)
run(885, 260, 15, 256, "L4 conv1")
run(100654, 32, 1, 32, "L0 conv2")
run(885, 256, 1, 256, "L4 conv2", show=60)
L.efgh_debug_set_conv_flags(0)
