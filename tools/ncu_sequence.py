#!/usr/bin/env python
"""One eager launch sequence of the bench's pipeline inside a cudaProfilerStart/Stop range - the target of bench.py's
in-run `ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum` pass (per-stage DRAM traffic)
and of the `ncu --set full` captures under profiles/.

    ncu --profile-from-start off --set full -k regex:k_conv_tc -o gpurun_out/conv python tools/ncu_sequence.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from efgh_b200 import synth
from efgh_b200.pipeline import ScanPipeline, make_enet_weights


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan-batch", type=int, default=16)
    ap.add_argument("--sensor", default="os1-64")
    ap.add_argument("--no-stem", action="store_true")
    ap.add_argument("--int32-only", action="store_true")
    ap.add_argument("--atomic-splat", action="store_true")
    ap.add_argument("--passes", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    G = a.scan_batch
    clouds = [synth.synth_scan(s, a.sensor) for s in range(G)]
    N = clouds[0].shape[1]
    weights = make_enet_weights(synth.ENET_BCL)
    gs_ = torch.Generator().manual_seed(5)
    stem = ([(torch.randn(co, ci, 1, generator=gs_) * 0.3, torch.randn(co, generator=gs_) * 0.1) for ci, co in ((3, 32), (32, 32), (32, 32))], True)
    pipe = ScanPipeline(N, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=1.0, batch=G, gather_splat=not a.atomic_splat,
                        stem=None if a.no_stem else stem, emit_int64=not a.int32_only)
    pipe.overlap_lattice = False                       # one stream: launches appear in stage order
    pc = torch.from_numpy(np.concatenate(clouds, axis=1)).to(dev)
    ft = torch.randn(32, G * N, device=dev) if a.no_stem else None
    for _ in range(2):
        pipe.enqueue(pc, ft)
    torch.cuda.synchronize()
    pipe.counts()
    torch.cuda.profiler.start()
    for _ in range(a.passes):
        pipe.enqueue(pc, ft)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
