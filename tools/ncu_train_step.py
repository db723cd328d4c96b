#!/usr/bin/env python
"""One BatchedTrainer step (8 x 65 536-point scans) inside a cudaProfilerStart/Stop range, for ncu launch lists:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
        python tools/ncu_train_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from efgh_b200 import synth, training


def main():
    dev = torch.device("cuda:0")
    scans = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    clouds = [torch.from_numpy(synth.synth_scan(i, "os1-64-64k")).to(dev) for i in range(scans)]
    tr = training.BatchedTrainer(clouds, dev, use_graph=False)     # eager launches: ncu lists every kernel of the step
    with torch.no_grad():
        for p in tr.params:
            p.normal_(0, 0.1 if p.dim() > 1 else 0.05)
    for _ in range(2):
        tr.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
