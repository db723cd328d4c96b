#!/usr/bin/env python
"""E-Net training step on sharded scans (BASELINE.json config 4): forward + backward through the five BCLs
(splat / gather-conv / dgrad / wgrad kernels of this repo), one bucketed NCCL all-reduce of the gradients,
Adam step.  Launch with torchrun (one process per GPU) or as a single process.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_step.py --steps 5 --scans-per-gpu 8
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from efgh_b200 import sharding, synth
from efgh_b200.enet import Enet


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--scans-per-gpu", type=int, default=8)
    ap.add_argument("--sensor", default="os1-64-64k")   # the shipped config trains on 65 536-point clouds
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                                  # identical replicas
    args = {"dim": 3, "scale_map": synth.SCALE_MAP, "DEVICE": "cuda", "use_leaky": True, "bcn_use_bias": True,
            "bcn_use_norm": True, "last_relu": False}
    model = Enet(args).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    n_scans = a.scans_per_gpu * world
    mine = sharding.scan_indices_for_rank(n_scans, rank, world)
    clouds = [torch.from_numpy(synth.synth_scan(i, a.sensor))[None].to(dev) for i in mine]
    target = torch.tensor([0.0, 0.0, 1.0], device=dev)[None, :, None]

    def step():
        opt.zero_grad(set_to_none=True)
        total = 0.0
        for pc in clouds:                                 # batch = independent scans (the reference is batch-size-1 too)
            out = model(pc)
            loss = (1 - (out["e_gn_abs"] * target).sum()) + torch.nn.functional.cross_entropy(out["e_gn_sgn"], torch.tensor([7], device=dev))
            (loss / len(clouds)).backward()
            total += float(loss.detach())
        calls = sharding.allreduce_gradients(model.parameters(), world)
        opt.step()
        return total / len(clouds), calls

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, calls = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    chk = torch.tensor([float(sum(p.detach().double().sum() for p in model.parameters()))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool((hi - lo).abs() <= 1e-9 * hi.abs().clamp(min=1))
    else:
        in_sync = True
    if rank == 0:
        print(json.dumps({"metric": "E-Net train step (fwd+bwd+allreduce+Adam) scans/s", "value": n_scans * a.steps / (float(ms) * 1e-3),
                          "n_gpus": world, "steps": a.steps, "ms_per_step": float(ms) / a.steps, "scans_per_gpu": a.scans_per_gpu,
                          "points_per_scan": int(clouds[0].shape[-1]), "allreduce_calls_per_step": calls, "loss": loss,
                          "replicas_in_sync": in_sync}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
