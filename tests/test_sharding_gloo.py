"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (scan sharding + gradient all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from efgh_b200 import sharding


def test_scan_sharding_partitions_every_scan_once():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            idx = sharding.scan_indices_for_rank(67, r, world)
            assert all(sharding.owner_of_scan(i, world) == r for i in idx)
            seen += idx
        assert sorted(seen) == list(range(67))
    with pytest.raises(ValueError):
        sharding.scan_indices_for_rank(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    # same weights on every rank (replicas), different per-rank gradients (different scans)
    shapes = [(32, 36, 15, 1), (32,), (32, 32, 1, 1), (32,), (7,)]
    params = [torch.nn.Parameter(torch.randn(*s)) for s in shapes]
    for i, p in enumerate(params):
        if i != 4 or rank == 0:                       # last parameter has no grad on rank 1
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    calls = sharding.allreduce_gradients(params, bucket_bytes=40000)
    want = [((1 + 2) / 2.0) * (i + 1) for i in range(4)] + [5.0 / 2]
    ok = all(torch.allclose(p.grad, torch.full_like(p, w)) for p, w in zip(params, want))
    mine = sharding.scan_indices_for_rank(9, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((ok, calls, sorted(sum(gathered, []))))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, calls, scans = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "averaged gradients differ from the expected mean"
    assert calls >= 2, "bucketing should have split the parameters into several all-reduces"
    assert scans == list(range(9))
