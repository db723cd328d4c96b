"""torchrun worker of tests/test_gpu_backward.py::test_gradient_allreduce_nccl (one process per GPU, NCCL)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from efgh_b200 import sharding


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    shapes = [(32, 36, 15, 1), (32,), (256, 260, 15, 1), (256,), (7,)]
    params = [torch.nn.Parameter(torch.randn(*s, device=dev)) for s in shapes]
    for i, p in enumerate(params):
        if i != 4 or rank == 0:                        # last parameter has no grad except on rank 0
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    calls = sharding.allreduce_gradients(params, bucket_bytes=1 << 20)
    mean_rank = sum(range(1, world + 1)) / world
    want = [mean_rank * (i + 1) for i in range(4)] + [5.0 / world]
    ok = all(torch.allclose(p.grad, torch.full_like(p, w)) for p, w in zip(params, want))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and float(flag) == 1.0 and calls >= 2:
        print("NCCL_ALLREDUCE_OK calls=%d world=%d" % (calls, world))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
