"""Multi-GPU layout of the hot path: scans are independent units.

One process per GPU.  Inference / lattice build: scan i goes to rank i mod world, every rank runs whole
scans end to end and there is NO collective on the data path.  Training (BASELINE.json config 4): replicas
hold identical weights and the only exchange is one all-reduce (mean) of the gradients per step, bucketed
into flat buffers so that NCCL sees a few large messages over NVLink / NVSwitch instead of one per
parameter.  The reference wraps the model in nn.DataParallel (reference main.py:127), which with its
batch_size 1 is a single GPU; this replaces it.
"""
import torch
import torch.distributed as dist


def scan_indices_for_rank(n_scans, rank, world):
    """Global scan indices owned by `rank` (round-robin: scan i -> rank i mod world)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_scans, world))


def owner_of_scan(scan_index, world):
    return scan_index % world


def allreduce_gradients(params, world=None, bucket_bytes=32 << 20, group=None):
    """Average .grad of `params` across ranks with bucketed flat all-reduces.  Works for NCCL (CUDA tensors)
    and gloo (CPU tensors - used by the CPU tests).  Parameters without a gradient are treated as zero so that
    every rank issues the same sequence of collectives.  Returns the number of all-reduce calls issued."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    params = [p for p in params if p.requires_grad]
    if world == 1 or not params:
        return 0
    calls = 0
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size, calls
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        o = 0
        for p in bucket:
            n = p.numel()
            g = flat[o:o + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            o += n
        calls += 1
        bucket, size = [], 0

    for p in params:
        nbytes = p.numel() * p.element_size()
        if bucket and (size + nbytes > bucket_bytes or p.dtype != bucket[0].dtype or p.device != bucket[0].device):
            flush()
        bucket.append(p)
        size += nbytes
    flush()
    return calls
