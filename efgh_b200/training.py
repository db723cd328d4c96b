"""Training step of the lattice path on sharded scans (BASELINE.json configs[3]).

reference iterater.py:35-43 does, per iteration: forward -> loss -> zero_grad -> backward -> optimizer.step, on a
model wrapped in nn.DataParallel (reference main.py:127) with batch_size 1 (configs/train_rellis.yaml:29).  Here a
step is: every rank runs its own `scans_per_gpu` scans forward + backward (replicas hold identical weights; scans are
independent), the gradients are averaged with ONE bucketed all-reduce over NCCL
(efgh_b200.sharding.allreduce_gradients), then Adam steps.

Two implementations behind one interface (`step() -> loss`):

  ModulePathTrainer  the drop-in modules exactly as reference nets/enet.py wires them (GenerateData +
                     5 x BilateralConvFlex through torch.autograd), one scan per call - what an EFGH user gets by
                     swapping the two imports.  Host-bound: ~70 launches and one autograd graph
                     per scan.
  BatchedTrainer     all scans of the rank through ONE ScanPipeline launch sequence forward and one backward
                     (ScanPipeline.backward): the stem runs in torch (autograd) over the concatenated clouds, the
                     lattice / BCL part forward + backward on this repo's kernels with the batch's buffers reused
                     from step to step.
"""
import torch
import torch.nn as nn

from . import sharding, synth
from .enet import Enet

ENET_ARGS = {"dim": 3, "scale_map": synth.SCALE_MAP, "DEVICE": "cuda", "use_leaky": True, "bcn_use_bias": True,
             "bcn_use_norm": True, "last_relu": False}


def enet_loss(out, dev):
    """A stand-in for the reference's E loss (losses/efghloss.py:23: cosine + cross-entropy on the gravity normal):
    same heads, fixed target (+z, all signs positive)."""
    target = torch.tensor([0.0, 0.0, 1.0], device=dev)[None, :, None]
    return (1 - (out["e_gn_abs"] * target).sum()) + nn.functional.cross_entropy(out["e_gn_sgn"], torch.tensor([7], device=dev))


class ModulePathTrainer(object):
    def __init__(self, clouds, dev, world=1, lr=1e-4, seed=0):
        """clouds: list of (3, N) float32 CUDA tensors - this rank's scans."""
        torch.manual_seed(seed)                               # identical replicas
        self.dev, self.world = dev, world
        self.model = Enet(dict(ENET_ARGS)).to(dev)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, fused=True)    # one multi-tensor kernel instead of ~60 small ones
        self.clouds = [c[None] for c in clouds]
        self.allreduce_calls = 0

    def parameters(self):
        return list(self.model.parameters())

    def step(self):
        self.opt.zero_grad(set_to_none=True)
        total = torch.zeros((), device=self.dev)
        for pc in self.clouds:
            loss = enet_loss(self.model(pc), self.dev)
            (loss / len(self.clouds)).backward()
            total += loss.detach()
        self.allreduce_calls = sharding.allreduce_gradients(self.model.parameters(), self.world)
        self.opt.step()
        return total / len(self.clouds)


class BatchedTrainer(object):
    """Stem (torch) -> ScanPipeline forward (B scans, one launch sequence) -> loss on the last level's output ->
    ScanPipeline.backward -> stem backward (torch) -> all-reduce -> Adam.

    Loss: mean over scans of  0.5 * mean(Z_b^2)  of the last BCL's output Z_b (H_b, 256) - it has a closed-form
    gradient that one small kernel-free torch expression produces on the batch buffer, so the step contains no
    per-scan host work.  (The E-Net head is a per-scan Conv1d / BatchNorm stack outside the lattice path;
    ModulePathTrainer includes it.)"""

    def __init__(self, clouds, dev, world=1, lr=1e-4, seed=0, vertex_cap_factor=2.0, precision="3xtf32", use_graph=True):
        from .pipeline import ScanPipeline
        torch.manual_seed(seed)
        self.dev, self.world = dev, world
        B, n = len(clouds), clouds[0].shape[-1]
        assert all(c.shape[-1] == n for c in clouds)
        self.B, self.n = B, n
        with torch.cuda.device(dev):
            self.pc = torch.cat([c[:3].float() for c in clouds], dim=1).contiguous()      # (3, B*n)
            model = Enet(dict(ENET_ARGS)).to(dev)                                        # same init as the module path
            self.stem = model.conv_in
            self.bcns = nn.ModuleList([model.bcn1, model.bcn2, model.bcn3, model.bcn4, model.bcn5])
            self.params = list(self.stem.parameters()) + list(self.bcns.parameters())
            self.opt = torch.optim.Adam(self.params, lr=lr, fused=True)           # one multi-tensor kernel instead of ~60 small ones
            plan = [(m.num_input, list(m.num_output)) for m in self.bcns]
            self.pipe = ScanPipeline(n, synth.SCALE_MAP, plan, self._weights(), dev, vertex_cap_factor=vertex_cap_factor,
                                     emit_int64=False, batch=B, precision=precision, train=True)
            self._feat0 = torch.empty((plan[0][0] - 4, B * n), dtype=torch.float32, device=dev)   # static input of the captured step
        self.allreduce_calls = 0
        self._checked = False
        self.use_graph = use_graph
        self._graph = None
        self._gstream = None
        self.launches_per_step = 0

    def _weights(self):
        return [[(c.weight.detach(), c.bias.detach()) for c in m.blur_conv if isinstance(c, nn.Conv2d)] for m in self.bcns]

    def parameters(self):
        return self.params

    def _lattice_part(self):
        """Weight re-layout + lattice build + 5 BCLs forward + loss + backward on this repo's kernels (~130 launches):
        everything between the stem's forward and its backward.  Reads self._feat0, leaves d feat0 in the pipeline."""
        pipe = self.pipe
        pipe.load_weights(self._weights())                 # re-lay the updated weights (device-side, no sync)
        pipe.enqueue(self.pc, self._feat0)
        self._loss_dev, dZ = pipe.loss_half_mean_square()   # per-scan means, device-side row counts
        self._dfeat0 = pipe.backward(dZ)

    def step(self):
        pipe = self.pipe
        with torch.cuda.device(self.dev):
            self.opt.zero_grad(set_to_none=True)
            feat0 = self.stem(self.pc[None])[0]                # (32, B*n), autograd
            self._feat0.copy_(feat0.detach())
            if not self._checked:
                # first step (eager, synchronising once): capacities and the symmetry backward() relies on
                from . import _capi
                l0 = _capi.lib().efgh_launch_count()
                self._lattice_part()
                self.launches_per_step = int(_capi.lib().efgh_launch_count() - l0)   # this library's kernels in one step (a replayed graph re-launches them)
                pipe.counts()
                bad = pipe.aliased_levels()
                if bad:
                    raise RuntimeError("BatchedTrainer: neighbour tables of levels %s are not mirror-symmetric (aliased keys); "
                                       "use ModulePathTrainer for such clouds" % bad)
                self._checked = True
            elif not self.use_graph:
                self._lattice_part()
            else:
                if self._graph is None:
                    # capture once (the launch sequence is data-independent: every count stays on the device), then replay:
                    # one graph launch instead of ~130 kernel launches + ~60 small torch ops per step
                    self._gstream = torch.cuda.Stream(self.dev)
                    self._gstream.wait_stream(torch.cuda.current_stream(self.dev))
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._gstream):
                        self._lattice_part()
                    torch.cuda.current_stream(self.dev).wait_stream(self._gstream)
                    self._graph = g
                self._graph.replay()
            feat0.backward(self._dfeat0)
            for m, g in zip(self.bcns, pipe.weight_grads()):
                convs = [c for c in m.blur_conv if isinstance(c, nn.Conv2d)]
                for c, (gw, gb) in zip(convs, g):
                    # (the pipeline's gradients are permuted VIEWS of its (K, M) buffers; the optimizer's multi-tensor kernels
                    #  need the parameter's own layout - with strided gradients torch falls back to ~60 per-tensor kernels)
                    c.weight.grad, c.bias.grad = gw.contiguous(), gb
            self.allreduce_calls = sharding.allreduce_gradients(self.params, self.world)
            self.opt.step()
        return self._loss_dev
