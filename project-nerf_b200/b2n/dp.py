"""Ray-sharded data parallelism: one process per GPU, rays sharded,
weights + hash tables + occupancy grid replicated, ONE all-reduce per step over a single
flat gradient buffer (NCCL over NVLink 5 / NVSwitch; gloo on CPU for the tests).

The reference has no distributed code at all (run.py:59 picks one device); this is the
new multi-GPU path of the hot loop.  ``clip_grad_norm_`` and the optimizer must run AFTER
``allreduce()`` so every rank clips and steps identically.
"""
import torch
import torch.distributed as dist


class GradAllReducer:
    """Makes every ``p.grad`` a view into one contiguous fp32 buffer, so that zeroing the
    gradients is one memset and averaging them over ranks is one collective."""

    def __init__(self, module: torch.nn.Module, world_size: int = None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, device=ref.device, dtype=torch.float32)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self.nbytes = total * 4

    def zero_grad(self):
        self.flat.zero_()

    def allreduce(self):
        if self.world <= 1 or not dist.is_initialized():
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                                   # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world)


def shard_rays(n_rays: int, rank: int, world: int):
    """Contiguous [start, stop) slice of a global ray batch owned by ``rank``."""
    per = (n_rays + world - 1) // world
    start = min(rank * per, n_rays)
    return start, min(start + per, n_rays)
