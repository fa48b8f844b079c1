"""Ray-sharded data parallelism: one process per GPU, rays sharded,
weights + hash tables + occupancy grid replicated, the gradients of all parameters live in one
flat fp32 buffer that is averaged over ranks once per step (NCCL over NVLink 5 / NVSwitch; gloo
on CPU for the tests).

The reference has no distributed code at all (run.py:59 picks one device); this is the
new multi-GPU path of the hot loop.  ``clip_grad_norm_`` and the optimizer must run AFTER
``allreduce()`` so every rank clips and steps identically.

Overlap: the large parameters (hash tables: 50-100 MB each) sit at the END of the flat buffer and
are all-reduced asynchronously from a post-accumulate-grad hook, i.e. as soon as autograd has
finished their gradient -- for the dynamic models that is while the backward of the deformation
branch is still running.  ``allreduce()`` then reduces the small head of the buffer and waits.
"""
import torch
import torch.distributed as dist


class GradAllReducer:
    """Makes every ``p.grad`` a view into one contiguous fp32 buffer, so that zeroing the
    gradients is one memset and averaging them over ranks is one collective per large table plus
    one for everything else."""

    def __init__(self, module: torch.nn.Module, world_size: int = None, overlap: bool = True,
                 big_numel: int = 1 << 22):
        params = [p for p in module.parameters() if p.requires_grad]
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        small = [p for p in params if p.numel() < big_numel]
        big = [p for p in params if p.numel() >= big_numel]
        self.params = small + big
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, device=ref.device, dtype=torch.float32)
        off = 0
        self._views = {}
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._views[p] = self.flat[off:off + n]
            off += n
        self.n_small = sum(p.numel() for p in small)
        self.nbytes = total * 4
        self._pending = []
        self._hooks = []
        self.overlap = bool(overlap and big and self.world > 1 and dist.is_initialized())
        if self.overlap:
            for p in big:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))

    # ---- collectives
    def _reduce(self, t, async_op=False):
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        work = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)      # gloo has no AVG
        return work

    def _on_grad_ready(self, p):
        if p.grad is None or p.grad.data_ptr() != self._views[p].data_ptr():
            p.grad = None if p.grad is None else self._views[p].copy_(p.grad.reshape(-1)).view_as(p)
        self._pending.append((self._reduce(self._views[p], async_op=True), self._views[p]))

    def zero_grad(self):
        self.flat.zero_()

    def _rebind(self):
        """A caller that ran ``optimizer.zero_grad()`` (set_to_none) instead of ``reducer.zero_grad()`` made autograd
        allocate fresh .grad tensors: copy them back into the flat buffer and re-point .grad, so that the collective
        below never reduces stale data."""
        for p in self.params:
            v = self._views[p]
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad.reshape(-1))
            else:
                continue
            p.grad = v.view_as(p)

    def allreduce(self):
        if self.world <= 1 or not dist.is_initialized():
            return
        if not self._pending:          # (with overlap the hooks already re-bound the large tables they reduced)
            self._rebind()
        else:
            for p in self.params[: len(self.params) - len(self._hooks)]:
                v = self._views[p]
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad.reshape(-1))
                    p.grad = v.view_as(p)
        avg_in_op = dist.get_backend() == "nccl"
        if self.overlap:
            head = self.flat[: self.n_small]
            if head.numel():
                self._reduce(head)
                if not avg_in_op:
                    head.div_(self.world)
            for work, view in self._pending:
                work.wait()
                if not avg_in_op:
                    view.div_(self.world)
            self._pending = []
        else:
            self._reduce(self.flat)
            if not avg_in_op:
                self.flat.div_(self.world)

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_rays(n_rays: int, rank: int, world: int):
    """Contiguous [start, stop) slice of a global ray batch owned by ``rank``."""
    per = (n_rays + world - 1) // world
    start = min(rank * per, n_rays)
    return start, min(start + per, n_rays)
