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


class GradSink:
    """Direct gradient accumulation target of one hash table under data parallelism (b2n.dp.GradAllReducer).

    ``view`` is the table's slice of the reducer's flat gradient buffer (also ``table.grad``).  The table-gradient
    kernels accumulate straight into it -- no zero-filled temporary, no autograd accumulate pass -- and tell the
    reducer which element ranges are final, so that their all-reduce starts while the rest of the backward still
    runs.  ``uses`` counts the forward passes of the current step that will send a gradient into the table (run.py's
    regularisers call the encoders directly, besides render_rays): only the LAST backward announces ranges."""

    def __init__(self, view: torch.Tensor, on_ready):
        self.view, self.on_ready, self.uses, self.split_level = view, on_ready, 0, 8

    def reset(self):
        self.uses = 0


class GradAllReducer:
    """Makes every ``p.grad`` a view into one contiguous fp32 buffer, so that zeroing the
    gradients is one memset and averaging them over ranks is a handful of collectives over element RANGES of that
    buffer.  Ranges are reduced asynchronously as soon as they are final:

    * ``direct=True``: hash tables (flat 1-D parameters of >= ``big_numel`` elements that reach ``b2n.hash_encode``)
      get a ``GradSink`` -- the table-gradient kernels accumulate straight into the flat buffer and announce the fine
      levels' slice before the coarse levels are scattered (ops._HashEncode.backward).  CONTRACT: such a table receives
      gradient ONLY through ``b2n.hash_encode``; a loss term that touches the table through autograd (the reference's
      ``mean|params[1:] - params[:-1]|`` TV term, run.py:614-616) would add into a range whose reduction is already in
      flight -- it is refused with an error.  Put the TV term into ``b2n.optim.FusedAdamW(tv_weight=...)`` instead, or
      keep the default;
    * ``direct=False`` (default): every large parameter is reduced from its post-accumulate-grad hook, which autograd
      runs once all contributions to that parameter -- direct kernels and autograd alike -- are in;
    * ``allreduce()`` reduces whatever range is still untouched, then waits for everything.

    ONE backward per step reaches the hooks; wrap earlier backward passes of an accumulation step in ``no_sync()``."""

    def __init__(self, module: torch.nn.Module, world_size: int = None, overlap: bool = True,
                 big_numel: int = 1 << 22, direct: bool = False):
        params = [p for p in module.parameters() if p.requires_grad]
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        small = [p for p in params if p.numel() < big_numel]
        big = [p for p in params if p.numel() >= big_numel]
        self.params = small + big
        # every parameter's slice starts on a 256-byte boundary: the kernels that write gradients in place (vector
        # red.global of the hash-table gradient, the 16-byte accesses of FusedAdamW) need aligned bases
        ALIGN = 64
        offsets, off = [], 0
        for p in self.params:
            offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        total = off
        ref = self.params[0]
        self.flat = torch.zeros(total, device=ref.device, dtype=torch.float32)
        self._views, self._offset = {}, {}
        for p, off in zip(self.params, offsets):
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self._views[p] = self.flat[off:off + n]
            self._offset[p] = off
        self.n_small = offsets[len(small)] if big else total
        self.nbytes = total * 4
        self._pending = []          # (work, start, stop) of ranges handed to async collectives this step
        self._hooks = []
        self._sinks = []
        self._suspended = False
        self.overlap = bool(overlap and big and self.world > 1 and dist.is_initialized())
        if self.overlap:
            # only hash tables (the flat parameter of a HashGridEncoding: reached through b2n.hash_encode /
            # hash_tri_blend, whose backward implements the sink protocol) can take gradients in place
            tables = {id(m.params) for m in module.modules() if type(m).__name__ == "HashGridEncoding"}
            for p in big:
                if direct == "always" or (direct and id(p) in tables and p.is_cuda):
                    sink = GradSink(self._views[p], self._on_sink_ready)
                    sink.offset = self._offset[p]
                    p._b2n_grad_sink = sink
                    self._sinks.append((p, sink))
                    # a tensor hook sees every DEFINED gradient autograd delivers to the leaf (the direct path returns
                    # None): that is a contribution outside the contract -- refuse it rather than race with the
                    # reduction that may already be in flight
                    self._hooks.append(p.register_hook(self._refuse_autograd_gradient))
                    continue
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))

    # ---- collectives
    def _reduce(self, t, async_op=False):
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        work = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op)      # gloo has no AVG
        return work

    def _start(self, start, stop):
        if stop > start:
            self._pending.append((self._reduce(self.flat[start:stop], async_op=True), start, stop))

    def _covered(self, start, stop):
        return any(a < stop and start < b for _, a, b in self._pending)

    def _on_sink_ready(self, sink, lo, hi):
        if self._suspended:
            return
        if self._covered(sink.offset + lo, sink.offset + hi):
            raise RuntimeError("GradAllReducer: a gradient range was announced twice in one step (two backward passes "
                               "without reducer.no_sync() around the first?)")
        self._start(sink.offset + lo, sink.offset + hi)

    @staticmethod
    def _refuse_autograd_gradient(grad):
        if grad is None:            # what the direct path returns: nothing came through autograd
            return None
        raise RuntimeError("GradAllReducer(direct=True): a hash table with a direct gradient sink received a gradient "
                           "through autograd (a TV / regularisation term on the raw table in the loss?).  Move that "
                           "term into b2n.optim.FusedAdamW(tv_weight=...) or build the reducer with direct=False")

    def _on_grad_ready(self, p):
        """Hook path (a large parameter whose gradient came through autograd's accumulation, e.g. a hash table used by
        a foreign op).  A second backward before ``allreduce()`` would let autograd add into a view whose NCCL
        reduction is still in flight: detected and refused; use ``no_sync()`` around all but the last backward."""
        if self._suspended:
            return
        v, off = self._views[p], self._offset[p]
        if self._covered(off, off + v.numel()):
            raise RuntimeError("GradAllReducer: a second backward reached a parameter whose gradient all-reduce is "
                               "still in flight; wrap all but the last backward of a step in reducer.no_sync()")
        if p.grad is None or p.grad.data_ptr() != v.data_ptr():
            p.grad = None if p.grad is None else v.copy_(p.grad.reshape(-1)).view_as(p)
        self._start(off, off + v.numel())

    def no_sync(self):
        """Context manager: backward passes inside it only accumulate (no collective is started), like DDP.no_sync()."""
        reducer = self

        class _NoSync:
            def __enter__(self_inner):
                reducer._suspended = True

            def __exit__(self_inner, *exc):
                reducer._suspended = False
                return False
        return _NoSync()

    def zero_grad(self):
        self.flat.zero_()
        for _, sink in self._sinks:
            sink.reset()

    def _rebind(self):
        """A caller that ran ``optimizer.zero_grad()`` (set_to_none) instead of ``reducer.zero_grad()`` made autograd
        allocate fresh .grad tensors: copy them back into the flat buffer and re-point .grad, so that the collective
        below never reduces stale data.  (Ranges already in flight were re-bound by their hook.)"""
        for p in self.params:
            v, off = self._views[p], self._offset[p]
            if self._covered(off, off + v.numel()):
                continue
            if p.grad is None:
                if getattr(p, "_b2n_grad_sink", None) is None:      # a sink accumulates in place: the view IS the gradient
                    v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad.reshape(-1))
            else:
                continue
            p.grad = v.view_as(p)

    def allreduce(self):
        if self.world <= 1 or not dist.is_initialized():
            return
        self._rebind()
        avg_in_op = dist.get_backend() == "nccl"
        # every range not yet in flight, in address order
        done = sorted((a, b) for _, a, b in self._pending)
        pos, n = 0, self.flat.numel()
        gaps = []
        for a, b in done:
            if a > pos:
                gaps.append((pos, a))
            pos = max(pos, b)
        if pos < n:
            gaps.append((pos, n))
        for a, b in gaps:
            self._start(a, b)
        for work, a, b in self._pending:
            work.wait()
            if not avg_in_op:
                self.flat[a:b].div_(self.world)
        self._pending = []
        for _, sink in self._sinks:
            sink.reset()

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for p, _ in self._sinks:
            p._b2n_grad_sink = None
        self._sinks = []


def shard_rays(n_rays: int, rank: int, world: int):
    """Contiguous [start, stop) slice of a global ray batch owned by ``rank``."""
    per = (n_rays + world - 1) // world
    start = min(rank * per, n_rays)
    return start, min(start + per, n_rays)
