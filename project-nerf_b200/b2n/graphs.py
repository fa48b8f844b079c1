"""CUDA-graph capture of fixed-shape steps of the hot path.

The small-batch configurations (vanilla NeRF at 4096 rays x 64 samples: ~40 kernels of 5-500 us each) are bound by
launch and Python overhead, not by the kernels.  When every shape of a step is static -- no occupancy grid, so the
number of evaluated samples is B * N -- the whole step (march, encode, tcgen05 decoder, composite, loss, backward,
optimizer) can be captured once and replayed with one launch.  Steps behind an occupancy grid become capturable with
``b2n.march.set_static_capacity(True)``: capacity-sized compact buffers and a device-side row count that every kernel
honours (``b2n_set_active_rows``) replace the 4-byte host read of the active-sample count.

The abort flags of tcgen05 launches captured into a graph are re-zeroed and re-written by every replay;
``b2n.check_errors()`` inspects them (state of the latest replay) together with the eager launches' flags.

All kernels of this package are launched on ``torch.cuda.current_stream()`` and allocate through torch, so
``torch.cuda.graph`` captures them like any torch op; tensor maps (TMA) are encoded on the host at capture time from
addresses inside the graph's private memory pool, which stay valid for every replay.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    """``fn(*tensors)`` captured into a CUDA graph.  ``fn`` must be free of host synchronisation and of shapes that
    depend on data; for a training step the optimizer must be capturable (``torch.optim.Adam(..., capturable=True)``)
    and ``fn`` should call ``optimizer.zero_grad(set_to_none=True)`` itself.

        step = GraphedStep(train_step, (rays_o, rays_d, target))
        loss = step(rays_o, rays_d, target)        # copies the inputs into the static buffers, replays

    The returned tensors are static too: read them (or ``.clone()``) before the next call."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        if not all(t.is_cuda for t in example_inputs):
            raise ValueError("GraphedStep needs CUDA tensors: there is no CPU path in this package")
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the default stream (lazy inits, autotuning,
            for _ in range(warmup):                        # the cluster-occupancy query of the tcgen05 decoder, ...)
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import ops
        n0 = len(ops._GRAPH_FLAGS)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_outputs = fn(*self.static_inputs)
        # abort flags of the tcgen05 launches inside the graph: owned here (freed with the graph), read by b2n.check_errors()
        self._err_flags = ops._GRAPH_FLAGS[n0:]
        del ops._GRAPH_FLAGS[n0:]
        ops._GRAPH_OWNERS.add(self)

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs
