"""world_size-2 gloo test (CPU) of the data-parallel plumbing: ray sharding + flat-buffer
gradient averaging must reproduce the single-process gradient of the mean-over-rays loss."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_dp():
    # b2n/__init__ loads the CUDA library (present after build()); dp.py itself is pure torch
    import importlib.util
    spec = importlib.util.spec_from_file_location("b2n_dp", os.path.join(ROOT, "project-nerf_b200", "b2n", "dp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _make():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.ReLU(), torch.nn.Linear(32, 3))


BIG = 100          # parameters with >= BIG elements take the asynchronous (hook-driven) all-reduce path: the 6x32 weight


def _worker(rank, world, port, out, foreign_zero_grad=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = _load_dp()
    model = _make()
    red = dp.GradAllReducer(model, world, overlap=True, big_numel=BIG)
    assert red.overlap and len(red._hooks) == 1
    torch.manual_seed(1)
    rays, target = torch.randn(64, 6), torch.randn(64, 3)
    a, b = dp.shard_rays(64, rank, world)
    if foreign_zero_grad:
        for p in model.parameters():          # what optimizer.zero_grad(set_to_none=True) does: .grad leaves the flat buffer
            p.grad = None
    else:
        red.zero_grad()
    torch.nn.functional.mse_loss(model(rays[a:b]), target[a:b]).backward()
    red.allreduce()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.05)          # after the all-reduce: identical on all ranks
    if rank == 0:
        torch.save({n: p.grad.clone() for n, p in model.named_parameters()}, out)
    gathered = [torch.zeros_like(red.flat) for _ in range(world)]
    dist.all_gather(gathered, red.flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    dist.destroy_process_group()


def test_flat_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    dp = _load_dp()
    model = _make()
    red = dp.GradAllReducer(model, 1)
    torch.manual_seed(1)
    rays, target = torch.randn(64, 6), torch.randn(64, 3)
    torch.nn.functional.mse_loss(model(rays), target).backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.05)
    got = torch.load(out)
    for n, p in model.named_parameters():
        assert torch.allclose(got[n], p.grad, rtol=1e-5, atol=1e-7), n
    # p.grad are views of the flat buffer (zeroing is one memset)
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in model.parameters())


def test_allreduce_survives_a_foreign_zero_grad(tmp_path):
    """optimizer.zero_grad() (set_to_none) detaches .grad from the flat buffer; allreduce() must re-bind, not reduce stale zeros"""
    out = str(tmp_path / "g2.pt")
    mp.spawn(_worker, args=(2, 29519, out, True), nprocs=2, join=True)
    model = _make()
    torch.manual_seed(1)
    rays, target = torch.randn(64, 6), torch.randn(64, 3)
    torch.nn.functional.mse_loss(model(rays), target).backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.05)
    got = torch.load(out)
    for n, p in model.named_parameters():
        assert torch.allclose(got[n], p.grad, rtol=1e-5, atol=1e-7), n


def test_shard_rays_covers_batch():
    dp = _load_dp()
    for n, w in ((64, 2), (65, 8), (7, 8), (0, 4)):
        spans = [dp.shard_rays(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


# ---------------------------------------------------------------- direct (sink) path: ranges announced by the backward
class _SinkScale(torch.autograd.Function):
    """y = x * sum(table): mimics ops._HashEncode's data-parallel protocol on CPU -- the table gradient is accumulated
    straight into the reducer's flat buffer in two halves, each announced as soon as it is final"""

    @staticmethod
    def forward(ctx, x, table, sink):
        ctx.save_for_backward(x, table)
        ctx.sink = sink
        sink.uses += 1
        return x * table.sum()

    @staticmethod
    def backward(ctx, g):
        x, table = ctx.saved_tensors
        sink = ctx.sink
        gt = (g * x).sum() * torch.arange(1, table.numel() + 1, dtype=torch.float32)   # a position-dependent "gradient"
        sink.uses -= 1
        cut = table.numel() // 2
        sink.view[cut:] += gt[cut:]
        if sink.uses <= 0:
            sink.on_ready(sink, cut, table.numel())
        sink.view[:cut] += gt[:cut]
        if sink.uses <= 0:
            sink.on_ready(sink, 0, cut)
        return g * table.sum(), None, None


class _SinkModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.table = torch.nn.Parameter(torch.randn(300))
        self.lin = torch.nn.Linear(4, 4)

    def forward(self, x):
        return _SinkScale.apply(self.lin(x), self.table, self.table._b2n_grad_sink)


def _sink_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = _load_dp()
    model = _SinkModel()
    red = dp.GradAllReducer(model, world, overlap=True, big_numel=BIG, direct="always")
    assert len(red._sinks) == 1
    torch.manual_seed(1)
    x = torch.randn(32, 4)
    a, b = dp.shard_rays(32, rank, world)
    for step in range(2):                      # two steps: the use counter and the pending list must reset
        red.zero_grad()
        # two uses of the table in one step (run.py's regularisers call the encoders besides render_rays): only the
        # last backward may announce ranges
        loss = model(x[a:b]).pow(2).mean() + 0.1 * model(x[a:b] * 0.5).abs().mean()
        loss.backward()
        assert len(red._pending) == 2 and model.table.grad.data_ptr() == red._views[model.table].data_ptr()
        red.allreduce()
        assert not red._pending
    if rank == 0:
        torch.save(red.flat.clone(), out)
    # accumulation: the first backward under no_sync() starts nothing; a second backward without it is refused
    red.zero_grad()
    with red.no_sync():
        model(x[a:b]).pow(2).mean().backward()
    assert not red._pending
    model(x[a:b]).pow(2).mean().backward()
    with pytest.raises(RuntimeError, match="no_sync"):
        model(x[a:b]).pow(2).mean().backward()
    red.allreduce()
    # the contract of the direct path: no autograd-side gradient on a table with a sink
    red.zero_grad()
    with pytest.raises(RuntimeError, match="direct=True"):
        (model(x[a:b]).pow(2).mean() + model.table.abs().mean()).backward()
    dist.destroy_process_group()


def test_direct_sink_ranges_and_no_sync(tmp_path):
    out = str(tmp_path / "g3.pt")
    mp.spawn(_sink_worker, args=(2, 29523, out), nprocs=2, join=True)
    dp = _load_dp()
    model = _SinkModel()
    red = dp.GradAllReducer(model, 1, big_numel=BIG, direct="always")           # world 1: no sinks, no collectives
    model.table._b2n_grad_sink = dp.GradSink(red._views[model.table], lambda *a: None)
    torch.manual_seed(1)
    x = torch.randn(32, 4)
    ref = torch.zeros_like(red.flat)
    for a, b in ((0, 16), (16, 32)):           # the two shards, averaged
        red.zero_grad()
        model.table._b2n_grad_sink.reset()
        (model(x[a:b]).pow(2).mean() + 0.1 * model(x[a:b] * 0.5).abs().mean()).backward()
        ref += red.flat / 2
    assert torch.allclose(torch.load(out), ref, rtol=1e-5, atol=1e-6)
