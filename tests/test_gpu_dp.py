"""Ray-sharded data parallelism on real GPUs: a K-way ray split of one batch through NCCL, on the
real Part-4 model and the real kernels, must reproduce the gradient of the unsplit batch -- including the rank-identical
parameter-only regularisers (run.py:1112-1163), the overlapped table-gradient reduction (GradSink: fine hash levels
all-reduced while the coarse ones are still scattered) and AVG semantics.  Skipped below 2 GPUs."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import GOLDEN  # noqa: F401  (conftest puts the package on sys.path)

pytestmark = pytest.mark.gpu

CFG = dict(mode="part4", scene_bound=1.5, L_embed_dir=4, L_embed_time=10, time_modulation_dim=64, time_modulation_layers=2,
           deform_n_levels=8, deform_n_features_per_level=2, deform_log2_hashmap_size=12, deform_base_resolution=8,
           deform_per_level_scale=1.5, deform_hidden_dim=64, hidden_dim=64, n_levels=16, log2_hashmap_size=15,
           base_resolution=16, per_level_scale=1.4)
B, N = 512, 32


def _build(dev):
    import b2n
    from b2n import synthetic
    from src.core import NeuralField
    from src.renderer import DensityGrid
    b2n.set_mlp_precision("fp32")                       # the 1e-4 class: differences are summation order only
    torch.manual_seed(0)
    model = NeuralField(CFG).to(dev).train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("encoding.params"):
                p.mul_(3000.0)
    grid = DensityGrid(resolution=32, bound=1.5).to(dev)
    grid.binary_grid = synthetic.ball_occupancy(32, 1.5, 1.1).to(dev)
    ro, rd, tgt, t = (v.to(dev) for v in synthetic.random_rays(B, seed=5, n_views=30, with_time=True))
    u = torch.rand(B, N, generator=torch.Generator().manual_seed(6)).to(dev)
    return model, grid, (ro, rd, tgt[:, :3].contiguous(), t, u)


def _loss(model, grid, batch, sl, tv_in_loss=True):
    from src.renderer import render_rays
    ro, rd, tgt, t, u = (v[sl] for v in batch)
    bg = torch.ones(3, device=ro.device)
    c, _, _, ex = render_rays(model, ro, rd, 2.0, 6.0, N, True, density_grid=grid, times=t, bg_color=bg, _jitter=u)
    loss = ((c - tgt) ** 2).mean() + 0.01 * (ex["mean_delta_x"] ** 2).mean()
    # parameter-only regularisers, drawn with a seed every rank shares (run.py:1112-1163): TV on the canonical table and
    # a smoothness term evaluated by calling the canonical encoder directly on random points (a second use of the table)
    table = model.canonical_repr.encoding.params
    if tv_in_loss:
        loss = loss + 1e-3 * torch.mean(torch.abs(table[1:] - table[:-1]))
    pts = (torch.rand(128, 3, generator=torch.Generator().manual_seed(99)) * 2 - 1).to(ro.device)
    loss = loss + 1e-3 * (model.canonical_repr(pts) - model.canonical_repr(pts + 1e-2)).pow(2).mean()
    return loss


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from b2n.dp import GradAllReducer, shard_rays
    model, grid, batch = _build(dev)
    a, b = shard_rays(B, rank, world)
    res = {}
    # big_numel small enough that every hash table takes the overlapped paths.  direct=False: the reference-style loss
    # (TV term on the raw table through autograd), every table reduced from its post-accumulate hook.  direct=True: the
    # canonical table accumulates in place and announces its fine levels early (GradSink); the TV term then belongs
    # into FusedAdamW and a TV term in the loss is refused
    for direct in (False, True):
        red = GradAllReducer(model, world, overlap=True, big_numel=4096, direct=direct)
        assert red.overlap and bool(red._sinks) == direct
        for _ in range(2):                               # two steps: per-step state (use counters, pending list) resets
            red.zero_grad()
            _loss(model, grid, batch, slice(a, b), tv_in_loss=not direct).backward()
            assert red._pending, "nothing was reduced asynchronously"
            if direct:
                assert len(red._pending) >= 2            # the canonical table went out as two level windows
            red.allreduce()
        flat = red.flat.clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree after the all-reduce"
        if direct:
            red.zero_grad()
            with pytest.raises(RuntimeError, match="direct=True"):
                _loss(model, grid, batch, slice(a, b), tv_in_loss=True).backward()
            torch.cuda.synchronize()
        offsets = dict(red._offset)
        red.remove_hooks()
        # the unsplit batch on this GPU, no collective: mean over all rays == AVG of the shard means
        ref = GradAllReducer(model, 1)
        ref.zero_grad()
        _loss(model, grid, batch, slice(0, B), tv_in_loss=not direct).backward()
        worst = 0.0
        for p in ref.params:
            o_new, o_old = ref._offset[p], offsets[p]
            g_ref = ref.flat[o_new:o_new + p.numel()].double()
            g_dp = flat[o_old:o_old + p.numel()].double()
            worst = max(worst, float((g_dp - g_ref).abs().max() / (g_ref.abs().max() + 1e-30)))
        res[f"worst_direct_{direct}"] = worst
        for p in model.parameters():
            p.grad = None
    if rank == 0:
        torch.save(res, out)
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ray_split_through_nccl_equals_unsplit(tmp_path):
    from _util import record
    world = 2
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, 29531, out), nprocs=world, join=True)
    res = torch.load(out)
    assert record("dp_nccl_split_vs_unsplit:grad_max[hook path]", res["worst_direct_False"]) < 1e-4
    assert record("dp_nccl_split_vs_unsplit:grad_max[direct sinks]", res["worst_direct_True"]) < 1e-4
