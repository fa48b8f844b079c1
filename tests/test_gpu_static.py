"""Static-capacity mode (b2n.march.set_static_capacity): the number of active samples behind an occupancy grid stays on
the device (the reference syncs three times per render_rays call, renderer.py:309-323; the exact-size path
of this package once).  Results must equal the exact-size path, and a whole training step must be capturable into a CUDA
graph whose replays equal eager execution."""
import pytest
import torch

from _util import record, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

CFGS = {
    "part2_instant": dict(mode="part2_instant", scene_bound=1.5, n_levels=8, log2_hashmap_size=14, base_resolution=16,
                          per_level_scale=1.5, L_embed_dir=4, hidden_dim=64),
    "part3_instant": dict(mode="part3", canonical_type="instant", scene_bound=1.5, L_embed=6, L_embed_dir=4,
                          L_embed_time=6, hidden_dim=64, deform_hidden_dim=128, deform_num_layers=4, n_levels=8,
                          log2_hashmap_size=14, base_resolution=16, per_level_scale=1.5),
    "part4": dict(mode="part4", scene_bound=1.5, L_embed_dir=4, L_embed_time=10, time_modulation_dim=64,
                  time_modulation_layers=2, deform_n_levels=8, deform_n_features_per_level=2, deform_log2_hashmap_size=12,
                  deform_base_resolution=8, deform_per_level_scale=1.5, deform_hidden_dim=64, hidden_dim=64, n_levels=8,
                  log2_hashmap_size=14, base_resolution=16, per_level_scale=1.5),
}


def _setup(mode, B=900, N=48, R=32):
    import b2n
    from b2n import synthetic
    from src.core import NeuralField
    from src.renderer import DensityGrid
    b2n.set_mlp_precision("bf16")
    torch.manual_seed(0)
    model = NeuralField(CFGS[mode]).to(DEV).train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("encoding.params"):
                p.mul_(3000.0)
        if hasattr(model, "deform_net"):
            model.deform_net.net[-1].weight.mul_(300.0)
    grid = DensityGrid(resolution=R, bound=1.5).to(DEV)
    grid.binary_grid = synthetic.ball_occupancy(R, 1.5, 1.0).to(DEV)
    dyn = mode != "part2_instant"
    batches = []
    for seed in (1, 2, 3):
        ro, rd, tgt, t = (v.to(DEV) for v in synthetic.random_rays(B, seed=seed, n_views=20, with_time=True))
        u = torch.rand(B, N, device=DEV, generator=torch.Generator(DEV).manual_seed(seed))
        batches.append((ro, rd, tgt[:, :3].contiguous(), t if dyn else torch.zeros(B, 1, device=DEV), u))
    return model, grid, batches, dyn, N


def _loss(model, grid, batch, dyn, N):
    from src.renderer import render_rays
    ro, rd, tgt, t, u = batch
    bg = torch.ones(3, device=DEV)
    if dyn:
        c, d, a, ex = render_rays(model, ro, rd, 2.0, 6.0, N, True, density_grid=grid, times=t, bg_color=bg, _jitter=u)
        return ((c - tgt) ** 2).mean() + 0.01 * (ex["mean_delta_x"] ** 2).mean(), (c, d, a)
    c, d, a = render_rays(model, ro, rd, 2.0, 6.0, N, True, density_grid=grid, bg_color=bg, _jitter=u)
    return ((c - tgt) ** 2).mean(), (c, d, a)


@pytest.mark.parametrize("mode", list(CFGS))
def test_static_capacity_equals_exact_and_graph_equals_eager(mode):
    import b2n
    from b2n import march
    model, grid, batches, dyn, N = _setup(mode)
    params = [p for p in model.parameters() if p.requires_grad]
    try:
        # ---- exact-size path (one host read of the count)
        march.set_static_capacity(False)
        ref = []
        for batch in batches:
            loss, outs = _loss(model, grid, batch, dyn, N)
            grads = torch.autograd.grad(loss, params, allow_unused=True)
            ref.append((loss.detach().clone(), [o.detach().clone() for o in outs], grads))
        # ---- static capacity, eager: the same numbers (outputs bit-identical: same kernels on the same rows)
        march.set_static_capacity(True)
        for batch, (l0, o0, g0) in zip(batches, ref):
            loss, outs = _loss(model, grid, batch, dyn, N)
            grads = torch.autograd.grad(loss, params, allow_unused=True)
            assert torch.equal(loss.detach(), l0)
            for a_, b_ in zip(outs, o0):
                assert torch.equal(a_.detach(), b_)
            for p, a_, b_ in zip(params, grads, g0):
                assert (a_ is None) == (b_ is None)
                if a_ is not None:
                    assert torch.isfinite(a_).all()
                    assert record(f"static_vs_exact[{mode}]:grad", rel_err(a_, b_)) < 2e-5      # atomics order only
        # ---- the whole forward + backward as ONE CUDA graph, replayed on other batches
        # (no autograd graph of the eager iterations may stay alive: its AccumulateGrad nodes are bound to the default
        # stream and would be re-used -- and break the capture -- on the capture stream)
        del loss, outs, grads
        for p in params:
            p.grad = torch.zeros_like(p)

        def step(ro, rd, tgt, t, u):
            for p in params:
                p.grad.zero_()
            loss, _ = _loss(model, grid, (ro, rd, tgt, t, u), dyn, N)
            loss.backward()
            return loss

        graphed = b2n.graphs.GraphedStep(step, batches[0])
        for batch, (l0, o0, g0) in zip(batches[::-1], ref[::-1]):
            loss = graphed(*batch)
            torch.cuda.synchronize()
            assert record(f"graph_vs_eager[{mode}]:loss", abs(float(loss) - float(l0)) / abs(float(l0))) < 1e-6
            for p, b_ in zip(params, g0):
                if b_ is not None:
                    assert record(f"graph_vs_eager[{mode}]:grad", rel_err(p.grad, b_)) < 2e-5
        b2n.check_errors()
    finally:
        march.set_static_capacity(False)


def test_static_capacity_empty_grid_and_fp32_refusal():
    """an all-empty grid still evaluates sample 0 of ray 0 (reference renderer.py:309-311); the fp32 layer-by-layer path
    refuses the device-side count instead of computing on unwritten rows"""
    import b2n
    from b2n import march
    model, grid, batches, dyn, N = _setup("part2_instant", B=64)
    grid.binary_grid = torch.zeros_like(grid.binary_grid)
    try:
        march.set_static_capacity(False)
        l0, o0 = _loss(model, grid, batches[0], dyn, N)
        march.set_static_capacity(True)
        l1, o1 = _loss(model, grid, batches[0], dyn, N)
        assert torch.equal(l0, l1) and all(torch.equal(a_, b_) for a_, b_ in zip(o0, o1))
        b2n.set_mlp_precision("fp32")
        with pytest.raises(RuntimeError, match="device-side row count"):
            _loss(model, grid, batches[0], dyn, N)
    finally:
        b2n.set_mlp_precision("fp32")
        march.set_static_capacity(False)
