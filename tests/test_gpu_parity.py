"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI (b2n -> libb2nerf.so),
against (a) the golden vectors generated from the reference and (b) the CPU oracle on seeded
inputs.  Bars (north star): sample depths, occupancy decisions and hash indices BIT-EXACT;
RGB / depth / acc / gradients within 1e-4 relative (max|a-b| / max|b|) in fp32."""
import glob
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, full_nerf_state_dict, load, record, rel_err, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda"


@pytest.fixture(scope="module")
def mods():
    import b2n
    from b2n import march
    from src import renderer
    from src.core import NeuralField
    b2n.set_mlp_precision("fp32")          # the 1e-4 bar is for the fp32 path; bf16 tests switch explicitly
    return dict(b2n=b2n, march=march, renderer=renderer, NeuralField=NeuralField)


@pytest.fixture
def bf16_mode(mods):
    mods["b2n"].set_mlp_precision("bf16")
    yield
    mods["b2n"].set_mlp_precision("fp32")


def cu(t):
    return t.to(DEV) if isinstance(t, torch.Tensor) else t


# ------------------------------------------------------------------ sampling / occupancy (bit exact)
@pytest.mark.parametrize("N", [64, 128, 2, 1])
def test_sampling_bit_exact(mods, N):
    g = load(f"sampling_N{N}")
    B = g["u"].shape[0]
    dummy = torch.zeros(B, 3, device=DEV)
    m = mods["march"].march(dummy, dummy, g["near"], g["far"], N, cu(g["u"]))
    assert torch.equal(m.z.cpu(), g["z_pert"])
    m = mods["march"].march(dummy, dummy, g["near"], g["far"], N, None)
    assert torch.equal(m.z.cpu(), g["z_flat"])


@pytest.mark.parametrize("R", [4, 16, 128])
def test_active_mask_bit_exact(mods, R):
    g = load(f"mask_R{R}")
    grid = mods["renderer"].DensityGrid(resolution=R, bound=g["bound"], threshold=0.01).to(DEV)
    grid.binary_grid = cu(g["binary_grid"])
    assert torch.equal(grid.get_active_mask(cu(g["pts"])).cpu(), g["mask"])
    # in-place edits of the bool buffer must be picked up (bitfield cache keyed on _version)
    grid.binary_grid.fill_(False)
    assert not grid.get_active_mask(cu(g["pts"])).any()


def test_march_mask_and_compaction_vs_oracle(mods):
    from oracle import nerf_oracle as O
    torch.manual_seed(0)
    B, N, R, bound = 777, 128, 128, 1.5
    ro, rd, _ = O.synthetic_rays(B, seed=3)
    u = torch.rand(B, N)
    occ = O.ball_occupancy(R, bound, 0.75)
    z = O.sample_stratified(2.0, 6.0, N, B, u)
    pts = (ro[:, None] + rd[:, None] * z[..., None]).reshape(-1, 3)
    mask = O.active_mask(pts, occ, bound)
    bits = mods["march"].pack_occupancy(cu(occ))
    m = mods["march"].march(cu(ro), cu(rd), 2.0, 6.0, N, cu(u), bits=bits, R=R, bound=bound, want_idx=True)
    assert torch.equal(m.z.cpu(), z)
    assert m.n_active == int(mask.sum())
    assert torch.equal(m.sample_idx.cpu().long(), mask.nonzero().squeeze(-1))          # order preserving
    assert torch.equal(m.pts.cpu(), pts[mask])                                           # bit-exact points
    vd = (rd / rd.norm(dim=-1, keepdim=True))[:, None].expand(-1, N, -1).reshape(-1, 3)[mask]
    assert rel_err(m.dirs.cpu(), vd) < 1e-6
    offs = torch.cat([torch.zeros(1, dtype=torch.long), mask.view(B, N).sum(1).cumsum(0)])
    assert torch.equal(m.ray_offset.cpu().long(), offs)


def test_march_empty_grid_forces_first_sample(mods):
    """reference renderer.py:309-311: an all-empty mask still queries sample 0 of ray 0."""
    B, N, R = 50, 64, 16
    occ = torch.zeros(R, R, R, dtype=torch.bool, device=DEV)
    bits = mods["march"].pack_occupancy(occ)
    ro = torch.tensor([[0.0, 0.0, 4.0]], device=DEV).repeat(B, 1)
    rd = torch.tensor([[0.0, 0.0, -1.0]], device=DEV).repeat(B, 1)
    m = mods["march"].march(ro, rd, 2.0, 6.0, N, None, bits=bits, R=R, bound=1.5, want_idx=True)
    assert m.n_active == 1 and m.sample_idx.tolist() == [0]
    assert m.ray_offset.tolist() == [0] + [1] * B
    assert torch.equal(m.pts.cpu(), torch.tensor([[0.0, 0.0, 2.0]]))


def test_march_ragged_sizes(mods):
    for B, N in ((0, 64), (1, 1), (3, 33), (2049, 31), (5000, 64)):
        dummy = torch.zeros(B, 3, device=DEV)
        m = mods["march"].march(dummy, dummy + 1, 2.0, 6.0, N, None)
        assert m.z.shape == (B, N) and m.pts.shape == (B * N, 3)


# ------------------------------------------------------------------ compositing
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "composite_*.npz"))))
def test_composite_golden(mods, path):
    g = load(os.path.basename(path)[:-4])
    rgb, sigma = cu(g["rgb"]).requires_grad_(True), cu(g["sigma"]).requires_grad_(True)
    bg = cu(g["bg"]) if g["bg"].numel() else None
    c, d, a = mods["renderer"].volume_render(rgb, sigma, cu(g["z"]), cu(g["rays_d"]), bg_color=bg)
    assert rel_err(c.cpu(), g["color"]) < TOL and rel_err(d.cpu(), g["depth"]) < TOL and rel_err(a.cpu(), g["acc"]) < TOL
    loss = (c * cu(g["g_color"])).sum() + (d * cu(g["g_depth"])).sum() + (a * cu(g["g_acc"])).sum()
    gr, gs = torch.autograd.grad(loss, [rgb, sigma])
    assert rel_err(gr.cpu(), g["g_rgb"]) < TOL
    assert rel_err(gs.cpu(), g["g_sigma"]) < TOL


@pytest.mark.parametrize("N", [96, 64, 33, 128, 200, 7])
def test_composite_compact_with_dx_vs_oracle(mods, N):
    """N <= 128 runs the register-resident kernels, longer rays the chunk-loop kernels"""
    from oracle import nerf_oracle as O
    torch.manual_seed(1)
    B = 300
    mask = torch.rand(B, N) < 0.3
    mask[5] = False
    mask[6] = True
    z = torch.sort(torch.rand(B, N) * 4 + 2, dim=-1)[0]
    rd = torch.randn(B, 3)
    bg = torch.rand(B, 3)
    Pn = int(mask.sum())
    rgb_c = torch.rand(Pn, 3, dtype=torch.float64, requires_grad=True)
    sig_c = (torch.rand(Pn, dtype=torch.float64) * 30).requires_grad_(True)
    dx_c = torch.randn(Pn, 3, dtype=torch.float64, requires_grad=True)
    flat = mask.reshape(-1)
    rgb = torch.zeros(B * N, 3, dtype=torch.float64).index_put((flat,), rgb_c).view(B, N, 3)
    sig = torch.zeros(B * N, dtype=torch.float64).index_put((flat,), sig_c).view(B, N)
    dxd = torch.zeros(B * N, 3, dtype=torch.float64).index_put((flat,), dx_c).view(B, N, 3)
    c, d, a = O.volume_render(rgb, sig, z.double(), rd.double(), bg.double())
    w = O.composite_weights(sig, z.double(), rd.double())
    mdx = (w[..., None] * dxd).sum(1)
    gc, gd, ga, gm = torch.randn(B, 3), torch.randn(B), torch.randn(B), torch.randn(B, 3)
    loss = (c * gc).sum() + (d * gd).sum() + (a * ga).sum() + (mdx * gm).sum()
    ref = torch.autograd.grad(loss, [rgb_c, sig_c, dx_c])
    # CUDA: compact layout through mask words / offsets
    W = (N + 31) // 32
    bitsarr = np.zeros((B, W * 32), dtype=bool)
    bitsarr[:, :N] = mask.numpy()
    words = np.packbits(bitsarr.reshape(B, W, 32), axis=-1, bitorder="little").view(np.uint32).reshape(B, W)
    words = torch.from_numpy(words.view(np.int32).copy()).to(DEV)
    offs = torch.cat([torch.zeros(1, dtype=torch.long), mask.sum(1).cumsum(0)]).int().to(DEV)
    r2, s2, x2 = (t.detach().float().to(DEV).requires_grad_(True) for t in (rgb_c, sig_c, dx_c))
    c2, d2, a2, m2 = mods["b2n"].composite(r2, s2, cu(z), cu(rd), bg=cu(bg), dx=x2, mask_words=words, ray_offset=offs)
    for got, want in ((c2, c), (d2, d), (a2, a), (m2, mdx)):
        assert rel_err(got.cpu(), want) < TOL
    loss2 = (c2 * cu(gc)).sum() + (d2 * cu(gd)).sum() + (a2 * cu(ga)).sum() + (m2 * cu(gm)).sum()
    got = torch.autograd.grad(loss2, [r2, s2, x2])
    for gg, rr in zip(got, ref):
        assert rel_err(gg.cpu(), rr) < TOL


def test_composite_properties_full_size(mods):
    """C2-sized batch (2^18 rays x 128 samples): weights sum to <= 1, linear in rgb, and the
    dense and compact layouts agree when every sample is active."""
    torch.manual_seed(2)
    B, N = 2 ** 18, 128
    sigma = torch.rand(B, N, device=DEV) * 2
    rgb = torch.rand(B, N, 3, device=DEV)
    z = torch.linspace(2, 6, N, device=DEV).expand(B, N).contiguous()
    rd = torch.randn(B, 3, device=DEV)
    c1, d1, a1, _ = mods["b2n"].composite(rgb, sigma, z, rd)
    assert float(a1.max()) <= 1.0 + 1e-5 and float(a1.min()) >= 0.0
    assert float(d1.min()) >= 0.0 and float(d1.max()) <= 6.0 + 1e-4
    c2, _, _, _ = mods["b2n"].composite(rgb * 0.5, sigma, z, rd)
    assert rel_err(c2, c1 * 0.5) < 1e-6
    words = torch.full((B, 4), -1, dtype=torch.int32, device=DEV)
    offs = (torch.arange(B + 1, device=DEV) * N).int()
    c3, d3, a3, _ = mods["b2n"].composite(rgb, sigma, z, rd, mask_words=words, ray_offset=offs)
    assert torch.equal(c3, c1) and torch.equal(d3, d1) and torch.equal(a3, a1)


# ------------------------------------------------------------------ encoders
@pytest.mark.parametrize("D,L", [(3, 10), (3, 4), (1, 10), (1, 6), (3, 0)])
def test_fourier_golden(mods, D, L):
    g = load(f"pe_D{D}_L{L}")
    x = cu(g["x"]).requires_grad_(True)
    y = mods["b2n"].fourier_encode(x, cu(g["bands"]))
    # arguments are bit-identical to the reference's; sincosf vs the host libm differ by <= 2 ulp
    assert (y.cpu() - g["y"]).abs().max() < 1e-6
    if L:
        gx, = torch.autograd.grad((y * cu(g["g_y"])).sum(), x)
        assert rel_err(gx.cpu(), g["g_x"]) < TOL


@pytest.mark.parametrize("tag", ["c2", "c5canon", "c5deform", "small"])
def test_hash_encode_vs_oracle(mods, tag):
    from oracle import nerf_oracle as O
    g = load(f"hash_kat_{tag}")
    c = g["cfg"]
    geom = mods["b2n"].HashGeometry(c["n_levels"], c["base_resolution"], c["per_level_scale"],
                                    c["log2_hashmap_size"], c["n_features_per_level"])
    lv = O.hash_level_table(c["n_levels"], c["base_resolution"], c["per_level_scale"], c["log2_hashmap_size"])
    assert [(l.scale, l.res, l.size, l.offset, l.hashed) for l in lv] == geom.levels      # host tables identical
    torch.manual_seed(4)
    table = torch.randn(geom.n_params) * 0.5
    bound = 1.5
    xw = torch.cat([(g["x"] * 2 - 1) * bound, (torch.rand(3000, 3) * 2 - 1) * bound * 1.1])   # some points outside
    # integer KAT: a one-hot table per corner is overkill; instead compare features + both gradients
    t_ref, x_ref = table.clone().requires_grad_(True), xw.clone().requires_grad_(True)
    y_ref = O.hash_representation(x_ref, t_ref, lv, c["n_features_per_level"], bound)
    gy = torch.randn_like(y_ref)
    gt_ref, gx_ref = torch.autograd.grad((y_ref * gy).sum(), [t_ref, x_ref])
    t_cu, x_cu = cu(table).requires_grad_(True), cu(xw).requires_grad_(True)
    y = mods["b2n"].hash_encode(x_cu, t_cu, geom, bound)
    assert record(f"hash_{tag}:features", rel_err(y.cpu(), y_ref)) < 1e-5
    gt, gx = torch.autograd.grad((y * cu(gy)).sum(), [t_cu, x_cu])
    assert record(f"hash_{tag}:g_table", rel_err(gt.cpu(), gt_ref)) < TOL
    assert record(f"hash_{tag}:g_x", rel_err(gx.cpu(), gx_ref)) < TOL
    # which table entries receive gradient is an INDEX decision: must match exactly
    assert torch.equal(gt.cpu() != 0, gt_ref != 0)


def test_hash_index_kat_exact(mods):
    """Integer KAT: with table[e] = e (entry id, exactly representable < 2^24) and a point sitting
    on a lattice node, the feature equals the entry index of that node."""
    g = load("hash_kat_c2")
    c = g["cfg"]
    geom = mods["b2n"].HashGeometry(c["n_levels"], c["base_resolution"], c["per_level_scale"],
                                    c["log2_hashmap_size"], 1)
    table = torch.arange(geom.n_entries, dtype=torch.float32, device=DEV)
    ent = g["corner_entries"]                               # [L, 64, 8] (F-independent entry ids)
    x = cu(g["x"])
    y = mods["b2n"].hash_encode(x, table, geom, 0.0).cpu()   # unit-cube input
    # trilinear blend of the 8 corner ids with the oracle's weights must reproduce y
    for li, (scale, res, size, offset, hashed) in enumerate(geom.levels):
        pos = g["x"] * torch.tensor(scale) + 0.5
        w = pos - torch.floor(pos)
        acc = torch.zeros(64, dtype=torch.float64)
        for k in range(8):
            wt = torch.ones(64, dtype=torch.float64)
            for d in range(3):
                wd = w[:, d].double()
                wt = wt * (wd if (k >> d) & 1 else 1 - wd)
            acc += wt * ent[li, :, k].double()
        assert ((y[:, li].double() - acc).abs() / (acc.abs() + 1)).max() < 1e-5, li


@pytest.mark.parametrize("variant,merge_res", [(3, 64), (3, 0), (3, 200), (0, 64), (1, 64), (2, 24)])
@pytest.mark.parametrize("log2T", [19, 20])
def test_hash_table_grad_on_ray_ordered_samples(mods, variant, merge_res, log2T):
    """Consecutive samples of a ray share coarse cells: the table-gradient kernels merge such runs inside a warp before
    they issue red.global.  Points marched along rays (runs of 1..10 lanes, ragged total, rays ending mid-warp), every
    kernel variant / merge threshold, both canonical geometries (T = 2^20 makes level 4 a dense 81^3 level)."""
    from oracle import nerf_oracle as O
    b2n = mods["b2n"]
    lib = b2n._lib.lib
    geom = b2n.HashGeometry(16, 16, 1.5, log2T, 2)
    lv = O.hash_level_table(16, 16, 1.5, log2T)
    gen = torch.Generator().manual_seed(9)
    o = (torch.rand(37, 1, 3, generator=gen) * 2 - 1) * 1.2
    dirs = torch.nn.functional.normalize(torch.randn(37, 1, 3, generator=gen), dim=-1)
    tt = torch.arange(97).view(1, -1, 1) * 0.013
    xw = (o + dirs * tt).reshape(-1, 3)[: 37 * 97 - 5].contiguous()
    table = torch.randn(geom.n_params, generator=gen) * 0.5
    t_ref, x_ref = table.clone().requires_grad_(True), xw.clone().requires_grad_(True)
    y_ref = O.hash_representation(x_ref, t_ref, lv, 2, 1.5)
    gy = torch.randn(y_ref.shape, generator=gen)
    gy[::11] = 0.0                                  # samples that receive no gradient
    gt_ref, gx_ref = torch.autograd.grad((y_ref * gy).sum(), [t_ref, x_ref])
    try:
        lib.b2n_debug_hash_variant(variant, merge_res)
        t_cu, x_cu = cu(table).requires_grad_(True), cu(xw).requires_grad_(True)
        y = b2n.hash_encode(x_cu, t_cu, geom, 1.5)
        gt, gx = torch.autograd.grad((y * cu(gy)).sum(), [t_cu, x_cu])
    finally:
        lib.b2n_debug_hash_variant(3, 64)
    assert rel_err(y.cpu(), y_ref) < 1e-5
    assert record(f"hash_rays[v{variant},m{merge_res},T{log2T}]:g_table", rel_err(gt.cpu(), gt_ref)) < TOL
    assert rel_err(gx.cpu(), gx_ref) < TOL
    assert torch.equal(gt.cpu() != 0, gt_ref != 0)


def test_linear_vs_torch(mods):
    torch.manual_seed(5)
    for (Pn, K, N, act) in ((1000, 63, 256, "relu"), (517, 319, 256, "relu"), (300, 256, 1, "relu"),
                            (129, 283, 128, "none"), (64, 128, 3, "sigmoid"), (5, 21, 64, "relu"), (0, 8, 8, "none")):
        x = torch.randn(Pn, K, requires_grad=True)
        W = (torch.randn(N, K) / K ** 0.5).requires_grad_(True)
        b = torch.randn(N, requires_grad=True)
        y = torch.nn.functional.linear(x, W, b)
        y = {"relu": torch.relu, "sigmoid": torch.sigmoid, "none": lambda v: v}[act](y)
        gy = torch.randn_like(y)
        x2, W2, b2 = (cu(t.detach()).requires_grad_(True) for t in (x, W, b))
        y2 = mods["b2n"].linear(x2, W2, b2, act)
        assert rel_err(y2.cpu(), y) < 1e-5
        if Pn:
            ref = torch.autograd.grad((y * gy).sum(), [x, W, b])
            got = torch.autograd.grad((y2 * cu(gy)).sum(), [x2, W2, b2])
            for a_, b_ in zip(got, ref):
                assert rel_err(a_.cpu(), b_) < 2e-5


# ------------------------------------------------------------------ fields and render_rays vs golden
FIELDS = ["part2_nerf", "part2_instant", "part3_nerf", "part3_dtc", "part3_instant", "part4"]


def _model_from(mods, cfg, sd):
    model = mods["NeuralField"](cfg)
    model.load_state_dict(sd)
    return model.to(DEV)


def _fp64_truth(g, pert):
    """parameter gradients of a golden render case evaluated by the CPU oracle in fp64 (the exact result, to ~1e-12)"""
    from oracle import nerf_oracle as O
    dt = torch.float64
    sd0 = g["sd"] if g["sd"] else full_nerf_state_dict(int(g["seed"]))
    sd = {k: (v.to(dt) if v.is_floating_point() else v).clone().requires_grad_(v.is_floating_point() and "freq" not in k)
          for k, v in sd0.items()}
    if "deformation_grid.encoding.params" in sd:
        sd["deformation_grid.encoding.params"] = sd["deform_grid_start.encoding.params"]
    field = O.OracleField(g["cfg"], sd).train(False)
    out = O.render_rays(field, g["rays_o"].to(dt), g["rays_d"].to(dt), g["near"], g["far"], int(g["n_samples"]),
                        g["u"].to(dt) if pert else None, binary_grid=g.get("binary_grid"),
                        grid_bound=g.get("grid_bound", 1.0), bg_color=g["bg"].to(dt),
                        times=g["times"].to(dt) if "times" in g else None)
    loss = (out[0] * g["g_color"].to(dt)).sum()
    if "mean_delta_x" in g:
        loss = loss + (out[3]["mean_delta_x"] * g["g_mdx"].to(dt)).sum()
    names = [n for n in sd if sd[n].requires_grad and n != "deformation_grid.encoding.params"]
    grads = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
    return {n: gr for n, gr in zip(names, grads) if gr is not None}


def _sums(t):
    t = t.double()
    return torch.stack([t.sum(), t.abs().sum(), (t ** 2).sum()])


def _sum_err(s, ref):
    # sum(g) cancels; it is judged against sum|g| (the scale of its terms), the two norms relatively
    return max(float((s[0] - ref[0]).abs() / (ref[1] + 1e-30)), float((s[1] - ref[1]).abs() / (ref[1] + 1e-30)),
               float((s[2] - ref[2]).abs() / (ref[2] + 1e-30)))


def _check_grads(model, loss, g, tol, l2=False, tag="", truth=None):
    """Max-norm relative error per parameter tensor against the reference's autograd gradients (fp32 bar: 1e-4).
    l2=True: relative Frobenius error -- the metric for the 16-bit tensor-core decoders, whose per-point gradients
    differ from fp32 by isolated ReLU-mask flips (unbiased, not small in max-norm).

    ``truth`` (fp32 path only): a callable returning the fp64 gradients of the same case.  The reference's own fp32
    gradients sit 0.4 - 2.2e-4 from the exact ones on the sigma-head weights of the 256-wide decoder (long sums with
    cancellation: measured on the fixtures), so two fp32 evaluations with different summation
    orders cannot be asked to agree to 1e-4 there.  A tensor that misses ``tol`` against the fixture passes only if it
    is within ``tol`` of the EXACT gradient, or in the same error class as the reference's own fp32 result
    (<= 2 x the reference's distance to exact + tol / 2: two fp32 evaluations land on either side of the truth)."""
    names = list(g["grads"]) + list(g["gradsum"])
    params = dict(model.named_parameters())
    grads = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
    worst = 0.0
    for n, gr in zip(names, grads):
        gr = torch.zeros_like(params[n]) if gr is None else gr
        gr = gr.cpu()
        if l2:
            if n in g["grads"]:
                e = record(f"{tag}:grad_l2:{n}", rel_l2(gr, g["grads"][n]))
                worst = max(worst, e)
                assert e < tol, (n, e)
            continue
        if n in g["grads"]:
            e = record(f"{tag}:grad_max:{n}", rel_err(gr, g["grads"][n]))
            if e >= tol and truth is not None:
                exact = truth()[n]
                e_ours = record(f"{tag}:grad_max_vs_fp64:{n}", rel_err(gr, exact))
                e_ref = record(f"{tag}:reference_fp32_vs_fp64:{n}", rel_err(g["grads"][n], exact))
                assert e_ours < tol or e_ours <= 2 * e_ref + tol / 2, (n, e, e_ours, e_ref)
            else:
                assert e < tol, (n, e)
        else:
            ref = g["gradsum"][n].double()
            e = record(f"{tag}:gradsum:{n}", _sum_err(_sums(gr), ref))
            if e >= tol and truth is not None:
                exact = _sums(truth()[n])
                e_ours = record(f"{tag}:gradsum_vs_fp64:{n}", _sum_err(_sums(gr), exact))
                e_ref = record(f"{tag}:reference_fp32_vs_fp64:{n}", _sum_err(ref, exact))
                assert e_ours < tol or e_ours <= 2 * e_ref + tol / 2, (n, e, e_ours, e_ref)
            else:
                assert e < tol, (n, e)
            if n in g["gradhead"]:
                e = record(f"{tag}:gradhead:{n}", rel_err(gr.reshape(-1)[:4096], g["gradhead"][n]))
                assert e < tol, (n, e)
    return worst


@pytest.mark.parametrize("tag", FIELDS)
def test_field_golden(mods, tag):
    g = load(f"field_{tag}")
    model = _model_from(mods, g["cfg"], g["sd"]).eval()
    dyn = g["cfg"]["mode"] in ("part3", "part4")
    out = model(cu(g["x"]), cu(g["d"]), t=cu(g["t"])) if dyn else model(cu(g["x"]), cu(g["d"]))
    assert record(f"field_{tag}:rgb", rel_err(out[0].cpu(), g["rgb"])) < TOL
    assert record(f"field_{tag}:sigma", rel_err(out[1].cpu(), g["sigma"])) < TOL
    loss = (out[0] * cu(g["g_rgb"])).sum() + (out[1] * cu(g["g_sigma"])).sum()
    if dyn:
        assert rel_err(out[2].cpu(), g["dx"]) < TOL
        loss = loss + (out[2] * cu(g["g_dx"])).sum()
    _check_grads(model, loss, g, TOL, tag=f"field_{tag}")


@pytest.mark.parametrize("tag", FIELDS + ["part2_nerf_full"])
@pytest.mark.parametrize("pert", ["flat", "pert"])
def test_render_rays_golden(mods, tag, pert):
    g = load(f"render_{tag}_{pert}")
    sd = g["sd"] if g["sd"] else full_nerf_state_dict(int(g["seed"]))
    model = _model_from(mods, g["cfg"], sd).train(pert == "pert")
    grid = None
    if "binary_grid" in g:
        grid = mods["renderer"].DensityGrid(resolution=g["binary_grid"].shape[0], bound=g["grid_bound"]).to(DEV)
        grid.binary_grid = cu(g["binary_grid"])
    times = cu(g["times"]) if "times" in g else None
    out = mods["renderer"].render_rays(model, cu(g["rays_o"]), cu(g["rays_d"]), g["near"], g["far"],
                                       int(g["n_samples"]), pert == "pert", density_grid=grid, times=times,
                                       bg_color=cu(g["bg"]), _jitter=cu(g["u"]) if pert == "pert" else None)
    assert len(out) == (4 if times is not None else 3)
    assert record(f"render_{tag}_{pert}:color", rel_err(out[0].cpu(), g["color"])) < TOL
    assert record(f"render_{tag}_{pert}:depth", rel_err(out[1].cpu(), g["depth"])) < TOL
    assert record(f"render_{tag}_{pert}:acc", rel_err(out[2].cpu(), g["acc"])) < TOL
    loss = (out[0] * cu(g["g_color"])).sum()
    if "mean_delta_x" in g:
        assert rel_err(out[3]["mean_delta_x"].cpu(), g["mean_delta_x"]) < TOL
        loss = loss + (out[3]["mean_delta_x"] * cu(g["g_mdx"])).sum()
    cache = {}

    def truth():
        if not cache:
            cache.update(_fp64_truth(g, pert == "pert"))
        return cache
    _check_grads(model, loss, g, TOL, tag=f"render_{tag}_{pert}", truth=truth)


@pytest.mark.parametrize("tag", ["part2_instant", "part3_instant", "part4"])
def test_density_grid_update_golden(mods, tag):
    g = load(f"gridupdate_{tag}")
    model = _model_from(mods, g["cfg"], g["sd"]).eval()
    R = g["grid1"].shape[0]
    grid = mods["renderer"].DensityGrid(resolution=R, bound=g["cfg"]["scene_bound"], threshold=g["threshold"]).to(DEV)
    part3 = g["cfg"]["mode"] == "part3"
    r1 = grid.update(model, device=DEV, time=torch.tensor([[0.3]], device=DEV) if part3 else None)
    assert rel_err(grid.grid.cpu(), g["grid1"]) < TOL
    stable = (g["grid1"] - g["threshold"]).abs() > 1e-5 * max(1.0, g["threshold"])
    assert torch.equal(grid.binary_grid.cpu()[stable], g["binary1"][stable])
    assert abs(r1 - g["ratio1"]) < 3e-3
    # run.py:1982-1985 passes extra keywords the reference signature lacks: must be accepted
    r2 = grid.update(model, device=DEV, time=torch.tensor([[0.8]], device=DEV) if part3 else None, decay=0.95,
                     auto_prune=True, threshold_multiplier=1.0)
    assert rel_err(grid.grid.cpu(), g["grid2"]) < TOL
    stable = (g["grid2"] - g["threshold"]).abs() > 1e-5 * max(1.0, g["threshold"])
    assert torch.equal(grid.binary_grid.cpu()[stable], g["binary2"][stable])
    assert abs(r2 - g["ratio2"]) < 3e-3
    # the refreshed bitfield is what get_active_mask reads
    pts = (torch.rand(5000, 3, device=DEV) * 2 - 1) * g["cfg"]["scene_bound"]
    from oracle import nerf_oracle as O
    assert torch.equal(grid.get_active_mask(pts).cpu(), O.active_mask(pts.cpu(), grid.binary_grid.cpu(), grid.bound))


@pytest.mark.parametrize("tag", FIELDS)
def test_density_branch_equals_forward_sigma(mods, tag):
    """NeuralField.density (the sigma-only sweep of DensityGrid.update) returns exactly the sigma that
    forward(x, 0, t) returns, in both precision modes (same kernels, colour branch skipped)."""
    g = load(f"field_{tag}")
    model = _model_from(mods, g["cfg"], g["sd"]).eval()
    torch.manual_seed(3)
    n = 5000
    x = (torch.rand(n, 3, device=DEV) * 2 - 1) * float(g["cfg"].get("scene_bound", 1.0))
    t = torch.rand(n, 1, device=DEV)
    dyn = model.mode in ("part3", "part4")
    with torch.no_grad():
        full = model(x, torch.zeros_like(x), t=t)[1] if dyn else model(x, torch.zeros_like(x))[1]
        dens = model.density(x, t=t) if dyn else model.density(x)
        assert dens.shape == full.shape and torch.equal(dens, full)
        mods["b2n"].set_mlp_precision("bf16")
        try:
            full_b = model(x, torch.zeros_like(x), t=t)[1] if dyn else model(x, torch.zeros_like(x))[1]
            dens_b = model.density(x, t=t) if dyn else model.density(x)
        finally:
            mods["b2n"].set_mlp_precision("fp32")
    # 16-bit mode: hash-grid decoders take sigma from the SAME fused kernel with its colour network switched off
    # (b2n.instant_sigma) -- bit-identical; the 256-wide decoder runs forward and drops rgb -- identical by construction
    assert torch.equal(dens_b, full_b) and rel_err(dens_b.cpu(), full.cpu()) < 2e-2


def test_density_grid_update_uses_density_branch(mods):
    """update() through model.density and through the reference-style full forward (a model without .density) give the
    same grids; the density path launches fewer kernels per sweep."""
    g = load("gridupdate_part2_instant")
    model = _model_from(mods, g["cfg"], g["sd"]).eval()
    R = g["grid1"].shape[0]

    class NoDensity(torch.nn.Module):          # what a foreign model looks like to DensityGrid.update
        def __init__(self, m):
            super().__init__()
            self.m, self.mode = m, m.mode

        def forward(self, *a, **k):
            return self.m(*a, **k)

    grids = []
    for mdl in (model, NoDensity(model)):
        grid = mods["renderer"].DensityGrid(resolution=R, bound=g["cfg"]["scene_bound"], threshold=g["threshold"]).to(DEV)
        n0 = mods["b2n"]._lib.LAUNCHES["count"]
        ratio = grid.update(mdl, device=DEV)
        grids.append((grid.grid.clone(), grid.binary_grid.clone(), ratio, mods["b2n"]._lib.LAUNCHES["count"] - n0))
    assert torch.equal(grids[0][0], grids[1][0]) and torch.equal(grids[0][1], grids[1][1]) and grids[0][2] == grids[1][2]
    assert grids[0][3] < grids[1][3]


def test_state_dict_keys_match_reference(mods):
    for tag in FIELDS:
        g = load(f"field_{tag}")
        model = mods["NeuralField"](g["cfg"])
        assert set(model.state_dict()) == set(g["sd"])


def test_amp_autocast_compatible(mods):
    """run.py:1092 wraps render_rays in torch.amp.autocast('cuda'): ops must keep fp32 semantics."""
    g = load("render_part4_flat")
    model = _model_from(mods, g["cfg"], g["sd"]).eval()
    grid = mods["renderer"].DensityGrid(resolution=g["binary_grid"].shape[0], bound=g["grid_bound"]).to(DEV)
    grid.binary_grid = cu(g["binary_grid"])
    with torch.amp.autocast("cuda"):
        out = mods["renderer"].render_rays(model, cu(g["rays_o"]), cu(g["rays_d"]), g["near"], g["far"],
                                           int(g["n_samples"]), False, density_grid=grid, times=cu(g["times"]),
                                           bg_color=cu(g["bg"]))
    assert out[0].dtype == torch.float32
    assert rel_err(out[0].cpu(), g["color"]) < 2e-3      # autocast runs the torch glue (cat, blend) in fp16


# ------------------------------------------------------------------ bf16 tensor-core decoder (1e-2 class)
BF16_TOL = 1e-2


@pytest.mark.parametrize("pos_dim,Pn", [(32, 1000), (32, 64 * 37), (53, 777), (32, 5), (64, 130), (32, 40000)])
@pytest.mark.parametrize("gscale", [1.0, 65536.0, 2.0 ** -22])
def test_instant_mlp_fp16_vs_oracle(mods, pos_dim, Pn, gscale):
    """Fused Instant decoder (fp16 operands, split first layer, power-of-two scaled gradient chain) against
    (a) the oracle evaluated in the kernel's arithmetic model -- tight, forward and backward -- and
    (b) the plain fp32 oracle -- the north star's 1e-2 bar for the 16-bit MLP, outputs and gradients.
    gscale multiplies the incoming gradient (GradScaler-sized and underflow-sized): the result must scale with it."""
    from oracle import nerf_oracle as O
    torch.manual_seed(7)
    gen = torch.Generator().manual_seed(7)
    sd = {"d.sigma_net.params": O._fused_init(pos_dim, 16, 64, 1, gen).requires_grad_(True),
          "d.color_net.params": O._fused_init(16 + 27, 3, 64, 2, gen).requires_grad_(True)}
    x = (torch.randn(Pn, pos_dim) * 0.5).requires_grad_(True)
    d = torch.randn(Pn, 3)
    d = d / d.norm(dim=-1, keepdim=True)
    bands = O.fourier_bands(4)
    wrt = [x, sd["d.sigma_net.params"], sd["d.color_net.params"]]
    rgb, sigma = O.instant_decoder(sd, "d", x, O.fourier_encode(d, bands), 64)
    rgb_q, sigma_q = O.instant_decoder(sd, "d", x, O.fourier_encode(d, bands), 64, emulate_bf16="kernel")
    g_rgb, g_sigma = torch.randn_like(rgb) * gscale, torch.randn_like(sigma) * gscale
    ref32 = torch.autograd.grad((rgb * g_rgb).sum() + (sigma * g_sigma).sum(), wrt)
    ref_q = torch.autograd.grad((rgb_q * g_rgb).sum() + (sigma_q * g_sigma).sum(), wrt)
    x2 = cu(x.detach()).requires_grad_(True)
    sp, cp = (cu(sd[k].detach()).requires_grad_(True) for k in ("d.sigma_net.params", "d.color_net.params"))
    rgb2, sigma2 = mods["b2n"].instant_mlp(x2, cu(d), cu(bands), sp, cp)
    assert rgb2.shape == (Pn, 3) and sigma2.shape == (Pn, 1)
    tag = f"instant_mlp[{pos_dim},{Pn}]"
    assert record(f"{tag}:rgb_vs_fp32", rel_err(rgb2.cpu(), rgb)) < BF16_TOL
    assert record(f"{tag}:sigma_vs_fp32", rel_err(sigma2.cpu(), sigma)) < BF16_TOL
    assert record(f"{tag}:rgb_vs_kernel_model", rel_err(rgb2.cpu(), rgb_q)) < 1e-3
    assert record(f"{tag}:sigma_vs_kernel_model", rel_err(sigma2.cpu(), sigma_q)) < 1e-3
    got = torch.autograd.grad((rgb2 * cu(g_rgb)).sum() + (sigma2 * cu(g_sigma)).sum(), [x2, sp, cp])
    for a_, bq, b32, name in zip(got, ref_q, ref32, ("g_x", "g_sigma_params", "g_color_params")):
        a_ = a_.cpu()
        assert torch.isfinite(a_).all()
        # kernel vs the oracle in the kernel's arithmetic: what is left is accumulation order and rounding ties
        assert record(f"{tag}:{name}_l2_vs_kernel_model", rel_l2(a_, bq)) < 5e-3, name
        # kernel vs fp32.  This case is adversarial on purpose: i.i.d. N(0, 0.5) features, Xavier weights and, above
        # all, i.i.d. N(0, 1) OUTPUT gradients -- the parameter gradient is then a random walk over the points and the
        # ReLU-mask flips of individual points (inherent to any 11-bit forward; the arithmetic model above gives the
        # same figures) do not average out: 1.3 - 2.9e-2 whatever P.  With the coherent gradients of a rendering loss
        # they do: the golden render tests hold the same quantities to the north star's 1e-2.
        assert record(f"{tag}:{name}_l2_vs_fp32", rel_l2(a_, b32)) < 5e-2, name
    # padded rows/columns of the flat parameter vectors never receive gradient
    V3 = got[2][64 * 48 + 64 * 64:].view(16, 64)
    assert float(V3[3:].abs().max()) == 0.0


@pytest.mark.parametrize("pos_dim,Pn,pad", [(32, 128 * 11 + 3, 0.0), (53, 777, 0.0), (64, 130, 0.0), (32, 5, 0.0),
                                            (20, 1000, 1.0), (32, 300000, 0.0)])
def test_instant_fwd_tcgen05_matches_mma_sync(mods, pos_dim, Pn, pad):
    """The two forward kernels of the fused Instant decoder (tcgen05: b2n_mlp64tc.cu, mma.sync: b2n_mlp64.cu) implement
    the same arithmetic (fp16 operands, fp32 accumulation, split first layer): they may differ by accumulation order
    only.  Also the density-only mode and the ones-padded inputs of upstream checkpoints."""
    from oracle import nerf_oracle as O
    ops = mods["b2n"].ops
    gen = torch.Generator().manual_seed(11)
    sp = cu(O._fused_init(pos_dim, 16, 64, 1, gen))
    cp = cu(O._fused_init(43, 3, 64, 2, gen))
    x = torch.randn(Pn, pos_dim, device=DEV) * 0.5
    d = torch.nn.functional.normalize(torch.randn(Pn, 3, device=DEV), dim=-1)
    bands = cu(O.fourier_bands(4))
    lib = mods["b2n"]._lib.lib
    out = {}
    prev = ops.INSTANT_FWD_TC
    prev_slots = lib.b2n_debug_instant_fwd_slots(1)
    try:
        for key, tc, slots in (("mma", False, 1), ("tc1", True, 1), ("tc2", True, 2)):       # slots: see b2nerf_debug.h
            ops.INSTANT_FWD_TC = tc
            lib.b2n_debug_instant_fwd_slots(slots)
            rgb, sigma = mods["b2n"].instant_mlp(x, d, bands, sp, cp, pad_value=pad)
            out[key] = (rgb.clone(), sigma.clone(), mods["b2n"].instant_sigma(x, sp, pad_value=pad).clone())
    finally:
        ops.INSTANT_FWD_TC = prev
        lib.b2n_debug_instant_fwd_slots(prev_slots)
    mods["b2n"].check_errors()
    for key in ("tc1", "tc2"):
        tag = f"instant_fwd_{key}[{pos_dim},{Pn}]"
        assert record(f"{tag}:rgb", rel_err(out[key][0], out["mma"][0])) < 2e-5
        assert record(f"{tag}:sigma", rel_err(out[key][1], out["mma"][1])) < 2e-5
        assert torch.equal(out[key][2], out[key][1])            # density-only mode: the same sigma_net arithmetic
    assert torch.equal(out["tc2"][0], out["tc1"][0]) and torch.equal(out["tc2"][1], out["tc1"][1])      # same MMAs, other schedule


@pytest.mark.parametrize("pos_dim,Pn", [(32, 64 * 37 + 5), (53, 777), (64, 130), (32, 5), (32, 200000), (20, 64 * 148 * 3 + 1),
                                        (32, 1500000)])
def test_instant_bwd_tcgen05_wgrad_matches_mma_sync(mods, pos_dim, Pn):
    """The two backward kernels of the fused Instant decoder differ only in where the weight gradients are accumulated
    (tcgen05 + TMEM vs mma.sync + registers): same operands, fp32 accumulation in a different order."""
    from oracle import nerf_oracle as O
    ops = mods["b2n"].ops
    gen = torch.Generator().manual_seed(13)
    sp = cu(O._fused_init(pos_dim, 16, 64, 1, gen)).requires_grad_(True)
    cp = cu(O._fused_init(43, 3, 64, 2, gen)).requires_grad_(True)
    x = (torch.randn(Pn, pos_dim, device=DEV) * 0.5).requires_grad_(True)
    d = torch.nn.functional.normalize(torch.randn(Pn, 3, device=DEV), dim=-1)
    bands = cu(O.fourier_bands(4))
    rgb, sigma = mods["b2n"].instant_mlp(x, d, bands, sp, cp)
    g1, g2 = torch.randn_like(rgb), torch.randn_like(sigma)
    lib = mods["b2n"]._lib.lib
    out = {}
    prev = ops.INSTANT_BWD_TC
    prev_groups = lib.b2n_debug_instant_bwd_groups(1)
    try:
        for key, tc, groups in (("mma", False, 1), ("tc1", True, 1), ("tc3", True, 3)):      # groups: see b2nerf_debug.h
            ops.INSTANT_BWD_TC = tc
            lib.b2n_debug_instant_bwd_groups(groups)
            out[key] = torch.autograd.grad([rgb, sigma], [x, sp, cp], [g1, g2], retain_graph=True)
    finally:
        ops.INSTANT_BWD_TC = prev
        lib.b2n_debug_instant_bwd_groups(prev_groups)
    mods["b2n"].check_errors()
    for key in ("tc1", "tc3"):
        tag = f"instant_bwd_{key}[{pos_dim},{Pn}]"
        assert torch.equal(out[key][0], out["mma"][0])                 # g_x: the same mma.sync chain
        for a_, b_, name in zip(out[key][1:], out["mma"][1:], ("g_sigma_params", "g_color_params")):
            assert record(f"{tag}:{name}", rel_err(a_, b_)) < 2e-5, name
        V3 = out[key][2][64 * 48 + 64 * 64:].view(16, 64)
        assert float(V3[3:].abs().max()) == 0.0


def test_instant_mlp_nonfinite_gradient_propagates(mods):
    """GradScaler's overflow detection needs an inf / NaN incoming gradient to stay visible: the saturating fp16
    conversions of the kernel must not turn it into a large finite step."""
    from oracle import nerf_oracle as O
    gen = torch.Generator().manual_seed(3)
    sp = cu(O._fused_init(32, 16, 64, 1, gen)).requires_grad_(True)
    cp = cu(O._fused_init(43, 3, 64, 2, gen)).requires_grad_(True)
    x = (torch.randn(300, 32, device=DEV) * 0.5).requires_grad_(True)
    d = torch.nn.functional.normalize(torch.randn(300, 3, device=DEV), dim=-1)
    for bad in (float("inf"), float("nan")):
        rgb, sigma = mods["b2n"].instant_mlp(x, d, cu(O.fourier_bands(4)), sp, cp)
        g = torch.ones_like(rgb)
        g[17, 1] = bad
        grads = torch.autograd.grad((rgb * g).sum() + sigma.sum(), [x, sp, cp])
        assert all(not torch.isfinite(t).all() for t in grads[1:])
        assert not torch.isfinite(grads[0]).all()
    rgb, sigma = mods["b2n"].instant_mlp(x, d, cu(O.fourier_bands(4)), sp, cp)       # all-zero gradient: exact zeros out
    grads = torch.autograd.grad((rgb * 0).sum() + (sigma * 0).sum(), [x, sp, cp])
    assert all(float(t.abs().max()) == 0.0 for t in grads)



def _fmlp_case(kind, gen):
    """(state dict, oracle fn, module factory args) of the three fused dynamic-config networks"""
    from oracle import nerf_oracle as O
    if kind == "deform":          # DeformationNetwork 84 -> 128 x3 -> 3 (biases; src/decoders.py:171-195)
        sd = {}
        for i, (o, k) in enumerate([(128, 84), (128, 128), (128, 128), (3, 128)]):
            sd[f"n.net.{2 * i}.weight"], sd[f"n.net.{2 * i}.bias"] = O._linear_init(o, k, gen)
        sd["n.net.6.weight"] = sd["n.net.6.weight"] * 30.0      # the reference's U(+-1e-4) init would hide errors
        return sd, (63, 21), lambda s, a, b, q: O.deformation_net(s, "n", a, b, 4, emulate_bf16=q)
    if kind == "timemod":         # TimeModulationNetwork 21 -> 64 -> 64 sigmoid (src/decoders.py:340-371)
        sd = {}
        for i, (o, k) in enumerate([(64, 21), (64, 64)]):
            sd[f"n.net.{2 * i}.weight"], sd[f"n.net.{2 * i}.bias"] = O._linear_init(o, k, gen)
        return sd, (21, 0), lambda s, a, b, q: O.time_modulation(s, "n", a, 2, emulate_bf16=q)
    sd = {"n.deform_net.params": O._fused_init(88, 3, 64, 2, gen), "n.displacement_scale": torch.tensor(0.1)}
    return sd, (24, 64), lambda s, a, b, q: O.hash_deform_decoder(s, "n", a, b, 64, emulate_bf16=q)


@pytest.mark.parametrize("kind", ["deform", "timemod", "hashdeform"])
@pytest.mark.parametrize("Pn", [1000, 16 * 37 + 3, 5, 40000])
@pytest.mark.parametrize("gscale", [1.0, 65536.0])
def test_fused_mlp_16bit_vs_oracle(mods, bf16_mode, kind, Pn, gscale):
    """b2n_fmlp_fwd/bwd (fp16 operands, power-of-two scaled gradient chain) behind DeformationNetwork /
    TimeModulationNetwork / HashDeformationDecoder: tight against the oracle in the kernel's arithmetic model, the
    north star's 1e-2 against plain fp32."""
    from src import decoders as D
    gen = torch.Generator().manual_seed(11)
    sd, (d0, d1), ref_fn = _fmlp_case(kind, gen)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    a = (torch.randn(Pn, d0, generator=gen) * 0.7).requires_grad_(True)
    b = (torch.rand(Pn, d1, generator=gen)).requires_grad_(True) if d1 else None
    y = ref_fn(sd, a, b, False)
    y_q = ref_fn(sd, a, b, "kernel")
    g_y = torch.randn(y.shape, generator=gen) * gscale
    wrt = [a] + ([b] if d1 else []) + list(sd.values())
    ref = torch.autograd.grad((y_q * g_y).sum(), wrt)
    ref32 = torch.autograd.grad((y * g_y).sum(), wrt)
    if kind == "deform":
        mod = D.DeformationNetwork(63, 21, 128, 4)
    elif kind == "timemod":
        mod = D.TimeModulationNetwork(21, 64, 64, 2)
    else:
        mod = D.HashDeformationDecoder(24, 64, 64)
    mod.load_state_dict({k[2:]: v.detach() for k, v in sd.items()})
    mod = mod.to(DEV)
    a2 = cu(a.detach()).requires_grad_(True)
    b2 = cu(b.detach()).requires_grad_(True) if d1 else None
    launches0 = mods["b2n"]._lib.LAUNCHES["count"]
    y2 = mod(a2, b2) if d1 else mod(a2)
    assert mods["b2n"]._lib.LAUNCHES["count"] - launches0 == 1          # ONE kernel: the fused path ran
    assert y2.shape == y.shape
    tag = f"fmlp_{kind}[{Pn}]"
    assert record(f"{tag}:y_vs_fp32", rel_err(y2.cpu(), y)) < BF16_TOL
    assert record(f"{tag}:y_vs_kernel_model", rel_err(y2.cpu(), y_q)) < 1e-3
    params = dict(mod.named_parameters())
    got = torch.autograd.grad((y2 * cu(g_y)).sum(), [a2] + ([b2] if d1 else []) + [params[k[2:]] for k in sd])
    names = ["g_x0"] + (["g_x1"] if d1 else []) + list(sd)
    for a_, b_, b32, name in zip(got, ref, ref32, names):
        a_ = a_.cpu()
        assert torch.isfinite(a_).all()
        assert record(f"{tag}:{name}_l2_vs_kernel_model", rel_l2(a_, b_)) < 5e-3, name
        # (random-sign output gradients: see the note in test_instant_mlp_fp16_vs_oracle)
        assert record(f"{tag}:{name}_l2_vs_fp32", rel_l2(a_, b32)) < 5e-2, name
    if kind == "hashdeform":      # padded rows of the flat parameter vector never receive gradient
        gp = got[names.index("n.deform_net.params")].cpu()
        assert float(gp[64 * 96 + 64 * 64:].view(16, 64)[3:].abs().max()) == 0.0
        assert float(gp[: 64 * 96].view(64, 96)[:, 88:].abs().max()) == 0.0


def test_fused_mlp_rejects_bad_shapes(mods):
    b2n = mods["b2n"]
    x = torch.zeros(8, 100, device=DEV)
    with pytest.raises(ValueError):
        b2n.fused_mlp(x, None, [torch.zeros(64, 100, device=DEV), torch.zeros(3, 64, device=DEV)], [None, None])
    with pytest.raises(ValueError):
        b2n.fused_mlp(x[:, :8], None, [torch.zeros(32, 8, device=DEV), torch.zeros(3, 32, device=DEV)], [None, None])


# relative-L2 bars of the 16-bit path's PARAMETER GRADIENTS against the reference's fp32 autograd on the golden
# fixtures: the north star's 1e-2 for Part 2 Instant and Part 3 Instant.  Part 4 sums the gradient of its three
# deformation grids / time-modulation net over 60-odd rays only; the CPU model of the kernels' arithmetic
# (tools/bf16_error_budget.py) puts those tensors at 1-1.9e-2 in fp16 (4-5e-2 in bf16), dominated by single ReLU-mask
# flips of the 64-wide deformation decoder -- the bar there is that model's figure, not a tolerance fitted to the kernel.
GRAD_L2_BAR_16BIT = {"part2_instant": 1e-2, "part3_instant": 1e-2, "part4": 2e-2}


@pytest.mark.parametrize("tag", ["part2_instant", "part3_instant", "part4"])
@pytest.mark.parametrize("pert", ["flat", "pert"])
def test_render_rays_golden_bf16(mods, bf16_mode, tag, pert):
    g = load(f"render_{tag}_{pert}")
    model = _model_from(mods, g["cfg"], g["sd"]).train(pert == "pert")
    assert model.decoder.can_fuse(model.dir_representation)
    grid = mods["renderer"].DensityGrid(resolution=g["binary_grid"].shape[0], bound=g["grid_bound"]).to(DEV)
    grid.binary_grid = cu(g["binary_grid"])
    times = cu(g["times"]) if "times" in g else None
    out = mods["renderer"].render_rays(model, cu(g["rays_o"]), cu(g["rays_d"]), g["near"], g["far"],
                                       int(g["n_samples"]), pert == "pert", density_grid=grid, times=times,
                                       bg_color=cu(g["bg"]), _jitter=cu(g["u"]) if pert == "pert" else None)
    name = f"render16_{tag}_{pert}"
    assert record(f"{name}:color", rel_err(out[0].cpu(), g["color"])) < BF16_TOL
    assert record(f"{name}:depth", rel_err(out[1].cpu(), g["depth"])) < BF16_TOL
    assert record(f"{name}:acc", rel_err(out[2].cpu(), g["acc"])) < BF16_TOL
    loss = (out[0] * cu(g["g_color"])).sum()
    if "mean_delta_x" in g:
        loss = loss + (out[3]["mean_delta_x"] * cu(g["g_mdx"])).sum()
    _check_grads(model, loss, g, GRAD_L2_BAR_16BIT[tag], l2=True, tag=name)


@pytest.mark.parametrize("tag", ["part2_instant", "part3_instant", "part4"])
@pytest.mark.parametrize("pert", ["flat", "pert"])
def test_render_rays_16bit_vs_kernel_arithmetic_model(mods, bf16_mode, tag, pert):
    """The same golden inputs through the CPU oracle evaluated in the kernels' own arithmetic model (operand types,
    split first layer, rounded dZ: oracle.OracleField(emulate_bf16="kernel")): gradients agree to 1e-2 relative-L2 per
    tensor -- what separates the CUDA path from its own arithmetic model is accumulation order and rounding ties."""
    from oracle import nerf_oracle as O
    g = load(f"render_{tag}_{pert}")
    model = _model_from(mods, g["cfg"], g["sd"]).train(pert == "pert")
    grid = mods["renderer"].DensityGrid(resolution=g["binary_grid"].shape[0], bound=g["grid_bound"]).to(DEV)
    grid.binary_grid = cu(g["binary_grid"])
    times = cu(g["times"]) if "times" in g else None
    torch.manual_seed(0)
    out = mods["renderer"].render_rays(model, cu(g["rays_o"]), cu(g["rays_d"]), g["near"], g["far"],
                                       int(g["n_samples"]), pert == "pert", density_grid=grid, times=times,
                                       bg_color=cu(g["bg"]), _jitter=cu(g["u"]) if pert == "pert" else None)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "freq" not in k) for k, v in g["sd"].items()}
    if "deformation_grid.encoding.params" in sd:
        sd["deformation_grid.encoding.params"] = sd["deform_grid_start.encoding.params"]
    field = O.OracleField(g["cfg"], sd, emulate_bf16="kernel").train(False)      # the fixtures were made without input noise
    ref = O.render_rays(field, g["rays_o"], g["rays_d"], g["near"], g["far"], int(g["n_samples"]),
                        g["u"] if pert == "pert" else None, binary_grid=g["binary_grid"], grid_bound=g["grid_bound"],
                        bg_color=g["bg"], times=g.get("times"))
    name = f"render16_vs_model_{tag}_{pert}"
    assert record(f"{name}:color", rel_err(out[0].cpu(), ref[0])) < 2e-3
    loss = (out[0] * cu(g["g_color"])).sum()
    loss_ref = (ref[0] * g["g_color"]).sum()
    if "mean_delta_x" in g:
        loss = loss + (out[3]["mean_delta_x"] * cu(g["g_mdx"])).sum()
        loss_ref = loss_ref + (ref[3]["mean_delta_x"] * g["g_mdx"]).sum()
    names = [n for n in g["grads"] if n in sd and sd[n].requires_grad]
    params = dict(model.named_parameters())
    got = torch.autograd.grad(loss, [params[n] for n in names], allow_unused=True)
    want = torch.autograd.grad(loss_ref, [sd[n] for n in names], allow_unused=True)
    for n, a_, b_ in zip(names, got, want):
        if a_ is None or b_ is None:
            continue
        assert record(f"{name}:grad_l2:{n}", rel_l2(a_.cpu(), b_)) < 1e-2, n


# ------------------------------------------------------------------ tcgen05 256-wide decoder
@pytest.mark.parametrize("Pn", [256, 1000, 128 * 37 + 5, 40000])
def test_nerf_mlp256_tcgen05_forward(mods, Pn):
    """NeRFDecoder forward on tcgen05 vs the oracle: tight against the bf16-operand emulation,
    1e-2 against plain fp32; saved activation planes equal the layer outputs."""
    from oracle import nerf_oracle as O
    from b2n import ops
    sd = full_nerf_state_dict(31)
    cfg = dict(mode="part2_nerf", L_embed=10, L_embed_dir=4)
    model = _model_from(mods, cfg, sd).eval()
    torch.manual_seed(11)
    x = (torch.rand(Pn, 3) * 2 - 1) * 1.5
    d = torch.randn(Pn, 3)
    d = d / d.norm(dim=-1, keepdim=True)
    xe = O.fourier_encode(x, sd["representation.freq_bands"])
    de = O.fourier_encode(d, sd["dir_representation.freq_bands"])
    rgb_q, sig_q = O.nerf_decoder(sd, "decoder", xe, de, emulate_bf16=True)
    rgb_f, sig_f = O.nerf_decoder(sd, "decoder", xe, de)
    assert ops.nerf_mlp_supported(model.decoder, 63, 27)
    rgb, sigma, (planes, masks), err = ops.nerf_mlp_forward(model.decoder, cu(xe), cu(de), save=True)
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"tcgen05 pipeline aborted with code {int(err.item())}"
    assert rel_err(rgb.cpu(), rgb_q) < 3e-3 and rel_err(sigma.cpu(), sig_q) < 3e-3
    assert rel_err(rgb.cpu(), rgb_f) < 1e-2 and rel_err(sigma.cpu(), sig_f) < 2e-2
    # saved plane 0 = relu(W0 x + b0) in bf16
    h0 = torch.relu(torch.nn.functional.linear(O.bf16_round(xe), O.bf16_round(sd["decoder.pts_layers.0.weight"]),
                                               sd["decoder.pts_layers.0.bias"]))
    assert rel_err(planes[0].float().cpu(), h0) < 1e-2
    # ReLU bit masks: bit c of row p of slot s == (planes[s][p, c] > 0)
    bits = ((masks.cpu().view(10, Pn, 8, 1) >> torch.arange(32, dtype=torch.int32)) & 1).reshape(10, Pn, 256).bool()
    assert torch.equal(bits[:8], planes[:8].float().cpu() > 0)
    assert torch.equal(bits[9][:, :128], planes[9][:, :128].float().cpu() > 0)


@pytest.mark.parametrize("pos_dim,Pn", [(63, 512), (63, 128 * 37 + 5), (84, 1000), (63, 75000)])
def test_nerf_mlp256_cta_pair_equals_single_cta(mods, bf16_mode, pos_dim, Pn):
    """The cta_group::2 schedule (M = 256 MMAs over an SM pair, half of every weight chunk per CTA) accumulates every
    output element over k in the same order as the single-CTA kernel: forward outputs, saved planes, ReLU masks and
    parameter / input gradients are BIT-identical between the two schedules (ragged tails, odd pair counts, wide input)."""
    from b2n import ops
    from src.decoders import NeRFDecoder
    lib = mods["b2n"]._lib.lib
    torch.manual_seed(5)
    dec = NeRFDecoder(pos_dim=pos_dim, dir_dim=27).to(DEV)
    xe = torch.randn(Pn, pos_dim, device=DEV) * 0.6
    de = torch.randn(Pn, 27, device=DEV) * 0.6
    g_rgb, g_sig = torch.randn(Pn, 3, device=DEV), torch.randn(Pn, 1, device=DEV)
    res = {}
    try:
        for pair in (0, 1):
            lib.b2n_debug_mlp256_set_pair(pair)
            rgb, sigma, (planes, masks), err = ops.nerf_mlp_forward(dec, xe, de, save=True)
            torch.cuda.synchronize()
            assert int(err.item()) == 0, f"pair={pair}: tcgen05 pipeline aborted with code {int(err.item())}"
            x2 = xe.clone().requires_grad_(True)
            r2, s2 = dec(x2, de)
            params = list(dec.parameters())
            grads = torch.autograd.grad((r2 * g_rgb).sum() + (s2 * g_sig).sum(), [x2] + params)
            torch.cuda.synchronize()
            res[pair] = [rgb, sigma, planes[:9], planes[9][:, :128], masks[:9], r2, s2] + list(grads)
    finally:
        lib.b2n_debug_mlp256_set_pair(1)
    for i, (a_, b_) in enumerate(zip(res[0], res[1])):
        if i < 8:                                   # kernel outputs: bit-identical
            assert torch.equal(a_, b_), f"output {i} differs between the schedules"
        else:                                       # weight gradients go through fp32 atomics (order varies run to run)
            assert rel_err(a_.cpu(), b_.cpu()) < 1e-4, i


def test_nerf_mlp256_tcgen05_backward(mods, bf16_mode):
    """Parameter gradients of the tensor-core NeRFDecoder vs autograd through the bf16-emulating oracle."""
    from oracle import nerf_oracle as O
    sd = full_nerf_state_dict(31)
    cfg = dict(mode="part2_nerf", L_embed=10, L_embed_dir=4)
    model = _model_from(mods, cfg, sd).train()
    torch.manual_seed(12)
    Pn = 128 * 21 + 77
    x = (torch.rand(Pn, 3) * 2 - 1) * 1.5
    d = torch.randn(Pn, 3)
    d = d / d.norm(dim=-1, keepdim=True)
    sdg = {k: v.clone().requires_grad_("freq" not in k) for k, v in sd.items()}
    xe = O.fourier_encode(x, sd["representation.freq_bands"])
    de = O.fourier_encode(d, sd["dir_representation.freq_bands"])
    rgb_q, sig_q = O.nerf_decoder(sdg, "decoder", xe, de, emulate_bf16=True)
    g_rgb, g_sig = torch.randn_like(rgb_q), torch.randn_like(sig_q)
    names = [k for k in sdg if k.startswith("decoder.")]
    ref = torch.autograd.grad((rgb_q * g_rgb).sum() + (sig_q * g_sig).sum(), [sdg[k] for k in names])
    rgb, sigma = model(cu(x), cu(d))
    assert rel_err(rgb.cpu(), rgb_q) < 3e-3 and rel_err(sigma.cpu(), sig_q) < 3e-3
    params = dict(model.named_parameters())
    got = torch.autograd.grad((rgb * cu(g_rgb)).sum() + (sigma * cu(g_sig)).sum(), [params[k] for k in names])
    for k, a_, b_ in zip(names, got, ref):
        l2 = float((a_.cpu().double() - b_.double()).norm() / (b_.double().norm() + 1e-30))
        assert l2 < 3e-2, (k, l2)


@pytest.mark.parametrize("pos_dim,Pn", [(84, 128 * 9 + 50), (63, 1000), (84, 256), (96, 300)])
def test_nerf_mlp256_tcgen05_wide_input_and_dx(mods, bf16_mode, pos_dim, Pn):
    """Part 3 canonical decoder (pos_dim = 63 + 21 > 64: two accumulating MMA steps for the layers reading x) and
    the input gradient d x_enc = dZ0 W0 + dZ4 W4[:, 256:] (b2n_nerf_mlp_dx), vs the bf16-emulating oracle."""
    from oracle import nerf_oracle as O
    from src.decoders import NeRFDecoder
    torch.manual_seed(41)
    dec = NeRFDecoder(pos_dim=pos_dim, dir_dim=27)
    sd = {"decoder." + k: v.detach().clone() for k, v in dec.state_dict().items()}
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xe = (torch.randn(Pn, pos_dim) * 0.6).requires_grad_(True)
    de = torch.randn(Pn, 27) * 0.6
    rgb_q, sig_q = O.nerf_decoder(sdg, "decoder", xe, de, emulate_bf16=True)
    g_rgb, g_sig = torch.randn_like(rgb_q), torch.randn_like(sig_q)
    names = list(sdg)
    ref = torch.autograd.grad((rgb_q * g_rgb).sum() + (sig_q * g_sig).sum(), [xe] + [sdg[k] for k in names])
    dec = dec.to(DEV)
    xe2 = cu(xe.detach()).requires_grad_(True)
    launches0 = mods["b2n"]._lib.LAUNCHES["count"]
    rgb, sigma = dec(xe2, cu(de))
    assert mods["b2n"]._lib.LAUNCHES["count"] - launches0 == 2           # pack + the tcgen05 kernel (no fp32 fallback)
    assert rel_err(rgb.cpu(), rgb_q) < 3e-3 and rel_err(sigma.cpu(), sig_q) < 3e-3
    params = dict(dec.named_parameters())
    got = torch.autograd.grad((rgb * cu(g_rgb)).sum() + (sigma * cu(g_sig)).sum(),
                              [xe2] + [params[k[len("decoder."):]] for k in names])
    for k, a_, b_ in zip(["g_x"] + names, got, ref):
        l2 = float((a_.cpu().double() - b_.double()).norm() / (b_.double().norm() + 1e-30))
        assert l2 < 3e-2, (k, l2)


def test_new_entry_points_accept_empty_batches(mods, bf16_mode):
    """P == 0 is a no-op for every fused entry point (an all-empty occupancy mask yields zero active samples)"""
    from src import decoders as D
    dn = D.DeformationNetwork(63, 21, 128, 4).to(DEV)
    out = dn(torch.zeros(0, 63, device=DEV), torch.zeros(0, 21, device=DEV))
    assert out.shape == (0, 3)
    tm = D.TimeModulationNetwork(21, 64, 64, 2).to(DEV)
    assert tm(torch.zeros(0, 21, device=DEV)).shape == (0, 64)
    dec = D.NeRFDecoder(pos_dim=84, dir_dim=27).to(DEV)
    x = torch.zeros(0, 84, device=DEV, requires_grad=True)
    rgb, sigma = dec(x, torch.zeros(0, 27, device=DEV))
    assert rgb.shape == (0, 3) and sigma.shape == (0, 1)
    (rgb.sum() + sigma.sum()).backward()
    assert x.grad.shape == (0, 84)
    assert all(p.grad is not None and float(p.grad.abs().max()) == 0.0 for p in dec.parameters())


@pytest.mark.parametrize("Pn", [2000, 33, 1])
def test_hash_tri_blend_vs_module_form(mods, Pn):
    """b2n_hash_tri_fwd/bwd against w0*grid0(x) + w1*grid1(x) + w2*grid2(x) built from b2n_hash_fwd/bwd + torch ops (the
    form the golden Part-4 fixtures pinned before the fusion), incl. t = 0, 0.5, 1 where one or two weights vanish."""
    b2n = mods["b2n"]
    torch.manual_seed(3)
    geom = b2n.HashGeometry(12, 16, 1.5, 16, 2)
    tabs = [((torch.rand(geom.n_params) * 2 - 1) * 0.3).to(DEV).requires_grad_(True) for _ in range(3)]
    x = ((torch.rand(Pn, 3) * 2 - 1) * 1.6).to(DEV)
    t = torch.rand(Pn, 1)
    t[:: 7] = 0.0
    t[1:: 7] = 0.5
    t[2:: 7] = 1.0
    t = t.to(DEV)
    g = torch.randn(Pn, geom.out_dim, device=DEV)
    ws = [torch.clamp(1.0 - torch.abs(t - a) / 0.5, 0.0, 1.0) for a in (0.0, 0.5, 1.0)]
    tot = ws[0] + ws[1] + ws[2] + 1e-8
    ref = sum((w / tot) * b2n.hash_encode(x, tb, geom, 1.5) for w, tb in zip(ws, tabs))
    g_ref = torch.autograd.grad((ref * g).sum(), tabs)
    out = b2n.ops.hash_tri_blend(x, t, tabs, geom, 1.5)
    assert rel_err(out, ref) < 1e-6
    g_out = torch.autograd.grad((out * g).sum(), tabs)
    for a_, b_ in zip(g_out, g_ref):
        assert rel_err(a_, b_) < 1e-5


def test_hash_encode_properties_full_size(mods):
    """C2-sized point set (2^24 points, L = 16, T = 2^19): the encoding is linear in the table, and because the 8
    trilinear weights of a point sum to 1, the table gradient of sum(out * g) must sum (per level and feature) to the
    column sums of g -- a conservation check on the 2 x 10^9 atomics of the backward."""
    b2n = mods["b2n"]
    torch.manual_seed(4)
    geom = b2n.HashGeometry(16, 16, 1.5, 19, 2)
    Pn = 2 ** 24
    x = (torch.rand(Pn, 3, device=DEV) * 2 - 1) * 1.5
    ta = ((torch.rand(geom.n_params, device=DEV) * 2 - 1) * 0.1)
    tb = ((torch.rand(geom.n_params, device=DEV) * 2 - 1) * 0.1)
    ea, eb = b2n.hash_encode(x, ta, geom, 1.5), b2n.hash_encode(x, tb, geom, 1.5)
    eab = b2n.hash_encode(x, ta + tb, geom, 1.5)
    assert rel_err(eab, ea + eb) < 1e-5
    del ea, eb, eab
    t = ta.clone().requires_grad_(True)
    g = torch.rand(Pn, geom.out_dim, device=DEV)          # positive: no cancellation in the sums
    (b2n.hash_encode(x, t, geom, 1.5) * g).sum().backward()
    col = g.double().sum(0)                               # [L*F]
    off = 0
    for l, (_, _, size, offset, _) in enumerate(geom.levels):
        got = t.grad[offset * 2:(offset + size) * 2].view(-1, 2).double().sum(0)
        assert float((got - col[2 * l:2 * l + 2]).abs().max() / col[2 * l:2 * l + 2].abs().max()) < 1e-4, l


@pytest.mark.parametrize("Pn,kx", [(64, 64), (1000, 128), (64 * 37 + 5, 64), (200000, 64)])
def test_nerf_mlp_wgrad_tcgen05_vs_torch(mods, Pn, kx):
    """b2n_nerf_mlp_wgrad (tcgen05 with MN-major operands straight from the point-major planes, TMA loads, split over P,
    fused bias column sums) against fp32 matmuls of the same bf16 planes; ragged P exercises the zero-filled TMA tails."""
    from b2n._lib import call, ptr, stream
    torch.manual_seed(8)
    dz = (torch.randn(10, Pn, 256, device=DEV) * 0.1).to(torch.bfloat16)
    H = torch.relu(torch.randn(10, Pn, 256, device=DEV)).to(torch.bfloat16)
    xb = torch.randn(Pn, kx, device=DEV).to(torch.bfloat16)
    db = torch.randn(Pn, 64, device=DEV).to(torch.bfloat16)
    dW = torch.zeros(8, 256, 256, device=DEV)
    dW0, dW4x = torch.zeros(256, kx, device=DEV), torch.zeros(256, kx, device=DEV)
    dWv_h, dWv_d = torch.zeros(128, 256, device=DEV), torch.zeros(128, 64, device=DEV)
    gb = torch.zeros(10, 256, device=DEV)
    err = torch.zeros(1, device=DEV, dtype=torch.int32)
    args = (ptr(dz), ptr(H), ptr(xb), kx, ptr(db), Pn, ptr(dW), ptr(dW0), ptr(dW4x), ptr(dWv_h), ptr(dWv_d), ptr(gb), ptr(err),
            stream())
    call("b2n_nerf_mlp_wgrad", *args)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    for l in range(1, 8):
        assert rel_err(dW[l - 1], dz[9 - l].float().t() @ H[l - 1].float()) < 2e-5, l
    assert rel_err(dW[7], dz[1].float().t() @ H[7].float()) < 2e-5
    assert rel_err(dW0, dz[9].float().t() @ xb.float()) < 2e-5
    assert rel_err(dW4x, dz[5].float().t() @ xb.float()) < 2e-5
    assert rel_err(dWv_h, dz[0][:, :128].float().t() @ H[8].float()) < 2e-5
    assert rel_err(dWv_d, dz[0][:, :128].float().t() @ db.float()) < 2e-5
    ref_b = dz.float().sum(1)
    assert rel_err(gb[1:], ref_b[1:]) < 2e-5 and rel_err(gb[0, :128], ref_b[0, :128]) < 2e-5
    # accumulation semantics: a second call doubles the result
    call("b2n_nerf_mlp_wgrad", *args)
    assert rel_err(dW[0], 2 * (dz[8].float().t() @ H[0].float())) < 2e-5


# ------------------------------------------------------------------ render_image / render_image_safe
def _image_case(mods, H=23, W=31):
    """a full 8x256 vanilla field and the rays of one small synthetic view"""
    from b2n import synthetic
    sd = full_nerf_state_dict(31)
    model = _model_from(mods, dict(mode="part2_nerf", L_embed=10, L_embed_dir=4), sd).eval()
    pose = synthetic.hemisphere_poses(3, seed=4)[1]
    ro, rd = synthetic.image_rays(pose, H=H, W=W)
    return model, sd, ro.view(H, W, 3), rd.view(H, W, 3)


def test_render_image_chunked_equals_unchunked_equals_oracle(mods):
    """render_image (reference src/renderer.py:387-418): [H,W,3] in, chunk loop over render_rays(perturb=False), [H,W,3]
    out.  A ragged last chunk, chunk = 1 ray short of / beyond the image and one single chunk all give the same image,
    and that image is the CPU oracle's (fp32 path, 1e-4)."""
    from oracle import nerf_oracle as O
    model, sd, ro, rd = _image_case(mods)
    H, W = ro.shape[:2]
    render_image = mods["renderer"].render_image
    with torch.no_grad():
        whole = render_image(model=model, rays_o=cu(ro), rays_d=cu(rd), near=2.0, far=6.0, n_samples=48, chunk=10 ** 6,
                             white_bkgd=True)
        assert whole.shape == (H, W, 3) and whole.dtype == torch.float32
        for chunk in (100, H * W - 1, H * W, 7):
            img = render_image(model=model, rays_o=cu(ro), rays_d=cu(rd), near=2.0, far=6.0, n_samples=48, chunk=chunk,
                               white_bkgd=True)
            assert torch.equal(img, whole), chunk                  # per-ray arithmetic does not depend on the batch
        ref = O.render_rays(O.OracleField(dict(mode="part2_nerf", L_embed=10, L_embed_dir=4), sd), ro.reshape(-1, 3),
                            rd.reshape(-1, 3), 2.0, 6.0, 48, None, white_bkgd=True)[0].view(H, W, 3)
    assert record("render_image:rgb_vs_oracle", rel_err(whole.cpu(), ref)) < TOL
    with torch.no_grad():                                          # an empty scene shows the background flag
        model.decoder.sigma_layer.bias.fill_(-1e4)
        white = render_image(model=model, rays_o=cu(ro), rays_d=cu(rd), near=2.0, far=6.0, n_samples=48, chunk=256,
                             white_bkgd=True)
        black = render_image(model=model, rays_o=cu(ro), rays_d=cu(rd), near=2.0, far=6.0, n_samples=48, chunk=256,
                             white_bkgd=False)
    assert float((white - 1.0).abs().max()) < 1e-6 and float(black.abs().max()) < 1e-6


def test_render_image_800x800_single_chunk(mods, bf16_mode):
    """one NeRF-Synthetic-sized frame (640 000 rays x 64 samples = 41 M points) in a single chunk on the tcgen05 decoder;
    a strided sub-image rendered on its own agrees to the 16-bit class"""
    from b2n import synthetic
    sd = full_nerf_state_dict(31)
    model = _model_from(mods, dict(mode="part2_nerf", L_embed=10, L_embed_dir=4), sd).eval()
    pose = synthetic.hemisphere_poses(3, seed=4)[2]
    ro, rd = (t.view(800, 800, 3).to(DEV) for t in synthetic.image_rays(pose))
    with torch.no_grad():
        img = mods["renderer"].render_image(model=model, rays_o=ro, rays_d=rd, near=2.0, far=6.0, n_samples=64,
                                            chunk=800 * 800, white_bkgd=True)
        sub = mods["renderer"].render_image(model=model, rays_o=ro[::40, ::40].contiguous(), rays_d=rd[::40, ::40].contiguous(),
                                            near=2.0, far=6.0, n_samples=64, chunk=123, white_bkgd=True)
    mods["b2n"].check_errors()
    assert img.shape == (800, 800, 3) and torch.isfinite(img).all()
    assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0 + 1e-5
    assert rel_err(img[::40, ::40], sub) < 1e-5


def test_render_image_safe_halves_the_chunk_on_oom(mods):
    """render_image_safe (reference src/utils.py:39-76): a CUDA OOM inside the renderer halves ``chunk`` (not below
    1024) and retries; any other error and an OOM at the floor propagate."""
    from src.utils import render_image_safe
    model, sd, ro, rd = _image_case(mods, H=40, W=64)
    seen = []

    def flaky(model, rays_o, rays_d, near, far, n_samples, chunk, white_bkgd):
        seen.append(chunk)
        if chunk > 2048:
            raise torch.cuda.OutOfMemoryError("synthetic OOM")
        return mods["renderer"].render_image(model=model, rays_o=rays_o, rays_d=rays_d, near=near, far=far,
                                             n_samples=n_samples, chunk=chunk, white_bkgd=white_bkgd)

    with torch.no_grad():
        img = render_image_safe(flaky, model, cu(ro), cu(rd), 2.0, 6.0, 32, 16384, True)
        ref = mods["renderer"].render_image(model=model, rays_o=cu(ro), rays_d=cu(rd), near=2.0, far=6.0, n_samples=32,
                                            chunk=4096, white_bkgd=True)
    assert seen == [16384, 8192, 4096, 2048] and torch.equal(img, ref)

    def always_oom(**kw):
        seen.append(kw["chunk"])
        raise torch.cuda.OutOfMemoryError("synthetic OOM")
    seen.clear()
    with pytest.raises(torch.cuda.OutOfMemoryError):
        render_image_safe(always_oom, model, cu(ro), cu(rd), 2.0, 6.0, 32, 3000, True)
    assert seen == [3000, 1500, 1024]

    def other(**kw):
        raise ValueError("not an OOM")
    with pytest.raises(ValueError):
        render_image_safe(other, model, cu(ro), cu(rd), 2.0, 6.0, 32, 3000, True)
