"""b2n.checkpoint: checkpoints written by the reference's run.py -- here by the reference's own files on the
tcnn shim, tests/golden/make_golden.py::gen_checkpoint: fp16 ``params``, FullyFusedMLP inputs padded with ONES -- load
into this package's modules and reproduce the reference's outputs and gradients."""
import os

import pytest
import torch

from _util import GOLDEN, load, record, rel_err

CKPT = os.path.join(GOLDEN, "ckpt_part4_upstream.pth")


def _model(cfg):
    import __graft_entry__ as ge
    ge.build()
    from src.core import NeuralField
    return NeuralField(cfg)


def test_loader_host_logic():
    """no GPU needed: dtype conversion, alias keys, report, density grid, strictness, the padding switch"""
    from b2n.checkpoint import load_reference_checkpoint
    from src.renderer import DensityGrid
    g = load("ckpt_part4_upstream_io")
    model = _model(g["cfg"])
    raw = torch.load(CKPT, map_location="cpu", weights_only=False)
    assert raw["model_state_dict"]["decoder.sigma_net.params"].dtype == torch.float16
    assert "deformation_grid.encoding.params" in raw["model_state_dict"]          # the alias is saved twice
    grid = DensityGrid(resolution=8, bound=g["cfg"]["scene_bound"])
    rep = load_reference_checkpoint(model, CKPT, pad_value=1.0, density_grid=grid)
    assert rep["step"] == 1234 and abs(rep["val_psnr"] - 29.56) < 1e-9 and rep["density_grid"] and rep["pad_value"] == 1.0
    assert "decoder.sigma_net.params" in rep["converted"] and not rep["missing"] and not rep["unexpected"]
    assert all(p.dtype == torch.float32 for p in model.parameters())
    assert torch.equal(model.decoder.sigma_net.params.detach(), raw["model_state_dict"]["decoder.sigma_net.params"].float())
    assert model.deformation_grid.encoding.params is model.deform_grid_start.encoding.params
    assert model.decoder.color_net.input_pad_value == 1.0 and model.deform_decoder.deform_net.input_pad_value == 1.0
    assert torch.equal(grid.binary_grid, g["binary_grid"]) and torch.equal(grid.grid, g["grid"])
    # a bare state_dict without the alias, module.-prefixed keys
    sd = {"module." + k: v for k, v in raw["model_state_dict"].items() if not k.startswith("deformation_grid.")}
    model2 = _model(g["cfg"])
    load_reference_checkpoint(model2, sd)
    assert torch.equal(model2.canonical_repr.encoding.params, model.canonical_repr.encoding.params)
    assert model2.decoder.color_net.input_pad_value == 0.0
    # strictness
    bad = dict(raw["model_state_dict"])
    bad.pop("decoder.color_net.params")
    with pytest.raises(KeyError, match="color_net"):
        load_reference_checkpoint(_model(g["cfg"]), bad)
    bad = dict(raw["model_state_dict"])
    bad["decoder.color_net.params"] = bad["decoder.color_net.params"][:-5]
    with pytest.raises(ValueError, match="FullyFusedMLP"):
        load_reference_checkpoint(_model(g["cfg"]), bad)


def test_hash_table_reindexing_between_neighbouring_resolutions():
    """level 3 of the stock geometry is 54^3 or 55^3 depending on the libm that computed ceil(scale): a
    table of the other layout is re-indexed by lattice coordinate"""
    import __graft_entry__ as ge
    ge.build()
    import b2n
    from b2n.checkpoint import convert_hash_table
    geom = b2n.HashGeometry(6, 16, 1.5, 19, 2)              # levels 16, 24, 36, 54, 81, 122 -- all dense but the last two? check below
    res = [l[1] for l in geom.levels]
    assert res[:4] == [16, 24, 36, 54] and not geom.levels[3][4]
    # build the "55" layout: same levels, level 3 stored with res 55
    sizes = [l[2] for l in geom.levels]
    alt = list(sizes)
    alt[3] = (55 ** 3 + 7) // 8 * 8
    src = torch.arange(sum(alt) * 2, dtype=torch.float32)
    out = convert_hash_table(src, geom)
    assert out.numel() == geom.n_params
    off_src3, off_dst3 = sum(alt[:3]), geom.levels[3][3]
    for (x, y, z) in ((0, 0, 0), (53, 0, 0), (7, 31, 2), (53, 53, 53)):
        e_src, e_dst = x + y * 55 + z * 55 * 55, x + y * 54 + z * 54 * 54
        assert torch.equal(out[(off_dst3 + e_dst) * 2:(off_dst3 + e_dst) * 2 + 2], src[(off_src3 + e_src) * 2:(off_src3 + e_src) * 2 + 2])
    # the other levels pass through unchanged (shifted by the size difference behind level 3)
    assert torch.equal(out[:off_dst3 * 2], src[:off_src3 * 2])
    o4s, o4d = sum(alt[:4]), geom.levels[4][3]
    assert torch.equal(out[o4d * 2:], src[o4s * 2:])
    with pytest.raises(ValueError, match="hash table"):
        convert_hash_table(src[:-16], geom)


def test_unpadded_fused_mlp_export_is_padded():
    import __graft_entry__ as ge
    ge.build()
    from b2n.checkpoint import convert_fused_mlp
    from src.decoders import FusedMLP
    mlp = FusedMLP(43, 3, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 64,
                           "n_hidden_layers": 2})
    W = [torch.randn(64, 43), torch.randn(64, 64), torch.randn(3, 64)]
    flat = convert_fused_mlp(torch.cat([w.reshape(-1) for w in W]).half(), mlp)
    mats = []
    off = 0
    for r, c in mlp.shapes:
        mats.append(flat[off:off + r * c].view(r, c))
        off += r * c
    assert torch.equal(mats[0][:, :43], W[0].half().float()) and float(mats[0][:, 43:].abs().max()) == 0.0
    assert torch.equal(mats[2][:3], W[2].half().float()) and float(mats[2][3:].abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_upstream_style_checkpoint_reproduces_reference_outputs(precision):
    """the fixture's outputs / gradients were produced by the reference's own modules on the shim with ones in the padded
    FullyFusedMLP inputs: fp32 path 1e-4, the 16-bit fused kernels 1e-2; with pad_value = 0 the outputs must differ"""
    import b2n
    from b2n.checkpoint import load_reference_checkpoint
    g = load("ckpt_part4_upstream_io")
    dev = "cuda"
    b2n.set_mlp_precision(precision)
    try:
        model = _model(g["cfg"])
        load_reference_checkpoint(model, CKPT, pad_value=1.0)
        model = model.to(dev).eval()
        x, d, t = (g[k].to(dev) for k in ("x", "d", "t"))
        rgb, sigma, dx = model(x, d, t=t)
        tol = 1e-4 if precision == "fp32" else 1e-2
        assert record(f"ckpt[{precision}]:rgb", rel_err(rgb.cpu(), g["rgb"])) < tol
        assert record(f"ckpt[{precision}]:sigma", rel_err(sigma.cpu(), g["sigma"])) < tol
        assert record(f"ckpt[{precision}]:dx", rel_err(dx.cpu(), g["dx"])) < tol
        loss = (rgb * g["g_rgb"].to(dev)).sum() + (sigma * g["g_sigma"].to(dev)).sum() + (dx * g["g_dx"].to(dev)).sum()
        params = dict(model.named_parameters())
        names = list(g["grads"])
        grads = torch.autograd.grad(loss, [params[n] for n in names])
        for n, gr in zip(names, grads):
            ref = g["grads"][n]
            if precision == "fp32":
                assert record(f"ckpt[fp32]:grad:{n}", rel_err(gr.cpu(), ref)) < 1e-4, n
            else:
                l2 = float((gr.cpu().double() - ref.double()).norm() / ref.double().norm())
                assert record(f"ckpt[bf16]:grad_l2:{n}", l2) < 5e-2, n
        # the padded input columns are alive: color_net 43 -> 48 inputs, columns 43..47 of its first matrix
        gc = grads[names.index("decoder.color_net.params")][: 64 * 48].view(64, 48)
        assert float(gc[:, 43:].abs().max()) > 0.0
        ref_c = g["grads"]["decoder.color_net.params"][: 64 * 48].view(64, 48)
        assert rel_err(gc[:, 43:].cpu(), ref_c[:, 43:]) < (1e-4 if precision == "fp32" else 5e-2)
        # the same weights read with zero padding give a different network
        for m in (model.decoder.sigma_net, model.decoder.color_net, model.deform_decoder.deform_net):
            m.input_pad_value = 0.0
        rgb0, _, _ = model(x, d, t=t)
        assert rel_err(rgb0.cpu(), g["rgb"]) > 10 * tol
    finally:
        b2n.set_mlp_precision("fp32")
