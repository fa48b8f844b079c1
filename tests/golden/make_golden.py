#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's ``src`` package read-only (with oracle.tcnn_shim
standing in for the absent tinycudann), runs it on small seeded inputs on CPU
and stores inputs, weights and outputs as ``.npz``.  The fixtures are what
pins oracle/nerf_oracle.py (tests/test_oracle_golden.py) and, through the
same files, the CUDA path (tests/test_gpu_*.py).  Nothing here is imported by
the product.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("B2N_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import tcnn_shim  # noqa: E402

tcnn_shim.install()
sys.path.insert(0, REF)
from src.core import NeuralField  # noqa: E402  (reference)
from src import renderer as R  # noqa: E402  (reference)
from src.embeddings import FourierRepresentation  # noqa: E402

torch.set_num_threads(4)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        if isinstance(v, (dict, list)) and not isinstance(v, np.ndarray):
            v = np.frombuffer(json.dumps(v).encode(), dtype=np.uint8)
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


def sd_arrays(model, prefix="sd::"):
    return {prefix + k: v for k, v in model.state_dict().items()}


# ---------------------------------------------------------------- sampling
def gen_sampling():
    for N, B in ((64, 37), (128, 5), (2, 3), (1, 4)):
        torch.manual_seed(7)
        z_pert = R.sample_stratified(2.0, 6.0, N, B, "cpu", True)
        torch.manual_seed(7)
        u = torch.rand(B, N)
        z_flat = R.sample_stratified(2.0, 6.0, N, B, "cpu", False)
        save(f"sampling_N{N}", near=np.float64(2.0), far=np.float64(6.0), u=u, z_pert=z_pert, z_flat=z_flat)


# ---------------------------------------------------------------- occupancy
def gen_mask():
    torch.manual_seed(3)
    for Rr, bound in ((4, 1.5), (16, 1.0), (128, 1.5)):
        grid = R.DensityGrid(resolution=Rr, bound=bound, threshold=0.01)
        grid.binary_grid = torch.rand(Rr, Rr, Rr) < 0.4
        pts = (torch.rand(4000, 3) * 2 - 1) * bound * 1.3
        step = 2 * bound / Rr
        edge = torch.tensor([-bound, bound, -bound - 0.1 * step, -bound - 0.99 * step, -bound - step,
                             -bound - 1.01 * step, bound - 1e-7, bound + 1e-7, 0.0, -0.0, step, -step,
                             np.nextafter(np.float32(bound), np.float32(0)), -1.6, -1.9, 1.5])
        gx, gy, gz = torch.meshgrid(edge, edge[:6], edge[:4], indexing="ij")
        adv = torch.stack([gx, gy, gz], -1).reshape(-1, 3)
        pts = torch.cat([pts, adv, adv[:, [2, 0, 1]], adv[:, [1, 2, 0]]]).float()
        m = grid.get_active_mask(pts)
        save(f"mask_R{Rr}", bound=np.float64(bound), binary_grid=grid.binary_grid, pts=pts, mask=m)


# ---------------------------------------------------------------- compositing
def gen_composite():
    torch.manual_seed(11)
    for tag, B, N in (("small", 33, 16), ("n128", 9, 128), ("n1", 6, 1), ("n65", 5, 65)):
        rgb = torch.rand(B, N, 3, requires_grad=True)
        sigma = (torch.rand(B, N) * 4).requires_grad_(True)
        with torch.no_grad():
            sigma[0] = 0.0                       # empty ray
            sigma[1] = 1e4                       # opaque at first sample
            if N > 3:
                sigma[2, N // 2] = 1e6           # wall mid-ray
                sigma[3] *= 1e-4
        z = torch.sort(torch.rand(B, N) * 4 + 2, dim=-1)[0]
        rays_d = torch.randn(B, 3)
        rays_d[0] = rays_d[0] / rays_d[0].norm()
        for bgtag, bg in (("bgvec", torch.rand(3)), ("bgray", torch.rand(B, 3)), ("nobg", None)):
            c, d, a = R.volume_render(rgb, sigma, z, rays_d, bg_color=bg)
            gc, gd, ga = torch.randn_like(c), torch.randn_like(d), torch.randn_like(a)
            loss = (c * gc).sum() + (d * gd).sum() + (a * ga).sum()
            g_rgb, g_sigma = torch.autograd.grad(loss, [rgb, sigma])
            save(f"composite_{tag}_{bgtag}", rgb=rgb, sigma=sigma, z=z, rays_d=rays_d,
                 bg=(bg if bg is not None else torch.zeros(0)), color=c, depth=d, acc=a,
                 g_color=gc, g_depth=gd, g_acc=ga, g_rgb=g_rgb, g_sigma=g_sigma)


# ---------------------------------------------------------------- Fourier PE
def gen_pe():
    torch.manual_seed(5)
    for D, L in ((3, 10), (3, 4), (1, 10), (1, 6), (3, 0)):
        enc = FourierRepresentation(input_dim=D, L=L, use_encoding=True)
        x = ((torch.rand(257, D) * 2 - 1) * (6.0 if D == 3 else 1.0)).requires_grad_(True)
        y = enc(x)
        gy = torch.randn_like(y)
        gx, = torch.autograd.grad((y * gy).sum(), x)
        save(f"pe_D{D}_L{L}", x=x, bands=enc.freq_bands, y=y, g_y=gy, g_x=gx)


# ---------------------------------------------------------------- fields + render_rays
SMALL_HASH = dict(n_levels=6, n_features_per_level=2, log2_hashmap_size=11, base_resolution=4, per_level_scale=1.7)

FIELD_CFGS = {
    "part2_nerf": dict(mode="part2_nerf", L_embed=6, L_embed_dir=3, hidden_dim=64, num_layers=6, skip_layer=3, view_dim=32),
    "part2_nerf_full": dict(mode="part2_nerf", L_embed=10, L_embed_dir=4, hidden_dim=256, num_layers=8, skip_layer=4, view_dim=128),
    "part2_instant": dict(mode="part2_instant", scene_bound=1.5, L_embed_dir=4, hidden_dim=64, **SMALL_HASH),
    "part3_nerf": dict(mode="part3", canonical_type="nerf", L_embed=5, L_embed_dir=3, L_embed_time=4, L_embed_canon=5,
                       hidden_dim=64, num_layers=5, skip_layer=2, view_dim=32, deform_hidden_dim=32, deform_num_layers=3),
    "part3_dtc": dict(mode="part3", canonical_type="nerf", direct_time_conditioning=True, L_embed=5, L_embed_dir=3,
                      L_embed_time=6, L_embed_canon=5, hidden_dim=64, num_layers=5, skip_layer=2, view_dim=32,
                      deform_hidden_dim=32, deform_num_layers=3),
    "part3_instant": dict(mode="part3", canonical_type="instant", scene_bound=1.5, L_embed=5, L_embed_dir=4, L_embed_time=10,
                          hidden_dim=64, deform_hidden_dim=32, deform_num_layers=4, **SMALL_HASH),
    "part4": dict(mode="part4", scene_bound=1.5, L_embed_dir=4, L_embed_time=10, time_modulation_dim=64,
                  time_modulation_layers=2, deform_n_levels=4, deform_n_features_per_level=2, deform_log2_hashmap_size=10,
                  deform_base_resolution=4, deform_per_level_scale=1.6, deform_hidden_dim=64, hidden_dim=64, **SMALL_HASH),
}


def build(cfg, seed):
    torch.manual_seed(seed)
    model = NeuralField(cfg)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("encoding.params"):
                p.mul_(3000.0)                     # U(+-1e-4) init would make hash features negligible
        if hasattr(model, "deform_net"):
            last = model.deform_net.net[-1]
            last.weight.mul_(300.0)                # displacement large enough to matter
    return model


def gen_fields():
    for tag, cfg in FIELD_CFGS.items():
        if tag == "part2_nerf_full":
            continue
        model = build(cfg, 21).eval()
        torch.manual_seed(22)
        P = 301
        x = ((torch.rand(P, 3) * 2 - 1) * 1.7)
        d = torch.randn(P, 3)
        d = d / d.norm(dim=-1, keepdim=True)
        t = torch.rand(P, 1)
        dyn = cfg["mode"] in ("part3", "part4")
        out = model(x, d, t=t) if dyn else model(x, d)
        gs = [torch.randn_like(o) for o in out]
        loss = sum((o * g).sum() for o, g in zip(out, gs))
        names = [n for n, p in model.named_parameters()]
        grads = torch.autograd.grad(loss, [p for _, p in model.named_parameters()], allow_unused=True)
        arrs = dict(cfg=cfg, x=x, d=d, t=t, rgb=out[0], sigma=out[1], g_rgb=gs[0], g_sigma=gs[1])
        if dyn:
            arrs.update(dx=out[2], g_dx=gs[2])
        arrs.update(sd_arrays(model))
        for n, g in zip(names, grads):
            if g is not None and g.numel() <= 70000:
                arrs["grad::" + n] = g
            elif g is not None:
                arrs["gradsum::" + n] = torch.stack([g.double().sum(), g.double().abs().sum(), (g.double() ** 2).sum()])
                arrs["gradhead::" + n] = g.reshape(-1)[:4096]
        save(f"field_{tag}", **arrs)


def gen_render():
    for tag, cfg in FIELD_CFGS.items():
        full = tag == "part2_nerf_full"
        model = build(cfg, 31)
        dyn = cfg["mode"] in ("part3", "part4")
        hashy = "n_levels" in cfg
        B, N = (16, 64) if full else (41, 24)
        torch.manual_seed(32)
        ang = torch.rand(B) * 6.28
        ro = torch.stack([4 * torch.cos(ang), 4 * torch.sin(ang), torch.rand(B) * 2], -1)
        rd = -ro / ro.norm(dim=-1, keepdim=True) + 0.15 * torch.randn(B, 3)
        rd = rd / rd.norm(dim=-1, keepdim=True)
        times = torch.rand(B, 1) if dyn else None
        bg = torch.rand(3)
        grid = None
        if hashy:
            grid = R.DensityGrid(resolution=16, bound=cfg["scene_bound"], threshold=0.01)
            gp = torch.stack(torch.meshgrid(*[torch.linspace(-1, 1, 16)] * 3, indexing="ij"), -1)
            grid.binary_grid = gp.norm(dim=-1) < 0.8
        for perturb in (False, True):
            model.train(perturb)       # train mode only matters for coord noise (off in these configs)
            torch.manual_seed(33)
            out = R.render_rays(model, ro, rd, 2.0, 6.0, N, perturb, density_grid=grid, times=times, bg_color=bg)
            torch.manual_seed(33)
            u = torch.rand(B, N) if perturb else torch.zeros(0)
            c, dep, acc = out[:3]
            gc = torch.randn_like(c)
            loss = (c * gc).sum()
            extra = {}
            if dyn and "mean_delta_x" in out[3]:
                mdx = out[3]["mean_delta_x"]
                gm = torch.randn_like(mdx)
                loss = loss + (mdx * gm).sum()
                extra = dict(mean_delta_x=mdx, g_mdx=gm)
            names = [n for n, _ in model.named_parameters()]
            grads = torch.autograd.grad(loss, [p for _, p in model.named_parameters()], allow_unused=True)
            arrs = dict(cfg=cfg, rays_o=ro, rays_d=rd, near=np.float64(2.0), far=np.float64(6.0),
                        n_samples=np.int64(N), u=u, bg=bg, color=c, depth=dep, acc=acc, g_color=gc, **extra)
            if times is not None:
                arrs["times"] = times
            if grid is not None:
                arrs["binary_grid"] = grid.binary_grid
                arrs["grid_bound"] = np.float64(grid.bound)
            if not full:
                arrs.update(sd_arrays(model))
            else:
                arrs["seed"] = np.int64(31)
            for n, g in zip(names, grads):
                if g is None:
                    continue
                if g.numel() <= 20000 and not full:
                    arrs["grad::" + n] = g
                else:
                    arrs["gradsum::" + n] = torch.stack([g.double().sum(), g.double().abs().sum(), (g.double() ** 2).sum()])
            save(f"render_{tag}_{'pert' if perturb else 'flat'}", **arrs)
        # the 256-wide weights are not stored (2.4 MB): tests rebuild them bit-identically with
        # torch.manual_seed(31) + nn.Linear in construction order (tests/_util.py::full_nerf_state_dict)


# ---------------------------------------------------------------- hash KAT
def gen_hash_kat():
    import tinycudann as tcnn
    cfgs = {"c2": dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=1.5),
            "c5canon": dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=20, base_resolution=16, per_level_scale=1.5),
            "c5deform": dict(n_levels=12, n_features_per_level=2, log2_hashmap_size=16, base_resolution=16, per_level_scale=1.5),
            "small": SMALL_HASH}
    for tag, c in cfgs.items():
        enc = tcnn.Encoding(3, dict(otype="HashGrid", **c))
        lv = enc.levels
        torch.manual_seed(41)
        x = torch.rand(64, 3)
        x[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.5, 0.5, 0.5], [1, 1, 0], [0.999999, 1e-7, 0.25]])
        # integer KAT: base corner index per level (python ints, independent of torch)
        kat = np.zeros((len(lv), 64, 8), dtype=np.int64)
        xs = x.numpy()
        for li, L in enumerate(lv):
            s = np.float32(L.scale)
            pos = (xs * s + np.float32(0.5)).astype(np.float32)
            g = np.floor(pos).astype(np.int64)
            for p in range(64):
                for cidx in range(8):
                    cx, cy, cz = (int(g[p, 0]) + (cidx & 1), int(g[p, 1]) + ((cidx >> 1) & 1), int(g[p, 2]) + ((cidx >> 2) & 1))
                    if L.hashed:
                        h = (cx & 0xFFFFFFFF) ^ ((cy * 2654435761) & 0xFFFFFFFF) ^ ((cz * 805459861) & 0xFFFFFFFF)
                    else:
                        h = (cx + cy * L.res + cz * L.res * L.res) & 0xFFFFFFFF
                    kat[li, p, cidx] = h % L.size + L.offset
        save(f"hash_kat_{tag}", cfg=c, x=x, corner_entries=kat,
             level_scale=np.array([L.scale for L in lv], dtype=np.float32),
             level_res=np.array([L.res for L in lv]), level_size=np.array([L.size for L in lv]),
             level_offset=np.array([L.offset for L in lv]), level_hashed=np.array([L.hashed for L in lv]))


# ---------------------------------------------------------------- density grid update
def gen_grid_update():
    for tag in ("part2_instant", "part3_instant", "part4"):
        cfg = FIELD_CFGS[tag]
        model = build(cfg, 51).eval()
        with torch.no_grad():                      # push sigma around the threshold so the grid is mixed
            model.decoder.sigma_net.params.mul_(6.0)
        time = torch.tensor([[0.3]]) if cfg["mode"] == "part3" else None
        probe = R.DensityGrid(resolution=12, bound=cfg["scene_bound"], threshold=0.0)
        probe.update(model, device="cpu", time=time)
        thr = float(probe.grid.flatten().kthvalue(int(0.6 * 12 ** 3))[0]) * 1.0001   # ~40 % of voxels active
        grid = R.DensityGrid(resolution=12, bound=cfg["scene_bound"], threshold=thr)
        r1 = grid.update(model, device="cpu", time=time)
        g1, b1 = grid.grid.clone(), grid.binary_grid.clone()
        time2 = torch.tensor([[0.8]]) if cfg["mode"] == "part3" else None
        r2 = grid.update(model, device="cpu", time=time2, decay=0.95)
        save(f"gridupdate_{tag}", cfg=cfg, threshold=np.float64(thr), ratio1=np.float64(r1), ratio2=np.float64(r2),
             grid1=g1, binary1=b1, grid2=grid.grid, binary2=grid.binary_grid, **sd_arrays(model))


# ---------------------------------------------------------------- an upstream-style checkpoint
def gen_checkpoint():
    """What run.py:2084-2092 writes for a Part-4 model, in the two ways an upstream-tcnn checkpoint can differ from this
    package's own: the tcnn ``params`` tensors stored in fp16, and FullyFusedMLP inputs padded with ONES (the padded
    weight columns act as a bias).  The reference's own files produce the outputs, running on the shim with that
    padding and with the fp16-rounded parameters."""
    tcnn_shim.INPUT_PAD_VALUE = 1.0
    try:
        cfg = FIELD_CFGS["part4"]
        model = build(cfg, 61).eval()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith(".params"):
                    p.copy_(p.half().float())                      # the values an fp16 export can hold
                if n.endswith("_net.params"):                      # make the padded columns matter
                    p.add_(torch.randn_like(p) * 0.05)
                    p.copy_(p.half().float())
        torch.manual_seed(62)
        P = 200
        x = (torch.rand(P, 3) * 2 - 1) * 1.4
        d = torch.randn(P, 3)
        d = d / d.norm(dim=-1, keepdim=True)
        t = torch.rand(P, 1)
        rgb, sigma, dx = model(x, d, t=t)
        gs = [torch.randn_like(o) for o in (rgb, sigma, dx)]
        loss = (rgb * gs[0]).sum() + (sigma * gs[1]).sum() + (dx * gs[2]).sum()
        names = ["decoder.sigma_net.params", "decoder.color_net.params", "deform_decoder.deform_net.params"]
        params = dict(model.named_parameters())
        grads = torch.autograd.grad(loss, [params[n] for n in names])
        grid = R.DensityGrid(resolution=8, bound=cfg["scene_bound"], threshold=0.01)
        grid.binary_grid = torch.rand(8, 8, 8) < 0.5
        grid.grid = torch.rand(8, 8, 8)
        sd = {k: (v.half() if k.endswith(".params") else v.clone()) for k, v in model.state_dict().items()}
        torch.save({"model_state_dict": sd, "config": cfg, "step": 1234, "val_psnr": 29.56,
                    "density_grid": grid.state_dict()}, os.path.join(HERE, "ckpt_part4_upstream.pth"))
        save("ckpt_part4_upstream_io", cfg=cfg, x=x, d=d, t=t, rgb=rgb, sigma=sigma, dx=dx, g_rgb=gs[0], g_sigma=gs[1],
             g_dx=gs[2], grid=grid.grid, binary_grid=grid.binary_grid,
             **{"grad::" + n: g for n, g in zip(names, grads)})
        print("ckpt_part4_upstream.pth", os.path.getsize(os.path.join(HERE, "ckpt_part4_upstream.pth")) // 1024, "KiB")
    finally:
        tcnn_shim.INPUT_PAD_VALUE = 0.0


# ---------------------------------------------------------------- datasets (tiny synthetic scene on disk)
def gen_dataset():
    from PIL import Image
    from src.dataset import BlenderDataset, DynamicDataset          # reference
    root = os.path.join(HERE, "dataset_tiny")
    os.makedirs(os.path.join(root, "train"), exist_ok=True)
    rng = np.random.RandomState(5)
    frames = []
    for i in range(6):
        img = rng.randint(0, 256, size=(20, 16, 4), dtype=np.uint8)
        img[:5, :, 3] = 0                                            # some fully transparent pixels
        Image.fromarray(img, "RGBA").save(os.path.join(root, "train", f"r_{i}.png"))
        a = 0.7 * i
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=np.float32)
        c2w[:3, 3] = [4 * np.cos(a), 4 * np.sin(a), 0.5 * i]
        frames.append({"file_path": f"./train/r_{i}", "transform_matrix": c2w.tolist(), "time": i / 5.0})
    with open(os.path.join(root, "transforms_train.json"), "w") as f:
        json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, f)
    for cls, tag in ((BlenderDataset, "blender"), (DynamicDataset, "dynamic")):
        for scale in (1.0, 0.5):
            ds = cls(root, split="train", downscale=1, white_bkgd=True, scene_scale=scale)
            torch.manual_seed(77)
            out = ds.sample_random_rays(257, "cpu")
            img = ds.get_image_rays(3, "cpu")
            arrs = dict(rays_o=out[0], rays_d=out[1], target=out[2], img_rays_o=img[0], img_rays_d=img[1], img_target=img[2],
                        focal=np.float64(ds.focal), H=np.int64(ds.H), W=np.int64(ds.W))
            if tag == "dynamic":
                arrs.update(times=out[3], img_time=img[3], all_times=ds.times)
            save(f"dataset_{tag}_s{scale}", **arrs)


if __name__ == "__main__":
    if len(sys.argv) > 1:                      # regenerate selected groups only: make_golden.py hash_kat grid_update ...
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_dataset()
    gen_sampling()
    gen_mask()
    gen_composite()
    gen_pe()
    gen_hash_kat()
    gen_fields()
    gen_render()
    gen_grid_update()
