#!/usr/bin/env python
"""Drop-in check: run the REFERENCE's unchanged run.py + YAML config on top of this repo's `src`
package (project-nerf_b200/ first on sys.path), on a small synthetic Blender-format scene.

    python tools/dropin_check.py --mode part2_instant [--iters 300]

run.py and configs/ are looked up in $B2N_REFERENCE (default /root/reference, else baseline/_ref --
a git-ignored scratch copy that travels to the GPU box; reference sources are never committed).
"""
import argparse
import json
import math
import os
import runpy
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "project-nerf_b200")

CONFIGS = {"part2": "part2.yaml.example", "part2_instant": "part2_instant.yaml.example", "part3": "part3.yaml.example",
           "part3_instant": "part3_instant.yaml.example", "part4": "part4.yaml.example", "part3_dtc": "part3_dtc.yaml.example"}


def find_reference():
    for d in (os.environ.get("B2N_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.exists(os.path.join(d, "run.py")):
            return d
    raise SystemExit("reference run.py not found (set B2N_REFERENCE or copy run.py + configs/ to baseline/_ref/)")


def make_scene(root, res=64, n_train=24, n_test=8, dynamic=False):
    """an opaque shaded sphere (moving along x when dynamic) seen from cameras on a ring; RGBA PNGs"""
    from PIL import Image
    fov = 0.6911112070083618
    focal = 0.5 * res / math.tan(0.5 * fov)
    os.makedirs(root, exist_ok=True)
    rng = np.random.RandomState(0)
    for split, n in (("train", n_train), ("val", max(n_test // 2, 2)), ("test", n_test)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(n):
            az, el = rng.uniform(0, 2 * math.pi), math.radians(rng.uniform(5, 50))
            c = 4.0311 * np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
            fwd = -c / np.linalg.norm(c)
            right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right)
            up = np.cross(right, fwd)
            c2w = np.eye(4); c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, c
            t = i / max(n - 1, 1)
            centre = np.array([0.4 * math.sin(2 * math.pi * t), 0, 0]) if dynamic else np.zeros(3)
            jj, ii = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")
            d = np.stack([(ii - res * 0.5) / focal, -(jj - res * 0.5) / focal, -np.ones_like(ii, dtype=float)], -1)
            d = d @ c2w[:3, :3].T
            d /= np.linalg.norm(d, axis=-1, keepdims=True)
            oc = c - centre
            b = d @ oc
            disc = b * b - (oc @ oc - 0.6 ** 2)
            hit = disc > 0
            tt = -b - np.sqrt(np.maximum(disc, 0))
            nrm = (oc + d * tt[..., None]) / 0.6
            img = np.zeros((res, res, 4), dtype=np.uint8)
            img[..., 0] = np.clip(255 * (0.5 + 0.5 * nrm[..., 2]), 0, 255) * hit
            img[..., 1] = 60 * hit
            img[..., 2] = np.clip(255 * (0.5 - 0.5 * nrm[..., 2]), 0, 255) * hit
            img[..., 3] = 255 * hit
            Image.fromarray(img, "RGBA").save(os.path.join(root, split, f"r_{i:03d}.png"))
            fr = {"file_path": f"./{split}/r_{i:03d}", "transform_matrix": c2w.tolist()}
            if dynamic:
                fr["time"] = t
            frames.append(fr)
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump({"camera_angle_x": fov, "frames": frames}, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="part2_instant", choices=sorted(CONFIGS))
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--out", default="/tmp/b2n_dropin")
    args = ap.parse_args()
    ref = find_reference()
    import yaml
    cfg = yaml.safe_load(open(os.path.join(ref, "configs", CONFIGS[args.mode])))
    dynamic = cfg["mode"] in ("part3", "part4")
    data = os.path.join(args.out, "data", "sphere_dyn" if dynamic else "sphere")
    make_scene(data, dynamic=dynamic)
    cfg.update(train_iters=args.iters, batch_size=4096, log_every=max(args.iters // 6, 1), save_every=10 ** 9,
               val_every=args.iters, downscale=1, log_dir=os.path.join(args.out, "out", args.mode), chunk=4096,
               grid_warmup_iters=min(cfg.get("grid_warmup_iters", 256), args.iters // 3), n_samples=64, render_n_samples=64,
               grid_resolution=64)
    cfg_path = os.path.join(args.out, f"{args.mode}.yaml")
    os.makedirs(args.out, exist_ok=True)
    yaml.safe_dump(cfg, open(cfg_path, "w"))

    # run.py imports matplotlib only to write PNGs: a 10-line stand-in is enough
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            from PIL import Image
            plt = types.ModuleType("matplotlib.pyplot")
            plt.imsave = lambda path, arr, **kw: Image.fromarray((np.clip(np.asarray(arr), 0, 1) * 255).astype(np.uint8)).save(path)
            plt.close = lambda *a, **k: None
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = plt
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt

    sys.path[:0] = [PKG, ROOT]                    # OUR src package shadows the reference's
    import src.core
    assert os.path.realpath(src.core.__file__).startswith(os.path.realpath(PKG)), "reference src would be imported"
    sys.argv = [os.path.join(ref, "run.py"), "--config", cfg_path, "--data_dir", data]
    print(f">>> drop-in check: {ref}/run.py with {CONFIGS[args.mode]} on {data} using {src.core.__file__}")
    runpy.run_path(os.path.join(ref, "run.py"), run_name="__main__")
    print(">>> drop-in check finished without error")


if __name__ == "__main__":
    main()
