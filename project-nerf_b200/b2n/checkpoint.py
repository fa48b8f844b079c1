"""Loading checkpoints written by the reference's drivers into this package's modules.

The reference saves ``{"model_state_dict": model.state_dict(), "config": cfg[, "step", "val_psnr", "density_grid"]}``
(run.py:212-215, :351-357, :707-715, :1325-1333, :2084-2092) and restores with ``load_state_dict``
(run.py:88-91, :296-297, :530-536, :1003-1006, :1664-1669).  The state_dict keys of this package's ``src`` modules are the
reference's, so a checkpoint of the same arithmetic loads with ``model.load_state_dict``.  A checkpoint trained on the
upstream tiny-cuda-nn build can differ in four ways this loader absorbs:

* dtype: fp16 ``params`` tensors (a half-precision export) -> fp32;
* the driver's wrapper dict, the ``deformation_grid`` alias of ``deform_grid_start`` (src/core.py:199) saved twice,
  ``module.`` prefixes;
* the resolution of a dense hash level: ``res = ceil(scale) + 1`` is computed in fp32 with libm, and level 3 of the
  stock geometry lands on 54 or 55 depending on the build.  A table whose total size reveals the other
  resolution is re-indexed level by level (entries keep their lattice coordinates);
* the content of the PADDED input columns of a FullyFusedMLP (``pad_value``): this package and its oracle pad with
  zeros, which makes the padded weight columns dead; an upstream build that pads with ones turns them into a learned
  first-layer bias.  Which one upstream does cannot be verified offline (no tiny-cuda-nn source), so it is a
  switch: ``pad_value=1.0`` keeps those columns alive (the kernels then feed ones into the padded inputs).

Nothing here touches the GPU; the loader works on CPU tensors and modules.
"""
from __future__ import annotations

import itertools
from typing import Dict, Optional, Union

import torch

from .ops import HashGeometry


def _unwrap(ckpt) -> Dict[str, torch.Tensor]:
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    if isinstance(ckpt, dict) and "model_state_dict" in ckpt:
        return ckpt, ckpt["model_state_dict"]
    return {}, ckpt


def _alt_levels(geom: HashGeometry, n_params: int):
    """per-level (res, size) of a table with ``n_params`` values whose dense levels may sit one lattice step off the
    resolutions of ``geom`` (libm-dependent ceil); None when no such layout has that size"""
    F = geom.n_features
    base = [(res, size, hashed) for (_, res, size, _, hashed) in geom.levels]
    cap = max(size for _, size, hashed in base if hashed) if any(h for _, _, h in base) else None
    dense = [i for i, (_, _, hashed) in enumerate(base) if not hashed]
    for deltas in itertools.product((0, 1, -1), repeat=len(dense)):
        if not any(deltas):
            continue
        cand = list(base)
        for i, d in zip(dense, deltas):
            res = base[i][0] + d
            size = (res ** 3 + 7) // 8 * 8
            if cap is not None and size > cap:
                size = cap            # would have been hashed upstream: not a re-indexable case
                cand = None
                break
            cand[i] = (res, size, False)
        if cand is not None and sum(size for _, size, _ in cand) * F == n_params:
            return cand
    return None


def convert_hash_table(params: torch.Tensor, geom: HashGeometry) -> torch.Tensor:
    """a flat hash-grid parameter vector in the tcnn layout (level-major, entry-major, feature-minor) -> the layout of
    ``geom``.  Identical sizes pass through; a dense level stored at the neighbouring resolution is re-indexed by lattice
    coordinate (x fastest): entries present in both lattices are copied, the rest keep the zero initialisation."""
    params = params.detach().to(torch.float32).reshape(-1)
    if params.numel() == geom.n_params:
        return params.clone()
    alt = _alt_levels(geom, params.numel())
    if alt is None:
        raise ValueError(f"hash table has {params.numel()} values, the geometry needs {geom.n_params}, and no layout with "
                         f"dense levels one lattice step off matches: different n_levels / log2_hashmap_size / "
                         f"base_resolution / per_level_scale?")
    F = geom.n_features
    out = torch.zeros(geom.n_params)
    src_off = 0
    for (res_s, size_s, hashed), (_, res_d, size_d, off_d, _) in zip(alt, geom.levels):
        src = params[src_off * F:(src_off + size_s) * F].view(size_s, F)
        dst = out[off_d * F:(off_d + size_d) * F].view(size_d, F)
        if hashed or res_s == res_d:
            dst.copy_(src)
        else:
            r = min(res_s, res_d)
            ax = torch.arange(r)
            z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
            i_s = (x + y * res_s + z * res_s * res_s).reshape(-1) % size_s
            i_d = (x + y * res_d + z * res_d * res_d).reshape(-1) % size_d
            dst[i_d] = src[i_s]
        src_off += size_s
    return out


def convert_fused_mlp(params: torch.Tensor, mlp) -> torch.Tensor:
    """a FullyFusedMLP parameter vector -> the padded row-major [out, in] layout of ``mlp`` (src.decoders.FusedMLP):
    fp16 -> fp32; an UNPADDED export (matrices of the logical in / out widths, row-major [out, in]) is zero-padded."""
    params = params.detach().to(torch.float32).reshape(-1)
    shapes = mlp.shapes
    n = sum(r * c for r, c in shapes)
    if params.numel() == n:
        return params.clone()
    logical = [(mlp.n_neurons, mlp.n_input_dims)] + [(mlp.n_neurons, mlp.n_neurons)] * (mlp.n_hidden - 1)
    logical.append((mlp.n_output_dims, mlp.n_neurons))
    if params.numel() == sum(r * c for r, c in logical):
        chunks, off = [], 0
        for (r, c), (rp, cp) in zip(logical, shapes):
            W = torch.zeros(rp, cp)
            W[:r, :c] = params[off:off + r * c].view(r, c)
            chunks.append(W.reshape(-1))
            off += r * c
        return torch.cat(chunks)
    raise ValueError(f"FullyFusedMLP parameter vector has {params.numel()} values; the padded layout "
                     f"{[tuple(s) for s in shapes]} needs {n}, the unpadded one {sum(r * c for r, c in logical)}")


def load_reference_checkpoint(model: torch.nn.Module, ckpt: Union[str, dict], pad_value: float = 0.0,
                              density_grid: Optional[torch.nn.Module] = None, strict: bool = True) -> dict:
    """Load a checkpoint written by the reference's run.py (or a bare state_dict) into ``model`` (src.core.NeuralField).

    ``pad_value``: content of the padded FullyFusedMLP input columns the weights were trained with (module docstring);
    it is installed on every ``FusedMLP`` of the model.  ``density_grid``: a ``DensityGrid`` to restore from the
    checkpoint's ``density_grid`` entry (run.py:713-714, :1331-1332, :2090-2091).  Returns a report dict."""
    from src.decoders import FusedMLP
    from src.embeddings import HashGridEncoding
    meta, sd = _unwrap(ckpt)
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    own = model.state_dict()
    report = {"converted": [], "missing": [], "unexpected": [], "pad_value": float(pad_value)}
    tables = {n + ".params": m for n, m in model.named_modules() if isinstance(m, HashGridEncoding)}
    mlps = {n + ".params": m for n, m in model.named_modules() if isinstance(m, FusedMLP)}
    new = {}
    for key, cur in own.items():
        src = sd.get(key)
        if src is None and key.startswith("deformation_grid."):          # alias saved (or not) by the reference
            src = sd.get("deform_grid_start." + key[len("deformation_grid."):])
        if src is None and key.startswith("deform_grid_start."):
            src = sd.get("deformation_grid." + key[len("deform_grid_start."):])
        if src is None:
            report["missing"].append(key)
            continue
        if key in tables:
            val = convert_hash_table(src, tables[key].geometry)
        elif key in mlps:
            val = convert_fused_mlp(src, mlps[key])
        else:
            val = src.detach().to(cur.dtype) if src.is_floating_point() else src.detach()
        if val.shape != cur.shape:
            if val.numel() != cur.numel():
                raise ValueError(f"{key}: checkpoint shape {tuple(src.shape)} does not fit {tuple(cur.shape)}")
            val = val.reshape(cur.shape)
        if src.dtype != cur.dtype or src.numel() != cur.numel():
            report["converted"].append(key)
        new[key] = val
    report["unexpected"] = sorted(k for k in sd if k not in own)
    if strict and (report["missing"] or report["unexpected"]):
        raise KeyError(f"checkpoint / model mismatch: missing {report['missing']}, unexpected {report['unexpected']}")
    model.load_state_dict(new, strict=False)
    for m in mlps.values():
        m.input_pad_value = float(pad_value)
    if density_grid is not None and isinstance(meta.get("density_grid"), dict):
        g = meta["density_grid"]
        density_grid.grid = g["grid"].to(density_grid.grid.device, torch.float32)
        density_grid.binary_grid = g["binary_grid"].to(density_grid.binary_grid.device, torch.bool)
        report["density_grid"] = True
    for k in ("step", "val_psnr", "config"):
        if k in meta:
            report[k] = meta[k]
    return report
