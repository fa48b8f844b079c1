"""FusedAdamW: the per-step parameter update of the training loops as two kernel launches.

The reference's loops (run.py:611-630, :1112-1120 + :1167-1178, :1840-1859 + :1940-1949) spend ~100 torch launches per
step on: the TV loss over the flat hash tables (forward + backward through slice / sub / abs / mean), GradScaler.unscale_,
clip_grad_norm_ and foreach-AdamW.  This optimizer does the same arithmetic in b2n_opt_prepare + b2n_opt_adamw:

    opt = b2n.optim.FusedAdamW([
        {"params": table_params, "tv_weight": 1e-6},           # + d/dp [ tv_weight * mean |p[1:] - p[:-1]| ] per tensor
        {"params": mlp_params}], lr=1e-2, weight_decay=1e-6)
    ...
    loss.backward()                    # WITHOUT the TV term in the loss
    opt.step(max_norm=1.0)             # == unscale_ + clip_grad_norm_(all params, 1.0) + AdamW.step
    # or, under AMP:  scaler.step(opt, max_norm=1.0); scaler.update()     (no unscale_ / clip_grad_norm_ calls)

It is a ``torch.optim.Optimizer`` (param groups, ``state_dict``, LR schedulers work); per-group options ``tv_weight`` and
``max_norm`` (clip that group on its own, like run.py:622-625 does for representation / decoder).  GradScaler hands it
``grad_scale`` / ``found_inf`` (``_step_supports_amp_scaling``): gradients are unscaled inside pass 1 and the update is
skipped on the device when an inf was found -- no host synchronisation anywhere.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import torch

from . import _lib
from ._lib import call, ptr, stream

MAX_TENSORS, MAX_GROUPS = 40, 8


class _OptTensor(ctypes.Structure):          # b2n_opt_tensor of include/b2nerf.h
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("n", ctypes.c_int64), ("lr", ctypes.c_float), ("weight_decay", ctypes.c_float),
                ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float),
                ("bias_corr1", ctypes.c_float), ("bias_corr2", ctypes.c_float), ("tv_scale", ctypes.c_float),
                ("clip_group", ctypes.c_int), ("reserved", ctypes.c_int)]


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, tv_weight=0.0, max_norm=None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, tv_weight=tv_weight,
                                      max_norm=max_norm))
        self._step_supports_amp_scaling = True      # GradScaler.step sets .grad_scale / .found_inf and calls step()
        self.last_grad_norm2 = None                 # device float[MAX_GROUPS]: squared norms of the last clipped step

    @torch.no_grad()
    def step(self, closure=None, max_norm: Optional[float] = None):
        """``max_norm``: one norm over ALL parameters (clip_grad_norm_(model.parameters(), max_norm)); otherwise every
        param group with a ``max_norm`` option is clipped on its own."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        descs, keep, limits, dev, step_t = [], [], [-1.0] * MAX_GROUPS, None, None
        next_group = 1 if max_norm is not None else 0
        if max_norm is not None:
            limits[0] = float(max_norm)
        for group in self.param_groups:
            b1, b2 = group["betas"]
            cg = -1
            if max_norm is not None:
                cg = 0
            elif group.get("max_norm") is not None:
                if next_group >= MAX_GROUPS:
                    raise ValueError(f"at most {MAX_GROUPS} separately clipped groups")
                cg, next_group = next_group, next_group + 1
                limits[cg] = float(group["max_norm"])
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise ValueError("FusedAdamW needs contiguous fp32 CUDA parameters: there is no CPU path in this package")
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32:
                    raise ValueError("FusedAdamW needs dense fp32 gradients")
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                # ONE device-side step counter shared by all parameters (state["step"] of each refers to it): a step
                # that GradScaler's inf flag cancels must not advance the bias corrections, and only the device knows
                if step_t is None:
                    step_t = st.get("step")
                    if step_t is None or not step_t.is_cuda:
                        step_t = torch.zeros(1, device=p.device) + (float(step_t) if step_t is not None else 0.0)
                st["step"] = step_t
                n = p.numel()
                tv = float(group.get("tv_weight", 0.0) or 0.0)
                dev = p.device
                keep.append(g)
                # the kernels write the parameter through its raw pointer: tell autograd (version counter), so that
                # anything keyed on "has this parameter changed" (DensityGrid's sweep cache, saved-tensor checks) sees it
                torch.autograd.graph.increment_version(p)
                descs.append(_OptTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                        n, float(group["lr"]), float(group["weight_decay"]), float(b1), float(b2),
                                        float(group["eps"]), 1.0, 1.0,
                                        tv / (n - 1) if (tv != 0.0 and n > 1) else 0.0, cg, 0))
        if not descs:
            return loss
        clipping = any(l >= 0 for l in limits)
        norm2 = None
        need_prepare = clipping or any(d.tv_scale != 0.0 for d in descs)
        gs_ptr = ptr(grad_scale) if grad_scale is not None else None
        fi_ptr = ptr(found_inf) if found_inf is not None else None
        lim = (ctypes.c_float * MAX_GROUPS)(*limits)
        with torch.cuda.device(dev):
            step_t.add_(1.0 if found_inf is None else (1.0 - found_inf.reshape(-1)[:1].to(step_t.dtype)))
            if need_prepare:
                norm2 = torch.zeros(MAX_GROUPS, device=dev)
                for i in range(0, len(descs), MAX_TENSORS):
                    chunk = descs[i:i + MAX_TENSORS]
                    arr = (_OptTensor * len(chunk))(*chunk)
                    call("b2n_opt_prepare", arr, len(chunk), gs_ptr, ptr(norm2), stream())
                self.last_grad_norm2 = norm2
            for i in range(0, len(descs), MAX_TENSORS):
                chunk = descs[i:i + MAX_TENSORS]
                arr = (_OptTensor * len(chunk))(*chunk)
                call("b2n_opt_adamw", arr, len(chunk), gs_ptr, int(need_prepare), fi_ptr, ptr(step_t),
                     ptr(norm2) if clipping else None, lim if clipping else None, stream())
        return loss
