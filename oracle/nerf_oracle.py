"""Functional torch-CPU restatement of the Project-NeRF ray-marching hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line it follows (paths relative to /root/reference).  The code
is a restatement, not a copy: the reference is built from nn.Modules, this is
a set of pure functions over a flat ``state_dict`` so that the very same
weights can be fed to the reference, to this oracle and to the CUDA path.

All arithmetic is done in the dtype of the inputs (fp32 for parity runs, fp64
for gradient-check runs).  Expressions whose results must be BIT-EXACT
(sample depths, sample positions, voxel indices, hash indices) keep the
reference's exact operation order; no fused multiply-adds are available to
torch eager on CPU, which is what the CUDA kernels reproduce with
``__fmul_rn`` / ``__fadd_rn``.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------
# A1  Fourier positional encoding            (src/embeddings.py:13-32)
# --------------------------------------------------------------------------

def fourier_bands(L: int, dtype=torch.float32) -> Tensor:
    """``freq_bands`` buffer, src/embeddings.py:15: 2 ** linspace(0, L-1, L)."""
    if L <= 0:
        return torch.empty(0, dtype=dtype)
    return (2.0 ** torch.linspace(0.0, L - 1, steps=L)).to(dtype)


def fourier_encode(x: Tensor, bands: Tensor) -> Tensor:
    """gamma(x) = [x, sin(x f0 pi), cos(x f0 pi), ...]  (src/embeddings.py:22-32).

    Evaluation order is ((x * f) * pi) with pi rounded to the working dtype;
    column order is identity block, then per frequency [sin(D), cos(D)].
    An empty ``bands`` is the identity (src/embeddings.py:23-25).
    """
    if bands.numel() == 0:
        return x
    cols = [x]
    for k in range(bands.numel()):
        arg = (x * bands[k].to(x.dtype)) * np.pi
        cols.append(torch.sin(arg))
        cols.append(torch.cos(arg))
    return torch.cat(cols, dim=-1)


def fourier_out_dim(d_in: int, L: int) -> int:
    """src/embeddings.py:17."""
    return d_in + 2 * d_in * L if L > 0 else d_in


# --------------------------------------------------------------------------
# A2  Multiresolution hash grid (tiny-cuda-nn ``HashGrid``) -- PARITY UNPINNED
#     call site src/embeddings.py:60-89; algorithm restated from the upstream
#     library (grid.h) as summarised in SURVEY.md 8a/A2.
# --------------------------------------------------------------------------

_PRIME_Y = 2654435761
_PRIME_Z = 805459861
_U32 = 0xFFFFFFFF


def _libm():
    name = ctypes.util.find_library("m") or "libm.so.6"
    lib = ctypes.CDLL(name)
    for fn in ("exp2f", "log2f", "ceilf"):
        getattr(lib, fn).restype = ctypes.c_float
        getattr(lib, fn).argtypes = [ctypes.c_float]
    return lib


@dataclass(frozen=True)
class HashLevel:
    scale: float      # fp32 value, stored as python float (exactly representable)
    res: int          # grid resolution of this level
    size: int         # number of table entries of this level
    offset: int       # first entry of this level in the flat table
    hashed: bool      # True -> spatial hash, False -> dense index


def hash_level_table(n_levels: int, base_resolution: int, per_level_scale: float,
                     log2_hashmap_size: int) -> List[HashLevel]:
    """Per-level geometry, computed in fp32 with the host libm like upstream
    does on the host: scale = exp2f(l * log2f(s)) * base - 1, res = ceil(scale)+1,
    size = min(round_up(res^3, 8), 2^log2T)."""
    lm = _libm()
    log2s = lm.log2f(ctypes.c_float(per_level_scale))
    out, offset = [], 0
    for lvl in range(n_levels):
        e = np.float32(np.float32(lvl) * np.float32(log2s))
        scale = np.float32(np.float32(lm.exp2f(ctypes.c_float(float(e)))) * np.float32(base_resolution)) - np.float32(1.0)
        scale = np.float32(scale)
        res = int(lm.ceilf(ctypes.c_float(float(scale)))) + 1
        dense = res ** 3
        cap = (1 << 32) // 2 - 1
        n = min(dense, cap)
        n = (n + 7) // 8 * 8
        size = min(n, 1 << log2_hashmap_size)
        out.append(HashLevel(float(scale), res, size, offset, hashed=(size < dense)))
        offset += size
    return out


def hash_table_entries(levels: Sequence[HashLevel]) -> int:
    return levels[-1].offset + levels[-1].size


def hash_corner_index(level: HashLevel, cx: Tensor, cy: Tensor, cz: Tensor) -> Tensor:
    """Entry index of integer lattice corner (cx,cy,cz) inside ``level``;
    uint32 arithmetic emulated in int64.  Dense: x + y*res + z*res^2; hashed:
    (x*1) ^ (y*2654435761) ^ (z*805459861); both modulo the level size."""
    if level.hashed:
        h = (cx & _U32) ^ ((cy * _PRIME_Y) & _U32) ^ ((cz * _PRIME_Z) & _U32)
    else:
        h = (cx + cy * level.res + cz * (level.res * level.res)) & _U32
    return h % level.size


def hash_encode(x01: Tensor, params: Tensor, levels: Sequence[HashLevel], n_feat: int) -> Tensor:
    """Trilinear multiresolution lookup.  x01 [P,3] in [0,1]; params flat
    [F * sum(size)] (level-major, entry-major, feature-minor) -> [P, L*F].
    Differentiable w.r.t. ``params`` (scatter-add) and ``x01`` (through the
    interpolation weights; floor() carries no gradient)."""
    table = params.view(-1, n_feat)
    feats = []
    for lv in levels:
        pos = x01 * torch.tensor(lv.scale, dtype=x01.dtype) + 0.5
        g = torch.floor(pos)
        w = pos - g
        gi = g.detach().to(torch.int64)
        acc = None
        for corner in range(8):
            wt = None
            idx3 = []
            for dim in range(3):
                bit = (corner >> dim) & 1
                wd = w[:, dim] if bit else (1.0 - w[:, dim])
                wt = wd if wt is None else wt * wd
                idx3.append(gi[:, dim] + bit)
            e = hash_corner_index(lv, idx3[0], idx3[1], idx3[2]) + lv.offset
            term = wt[:, None] * table[e]
            acc = term if acc is None else acc + term
        feats.append(acc)
    return torch.cat(feats, dim=-1)


def hash_representation(x: Tensor, params: Tensor, levels, n_feat: int, bound: float) -> Tensor:
    """World -> unit cube -> hash features (src/embeddings.py:75-89)."""
    x01 = ((x + bound) / (2 * bound)).clamp(0.0, 1.0)
    return hash_encode(x01, params, levels, n_feat)


# --------------------------------------------------------------------------
# A4/A6  Bias-free fused MLP (tiny-cuda-nn ``FullyFusedMLP``) -- PARITY UNPINNED
#     call sites src/decoders.py:111-134, :285-295.
# --------------------------------------------------------------------------

def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


def fused_mlp_shapes(n_in: int, n_out: int, n_neurons: int, n_hidden: int) -> List[Tuple[int, int]]:
    """[(rows=out, cols=in)] of each weight matrix inside the flat ``params``:
    in/out widths padded to multiples of 16, every matrix row-major [out,in]."""
    shapes = [(n_neurons, _pad16(n_in))]
    shapes += [(n_neurons, n_neurons)] * (n_hidden - 1)
    shapes.append((_pad16(n_out), n_neurons))
    return shapes


def fused_mlp_n_params(n_in, n_out, n_neurons, n_hidden) -> int:
    return sum(r * c for r, c in fused_mlp_shapes(n_in, n_out, n_neurons, n_hidden))


def bf16_round(x: Tensor) -> Tensor:
    """round-to-nearest-even to bfloat16 with a straight-through gradient (emulates the operand
    rounding of the tensor-core kernels; accumulation stays in the working dtype)."""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


# 16-bit operand type of each CUDA kernel family, for the "kernel" arithmetic model: the fused Instant decoder
# (csrc/b2n_mlp64.cu) and the fused deformation / time-modulation nets (csrc/b2n_fmlp.cu) compute in IEEE fp16 like
# tinycudann, the 256-wide tcgen05 decoder in bf16.
KERNEL_OPERAND = {"instant": torch.float16, "fmlp": torch.float16, "nerf256": torch.bfloat16}


def _q(x: Tensor, dt=torch.bfloat16) -> Tensor:
    return x.to(dt).to(x.dtype)


class _KernelMatmul(torch.autograd.Function):
    """h @ W^T with the operand rounding of the fused tensor-core kernels, forward AND backward
    (``emulate_bf16="kernel"``):

    forward : bf16(h) @ bf16(W)^T, fp32 accumulation -- or the exact product when ``exact_fwd``
              (the kernels evaluate the first layer of a hash-grid decoder as a split-bf16
              product hi*hi + lo*hi + hi*lo, i.e. to ~16 mantissa bits: csrc/b2n_mlp64.cu);
    backward: the incoming pre-activation gradient dZ is rounded to bf16 once (the kernels stage
              it in shared memory as bf16) and that rounded dZ feeds both the data gradient
              dZ bf16(W) and the weight gradient dZ^T bf16(h)."""

    @staticmethod
    def forward(ctx, h, W, exact_fwd, dt):
        hq, Wq = _q(h, dt), _q(W, dt)
        ctx.save_for_backward(hq, Wq)
        ctx.dt = dt
        return (h @ W.t()) if exact_fwd else (hq @ Wq.t())

    @staticmethod
    def backward(ctx, dz):
        hq, Wq = ctx.saved_tensors
        if ctx.dt == torch.float16:
            # the fp16 kernels run the gradient chain on S * dZ, S a power of two putting max|dL/dy| near 2^7: the
            # rounding is relative (no underflow) for everything within ~2^-21 of the largest gradient -- model it as
            # a pure 11-bit mantissa rounding by normalising with the tensor's own power-of-two scale
            m = float(dz.abs().max())
            sc = 2.0 ** (7 - math.floor(math.log2(m))) if (m > 0 and math.isfinite(m)) else 1.0
            dzq = _q(dz * sc, ctx.dt) / sc
        else:
            dzq = _q(dz, ctx.dt)
        return dzq @ Wq, dzq.t() @ hq, None, None


def _matmul_t(h: Tensor, W: Tensor, emulate_bf16, exact_fwd: bool = False, family: str = "fmlp") -> Tensor:
    """h @ W^T under the three arithmetic models of this oracle: False = working dtype; True = bf16 forward
    operands with a straight-through gradient; "kernel" = _KernelMatmul with the kernel family's operand type."""
    if emulate_bf16 == "kernel":
        return _KernelMatmul.apply(h, W, exact_fwd, KERNEL_OPERAND[family])
    if emulate_bf16:
        return bf16_round(h) @ bf16_round(W).t()
    return h @ W.t()


def fused_mlp(x: Tensor, params: Tensor, n_in: int, n_out: int, n_neurons: int,
              n_hidden: int, out_act: str = "None", emulate_bf16=False, first_exact: bool = False,
              family: str = "fmlp", pad_value: float = 0.0) -> Tensor:
    """ReLU MLP without bias terms.  Input columns beyond ``n_in`` are padded
    with ZEROS (SURVEY.md A4 fixes this choice), so the padded weight columns
    never contribute; the padded output rows are computed and sliced away.
    ``first_exact`` (kernel emulation only): the first layer's forward product is not rounded.
    ``pad_value``: what the padded input columns hold instead -- 1.0 models an upstream tiny-cuda-nn whose Network
    pads its inputs with ones (unverifiable offline; the padded weight columns then act as a bias)."""
    shapes = fused_mlp_shapes(n_in, n_out, n_neurons, n_hidden)
    h = x
    if shapes[0][1] != n_in:
        h = F.pad(h, (0, shapes[0][1] - n_in), value=float(pad_value))
    off = 0
    for li, (r, c) in enumerate(shapes):
        W = params[off:off + r * c].view(r, c).to(h.dtype)
        off += r * c
        h = _matmul_t(h, W, emulate_bf16, exact_fwd=first_exact and li == 0, family=family)
        if li < len(shapes) - 1:
            h = torch.relu(h)
    h = h[:, :n_out]
    if out_act == "Sigmoid":
        h = torch.sigmoid(h)
    elif out_act != "None":
        raise ValueError(out_act)
    return h


# --------------------------------------------------------------------------
# A3/A5/A6  torch.nn.Linear based decoders
# --------------------------------------------------------------------------

def _lin(sd: Dict[str, Tensor], prefix: str, h: Tensor) -> Tensor:
    return F.linear(h, sd[prefix + ".weight"].to(h.dtype), sd[prefix + ".bias"].to(h.dtype))


def nerf_decoder(sd, prefix: str, x: Tensor, d: Tensor, num_layers=8, skip_layer=4, emulate_bf16=False):
    """8x256 trunk with [h, x] skip concat, sigma/feature heads, view branch
    (src/decoders.py:68-87).  ``emulate_bf16`` rounds the operands of the trunk / feature / view
    GEMMs to bf16 like the tensor-core kernel does (the two small heads stay in fp32)."""
    def lin(name, v):
        return _matmul_t(v, sd[name + ".weight"].to(v.dtype), emulate_bf16, family="nerf256") + sd[name + ".bias"].to(v.dtype)
    h = x
    for i in range(num_layers):
        if i == skip_layer:
            h = torch.cat([h, x], dim=-1)
        h = torch.relu(lin(f"{prefix}.pts_layers.{i}", h))
    sigma = torch.relu(_lin(sd, f"{prefix}.sigma_layer", h))
    feat = lin(f"{prefix}.feature_layer", h)
    hv = torch.relu(lin(f"{prefix}.view_layer", torch.cat([feat, d], dim=-1)))
    rgb = torch.sigmoid(_lin(sd, f"{prefix}.rgb_layer", hv))
    return rgb, sigma


def instant_decoder(sd, prefix: str, x_enc: Tensor, d_enc: Tensor, hidden=64, emulate_bf16=False, pad_value: float = 0.0):
    """sigma_net (in->64->16), sigma = softplus(h0-5), color_net on
    cat[h(16), d_enc] (->64->64->3, sigmoid)   (src/decoders.py:136-162).
    Kernel emulation: sigma_net's first layer (the one fed by the hash features) is a split-bf16 product."""
    pos_dim, dir_dim = x_enc.shape[-1], d_enc.shape[-1]
    h = fused_mlp(x_enc, sd[f"{prefix}.sigma_net.params"], pos_dim, 16, hidden, 1, emulate_bf16=emulate_bf16,
                  first_exact=True, family="instant", pad_value=pad_value)
    sigma = F.softplus(h[..., 0:1] - 5.0)
    rgb = fused_mlp(torch.cat([h, d_enc], dim=-1), sd[f"{prefix}.color_net.params"],
                    16 + dir_dim, 3, hidden, 2, out_act="Sigmoid", emulate_bf16=emulate_bf16, family="instant",
                    pad_value=pad_value)
    return rgb, sigma


def _lin_q(sd: Dict[str, Tensor], prefix: str, h: Tensor, emulate_bf16: bool) -> Tensor:
    """nn.Linear with (optionally) bf16-rounded GEMM operands and an fp32 bias, like the tensor-core kernels"""
    if not emulate_bf16:
        return _lin(sd, prefix, h)
    return _matmul_t(h, sd[prefix + ".weight"].to(h.dtype), emulate_bf16) + sd[prefix + ".bias"].to(h.dtype)


def deformation_net(sd, prefix: str, x_feat: Tensor, t_feat: Tensor, num_layers=4, emulate_bf16=False):
    """Linear/ReLU stack -> 3-vector (src/decoders.py:171-195); nn.Sequential
    indices 0,2,4,... hold the Linear layers."""
    h = torch.cat([x_feat, t_feat], dim=-1)
    for i in range(num_layers):
        h = _lin_q(sd, f"{prefix}.net.{2 * i}", h, emulate_bf16)
        if i < num_layers - 1:
            h = torch.relu(h)
    return h


def time_modulation(sd, prefix: str, t_feat: Tensor, num_layers=2, emulate_bf16=False):
    """sigmoid(MLP(time features))   (src/decoders.py:342-371)."""
    h = t_feat
    for i in range(num_layers):
        h = _lin_q(sd, f"{prefix}.net.{2 * i}", h, emulate_bf16)
        if i < num_layers - 1:
            h = torch.relu(h)
    return torch.sigmoid(h)


def hash_deform_decoder(sd, prefix: str, hash_feat: Tensor, time_mod: Tensor, hidden=64, emulate_bf16=False,
                        pad_value: float = 0.0):
    """fused MLP (cat -> 64 -> 64 -> 3) * displacement_scale (src/decoders.py:300-318)."""
    h = torch.cat([hash_feat, time_mod], dim=-1)
    dx = fused_mlp(h, sd[f"{prefix}.deform_net.params"], h.shape[-1], 3, hidden, 2, emulate_bf16=emulate_bf16,
                   pad_value=pad_value)
    return dx * sd[f"{prefix}.displacement_scale"].to(dx.dtype)


# --------------------------------------------------------------------------
# A7  NeuralField graph wiring                 (src/core.py:227-363)
# --------------------------------------------------------------------------

class OracleField:
    """Pure-function twin of ``NeuralField`` driven by (config, state_dict).

    ``training`` + ``use_coord_noise`` reproduce the train-time input noise of
    src/core.py:254-262 / :289-294 by issuing the same ``torch.randn_like``
    calls in the same order (so a shared CPU seed gives identical noise).
    """

    def __init__(self, cfg: dict, sd: Dict[str, Tensor], emulate_bf16=False, pad_value: float = 0.0):
        self.cfg, self.sd = cfg, sd
        self.pad_value = pad_value             # content of the padded FullyFusedMLP input columns (see fused_mlp)
        self.mode = cfg["mode"]
        self.training = False
        # False | True | "kernel": arithmetic model of every decoder GEMM (see _matmul_t); the encoders, the blend and
        # the compositing stay in the working dtype, as they do in the CUDA path
        self.emulate_bf16 = emulate_bf16
        g = cfg.get
        self.use_coord_noise = g("use_coord_noise", False)
        self.coord_noise_std = g("coord_noise_std", 0.005)
        self.time_noise_std = g("time_noise_std", 0.02)
        self.hidden = g("hidden_dim", 64 if self._instant_canonical() else 256)
        if self.mode in ("part2_instant",) or (self.mode == "part3" and g("canonical_type", "nerf") == "instant"):
            bound_default = 1.0
        else:
            bound_default = 1.5
        self.bound = g("scene_bound", bound_default)
        if self._instant_canonical():
            self.levels = hash_level_table(g("n_levels", 16), g("base_resolution", 16),
                                           g("per_level_scale", 1.5), g("log2_hashmap_size", 19))
            self.n_feat = g("n_features_per_level", 2)
        if self.mode == "part4":
            self.d_levels = hash_level_table(g("deform_n_levels", 14), g("deform_base_resolution", 16),
                                             g("deform_per_level_scale", 1.5), g("deform_log2_hashmap_size", 19))
            self.d_feat = g("deform_n_features_per_level", 2)

    def _instant_canonical(self) -> bool:
        m = self.cfg["mode"]
        return m in ("part2_instant", "part4") or (m == "part3" and self.cfg.get("canonical_type", "nerf") == "instant")

    def train(self, flag=True):
        self.training = flag
        return self

    def eval(self):
        return self.train(False)

    # -- helpers ---------------------------------------------------------
    def _pe(self, name: str, x: Tensor) -> Tensor:
        return fourier_encode(x, self.sd[f"{name}.freq_bands"])

    def _hash(self, name: str, x: Tensor, levels, n_feat) -> Tensor:
        return hash_representation(x, self.sd[f"{name}.encoding.params"], levels, n_feat, self.bound)

    def _noise(self, x: Tensor, t: Tensor):
        xd, td = x, t
        if self.training and self.use_coord_noise:
            if self.coord_noise_std > 0:
                xd = x + torch.randn_like(x) * self.coord_noise_std
            if self.time_noise_std > 0:
                td = torch.clamp(t + torch.randn_like(t) * self.time_noise_std, 0.0, 1.0)
        return xd, td

    def _nerf_dec(self, prefix, h, d):
        g = self.cfg.get
        return nerf_decoder(self.sd, prefix, h, d, g("num_layers", 8), g("skip_layer", 4), emulate_bf16=self.emulate_bf16)

    # -- forward ----------------------------------------------------------
    def __call__(self, x: Tensor, d: Optional[Tensor] = None, t: Optional[Tensor] = None):
        m = self.mode
        if m == "part2_nerf":
            if d is None:
                raise ValueError("part2_nerf requires view directions.")
            return self._nerf_dec("decoder", self._pe("representation", x), self._pe("dir_representation", d))
        if m == "part2_instant":
            if d is None:
                raise ValueError("part2_instant requires view directions.")
            h = self._hash("representation", x, self.levels, self.n_feat)
            return instant_decoder(self.sd, "decoder", h, self._pe("dir_representation", d), self.hidden,
                                   emulate_bf16=self.emulate_bf16, pad_value=self.pad_value)
        if m == "part3":
            if t is None:
                raise ValueError("Part 3 requires time input 't'.")
            if self.cfg.get("direct_time_conditioning", False):     # src/core.py:237-247
                h = torch.cat([self._pe("pos_encoder_direct", x), self._pe("time_encoder", t)], dim=-1)
                rgb, sigma = self._nerf_dec("decoder_direct", h, self._pe("dir_representation", d))
                return rgb, sigma, torch.zeros_like(x)
            xd, td = self._noise(x, t)
            feat_t = self._pe("time_encoder", td)
            # (the CUDA path runs the deformation net on tensor cores only for hidden widths 64 / 128)
            emu_d = self.emulate_bf16 if self.cfg.get("deform_hidden_dim", 128) in (64, 128) else False
            dx = deformation_net(self.sd, "deform_net", self._pe("pos_encoder_for_deform", xd), feat_t,
                                 self.cfg.get("deform_num_layers", 4), emulate_bf16=emu_d)
            xc = x + dx                                             # un-noised x, src/core.py:268
            fd = self._pe("dir_representation", d)
            if self._instant_canonical():
                fc = self._hash("canonical_repr", xc, self.levels, self.n_feat)
                rgb, sigma = instant_decoder(self.sd, "decoder", torch.cat([fc, feat_t], dim=-1), fd, self.hidden,
                                             emulate_bf16=self.emulate_bf16, pad_value=self.pad_value)
            else:
                fc = self._pe("canonical_repr", xc)
                rgb, sigma = self._nerf_dec("decoder", torch.cat([fc, feat_t], dim=-1), fd)
            return rgb, sigma, dx
        if m == "part4":
            if t is None:
                raise ValueError("Part 4 requires time input 't'.")
            xd, td = self._noise(x, t)
            feat_t = self._pe("time_encoder", td)
            tmod = time_modulation(self.sd, "time_modulation", feat_t, self.cfg.get("time_modulation_layers", 2),
                                   emulate_bf16=self.emulate_bf16)
            f0 = self._hash("deform_grid_start", xd, self.d_levels, self.d_feat)
            f1 = self._hash("deform_grid_mid", xd, self.d_levels, self.d_feat)
            f2 = self._hash("deform_grid_end", xd, self.d_levels, self.d_feat)
            w0 = torch.clamp(1.0 - torch.abs(td - 0.0) / 0.5, 0.0, 1.0)   # src/core.py:324-332
            w1 = torch.clamp(1.0 - torch.abs(td - 0.5) / 0.5, 0.0, 1.0)
            w2 = torch.clamp(1.0 - torch.abs(td - 1.0) / 0.5, 0.0, 1.0)
            ws = w0 + w1 + w2 + 1e-8
            blend = (w0 / ws) * f0 + (w1 / ws) * f1 + (w2 / ws) * f2
            dx = hash_deform_decoder(self.sd, "deform_decoder", blend, tmod, self.cfg.get("deform_hidden_dim", 64),
                                     emulate_bf16=self.emulate_bf16, pad_value=self.pad_value)
            xc = x + dx
            fc = self._hash("canonical_repr", xc, self.levels, self.n_feat)
            rgb, sigma = instant_decoder(self.sd, "decoder", torch.cat([fc, feat_t], dim=-1),
                                         self._pe("dir_representation", d), self.hidden, emulate_bf16=self.emulate_bf16,
                                         pad_value=self.pad_value)
            return rgb, sigma, dx
        raise ValueError(f"mode {m!r} is outside the ray-marching hot path")


# --------------------------------------------------------------------------
# A8  stratified sampling                       (src/renderer.py:186-201)
# --------------------------------------------------------------------------

def sample_stratified(near: float, far: float, n_samples: int, n_rays: int,
                      u: Optional[Tensor] = None, dtype=torch.float32) -> Tensor:
    """Depths z [B,N].  ``u`` is the U[0,1) jitter the reference draws with
    ``torch.rand`` (renderer.py:198); None = unperturbed."""
    t = torch.linspace(0.0, 1.0, steps=n_samples, dtype=dtype)
    z = (near * (1.0 - t) + far * t).expand(n_rays, n_samples)
    if u is not None:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        hi = torch.cat([mid, z[:, -1:]], dim=-1)
        lo = torch.cat([z[:, :1], mid], dim=-1)
        z = lo + (hi - lo) * u
    return z


# --------------------------------------------------------------------------
# A11  occupancy lookup                         (src/renderer.py:134-166)
# --------------------------------------------------------------------------

def active_mask(pts: Tensor, binary_grid: Tensor, bound: float) -> Tensor:
    """voxel = trunc((p + bound) * (R / (2 bound))) (toward zero, so points up
    to one voxel outside the negative faces land in voxel 0 and are valid);
    x is the slowest axis of the [R,R,R] grid."""
    R = binary_grid.shape[0]
    idx = ((pts + bound) * (R / (2 * bound))).long()
    ok = ((idx >= 0) & (idx < R)).all(dim=-1)
    flat = (idx[:, 0] * R + idx[:, 1]) * R + idx[:, 2]
    flat = torch.where(ok, flat, torch.zeros_like(flat))
    return ok & binary_grid.reshape(-1)[flat]


# --------------------------------------------------------------------------
# A9  alpha compositing                         (src/renderer.py:204-237)
# --------------------------------------------------------------------------

def composite_weights(sigma: Tensor, z: Tensor, rays_d: Tensor) -> Tensor:
    """w_i = alpha_i * prod_{j<i}(1 - alpha_j + 1e-10); last interval 1e10;
    intervals scaled by |d|   (renderer.py:213-223)."""
    diff = z[:, 1:] - z[:, :-1]
    # N == 1 quirk kept from the reference: full_like of the EMPTY slice diff[:, :1] is empty, so every
    # tensor below is [B,0] (sigma broadcasts against it) and the ray composites to pure background.
    delta = torch.cat([diff, torch.full_like(diff[:, :1], 1e10)], dim=-1)
    delta = delta * torch.norm(rays_d[:, None, :], dim=-1)
    alpha = 1.0 - torch.exp(-sigma * delta)
    keep = torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10], dim=-1)
    trans = torch.cumprod(keep, dim=-1)[:, :-1]
    return alpha * trans


def volume_render(rgb: Tensor, sigma: Tensor, z: Tensor, rays_d: Tensor, bg: Optional[Tensor] = None):
    w = composite_weights(sigma, z, rays_d)
    if w.shape[1] == 0:                      # N == 1 (see composite_weights)
        rgb, z = rgb[:, :0], z[:, :0]
    color = (w[..., None] * rgb).sum(dim=-2)
    depth = (w * z).sum(dim=-1)
    acc = w.sum(dim=-1)
    if bg is not None:
        color = color + (1.0 - acc)[..., None] * (bg[None] if bg.dim() == 1 else bg)
    return color, depth, acc


# --------------------------------------------------------------------------
# A10  render_rays                              (src/renderer.py:240-384)
# --------------------------------------------------------------------------

def render_rays(field: Callable, rays_o: Tensor, rays_d: Tensor, near: float, far: float,
                n_samples: int, u: Optional[Tensor] = None, binary_grid: Optional[Tensor] = None,
                grid_bound: float = 1.0, times: Optional[Tensor] = None,
                white_bkgd: bool = True, bg_color: Optional[Tensor] = None):
    """Sample -> (occupancy mask -> compact) -> field -> scatter -> composite.
    Returns (rgb, depth, acc) or (rgb, depth, acc, extras) when ``times`` is
    given, like the reference.  ``binary_grid`` plays ``density_grid``."""
    B, N, dt = rays_o.shape[0], n_samples, rays_o.dtype
    mode = getattr(field, "mode", "unknown")
    dynamic = mode in ("part3", "part4")
    if bg_color is None:
        bg_color = torch.ones(3, dtype=dt) if white_bkgd else torch.zeros(3, dtype=dt)
    t_flat = None
    if dynamic:
        tt = times if times is not None else torch.zeros((B, 1), dtype=dt)
        t_flat = tt.expand(-1, N).reshape(-1, 1)
    z = sample_stratified(near, far, N, B, u, dtype=dt)
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]).reshape(-1, 3)
    vd = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    vd = vd[:, None, :].expand(-1, N, -1).reshape(-1, 3)

    dx_flat = None
    if binary_grid is not None:
        m = active_mask(pts, binary_grid, grid_bound)
        if not bool(m.any()):                    # renderer.py:309-311
            m = m.clone()
            m[0] = True
        if dynamic:
            rgb_c, sig_c, dx_c = field(pts[m], vd[m], t=t_flat[m])
        else:
            rgb_c, sig_c = field(pts[m], vd[m])
            dx_c = None
        rgb = torch.zeros(B * N, 3, dtype=dt).index_put((m,), rgb_c.to(dt))   # renderer.py:328-338
        sig = torch.zeros(B * N, 1, dtype=dt).index_put((m,), sig_c.to(dt))
        if dx_c is not None:
            dx_flat = torch.zeros(B * N, 3, dtype=dt).index_put((m,), dx_c.to(dt))
    else:
        if dynamic:
            rgb, sig, dx_flat = field(pts, vd, t=t_flat)
        else:
            rgb, sig = field(pts, vd)
    rgb, sig = rgb.to(dt).view(B, N, 3), sig.to(dt).view(B, N)
    color, depth, acc = volume_render(rgb, sig, z, rays_d, bg_color)
    if times is None:
        return color, depth, acc
    extras = {}
    if dynamic and dx_flat is not None:            # renderer.py:363-380
        w = composite_weights(sig, z, rays_d)
        extras["mean_delta_x"] = (w[..., None] * dx_flat.view(B, N, 3)).sum(dim=1)
    return color, depth, acc, extras


# --------------------------------------------------------------------------
# A12  occupancy-grid maintenance               (src/renderer.py:36-132)
# --------------------------------------------------------------------------

def grid_corner_points(R: int, bound: float, dtype=torch.float32) -> Tensor:
    """R^3 lattice of linspace(-b, b, R) corners, 'ij' order, x slowest (renderer.py:49-54)."""
    ax = torch.linspace(-bound, bound, R, dtype=dtype)
    xx, yy, zz = torch.meshgrid(ax, ax, ax, indexing="ij")
    return torch.stack([xx, yy, zz], dim=-1).reshape(-1, 3)


@torch.no_grad()
def density_grid_update(field: Callable, grid: Tensor, bound: float, threshold: float,
                        time: Optional[Tensor] = None, decay: float = 1.0, batch: int = 2 ** 18):
    """Returns (new_grid, new_binary, active_ratio)."""
    R = grid.shape[0]
    pts = grid_corner_points(R, bound, grid.dtype)
    mode = getattr(field, "mode", "unknown")

    def sweep(tval: Optional[Tensor]):
        out = []
        for i in range(0, pts.shape[0], batch):
            p = pts[i:i + batch]
            if tval is None:
                _, s = field(p, torch.zeros_like(p))
            else:
                _, s, _ = field(p, torch.zeros_like(p), t=tval.expand(p.shape[0], -1))
            out.append(s.squeeze(-1))
        return torch.cat(out)

    if mode == "part4":                            # three anchors, max (renderer.py:65-86)
        cur = torch.stack([sweep(torch.tensor([[a]], dtype=grid.dtype)) for a in (0.0, 0.5, 1.0)]).max(dim=0)[0]
    elif mode == "part3":
        if time is None:
            raise ValueError("Part 3 density grid update requires a time parameter")
        cur = sweep(time)
    else:
        cur = sweep(None)
    cur = cur.reshape(R, R, R)
    new = torch.maximum(grid * decay, cur) if mode in ("part3", "part4") else cur
    binary = new > threshold
    return new, binary, binary.float().mean().item()


# --------------------------------------------------------------------------
# parameter construction for synthetic runs (random init of the right shapes)
# --------------------------------------------------------------------------

def _linear_init(out_f: int, in_f: int, gen: torch.Generator) -> Tuple[Tensor, Tensor]:
    """nn.Linear default init (kaiming_uniform a=sqrt(5) == U(+-1/sqrt(in)))."""
    k = 1.0 / math.sqrt(in_f)
    W = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * k
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * k
    return W, b


def _fused_init(n_in, n_out, n_neurons, n_hidden, gen) -> Tensor:
    """Xavier-uniform per matrix, like upstream's default initialiser."""
    chunks = []
    for r, c in fused_mlp_shapes(n_in, n_out, n_neurons, n_hidden):
        lim = math.sqrt(6.0 / (r + c))
        chunks.append(((torch.rand(r, c, generator=gen) * 2 - 1) * lim).reshape(-1))
    return torch.cat(chunks)


def _hash_init(levels, n_feat, gen) -> Tensor:
    return (torch.rand(hash_table_entries(levels) * n_feat, generator=gen) * 2 - 1) * 1e-4


def make_state_dict(cfg: dict, seed: int = 0, table_scale: float = 1.0) -> Dict[str, Tensor]:
    """Random parameters with the reference's state_dict key names and shapes
    (src/core.py:11-225).  ``table_scale`` > 1 inflates the U(+-1e-4) hash
    init so that features are not numerically negligible in parity tests."""
    g = cfg.get
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    mode = cfg["mode"]

    def pe(name, L):
        sd[f"{name}.freq_bands"] = fourier_bands(L)

    def lin(name, o, i):
        sd[f"{name}.weight"], sd[f"{name}.bias"] = _linear_init(o, i, gen)

    def nerf_dec(prefix, pos_dim, dir_dim):
        H, nl, sk, vd = g("hidden_dim", 256), g("num_layers", 8), g("skip_layer", 4), g("view_dim", 128)
        for i in range(nl):
            ind = pos_dim if i == 0 else H
            if i == sk:
                ind += pos_dim
            lin(f"{prefix}.pts_layers.{i}", H, ind)
        lin(f"{prefix}.sigma_layer", 1, H)
        lin(f"{prefix}.feature_layer", H, H)
        lin(f"{prefix}.view_layer", vd, H + dir_dim)
        lin(f"{prefix}.rgb_layer", 3, vd)

    def inst_dec(prefix, pos_dim, dir_dim):
        H = g("hidden_dim", 64)
        sd[f"{prefix}.sigma_net.params"] = _fused_init(pos_dim, 16, H, 1, gen)
        sd[f"{prefix}.color_net.params"] = _fused_init(16 + dir_dim, 3, H, 2, gen)

    def hashgrid(name, levels, nf):
        sd[f"{name}.encoding.params"] = _hash_init(levels, nf, gen) * table_scale

    L_dir = g("L_embed_dir", 4)
    dir_dim = fourier_out_dim(3, L_dir)
    if mode == "part2_nerf":
        L = g("L_embed", 0) if g("use_positional_encoding", True) else 0
        L_dir = L_dir if g("use_viewdirs", True) else 0
        pe("representation", L)
        pe("dir_representation", L_dir)
        nerf_dec("decoder", fourier_out_dim(3, L), fourier_out_dim(3, L_dir))
        return sd
    levels = None
    if mode in ("part2_instant", "part4") or (mode == "part3" and g("canonical_type", "nerf") == "instant"):
        levels = hash_level_table(g("n_levels", 16), g("base_resolution", 16), g("per_level_scale", 1.5),
                                  g("log2_hashmap_size", 19))
        nf = g("n_features_per_level", 2)
    if mode == "part2_instant":
        hashgrid("representation", levels, nf)
        pe("dir_representation", L_dir)
        inst_dec("decoder", len(levels) * nf, dir_dim)
        return sd
    L_time = g("L_embed_time", 10)
    time_dim = fourier_out_dim(1, L_time)
    pe("dir_representation", L_dir)
    pe("time_encoder", L_time)
    if mode == "part3":
        Lp = g("L_embed", 10)
        pe("pos_encoder_for_deform", Lp)
        Hd, nd = g("deform_hidden_dim", 128), g("deform_num_layers", 4)
        ind = fourier_out_dim(3, Lp) + time_dim
        for i in range(nd):
            lin(f"deform_net.net.{2 * i}", 3 if i == nd - 1 else Hd, ind if i == 0 else Hd)
        last = f"deform_net.net.{2 * (nd - 1)}"
        sd[last + ".weight"] = (torch.rand(3, Hd if nd > 1 else ind, generator=gen) * 2 - 1) * 1e-4
        sd[last + ".bias"] = torch.zeros(3)
        if levels is not None:
            hashgrid("canonical_repr", levels, nf)
            inst_dec("decoder", len(levels) * nf + time_dim, dir_dim)
        else:
            Lc = g("L_embed_canon", 10)
            pe("canonical_repr", Lc)
            nerf_dec("decoder", fourier_out_dim(3, Lc) + time_dim, dir_dim)
        if g("direct_time_conditioning", False):
            pe("pos_encoder_direct", Lp)
            nerf_dec("decoder_direct", fourier_out_dim(3, Lp) + time_dim, dir_dim)
        return sd
    if mode == "part4":
        tm, tl = g("time_modulation_dim", 64), g("time_modulation_layers", 2)
        ind = time_dim
        for i in range(tl):
            lin(f"time_modulation.net.{2 * i}", tm, ind)
            ind = tm
        sd[f"time_modulation.net.{2 * (tl - 1)}.bias"] = torch.full((tm,), -1.0)
        dl = hash_level_table(g("deform_n_levels", 14), g("deform_base_resolution", 16),
                              g("deform_per_level_scale", 1.5), g("deform_log2_hashmap_size", 19))
        df = g("deform_n_features_per_level", 2)
        for nm in ("deform_grid_start", "deform_grid_mid", "deform_grid_end"):
            hashgrid(nm, dl, df)
        sd["deformation_grid.encoding.params"] = sd["deform_grid_start.encoding.params"]
        sd["deform_decoder.deform_net.params"] = _fused_init(len(dl) * df + tm, 3, g("deform_hidden_dim", 64), 2, gen)
        sd["deform_decoder.displacement_scale"] = torch.tensor(0.1)
        hashgrid("canonical_repr", levels, nf)
        inst_dec("decoder", len(levels) * nf + time_dim, dir_dim)
        return sd
    raise ValueError(mode)


# --------------------------------------------------------------------------
# synthetic NeRF-Synthetic-shaped inputs       (src/dataset.py:73-122,
#                                               look-at poses as run.py:1394-1417)
# (b2n/synthetic.py holds the same generators for the product side, which may not import oracle/;
#  tests/test_synthetic.py pins the two copies to each other)
# --------------------------------------------------------------------------

CAMERA_ANGLE_X = 0.6911112070083618


def synthetic_poses(n: int, seed: int = 0, radius: float = 4.0311) -> Tensor:
    """n camera-to-world matrices on the upper hemisphere looking at the origin
    (x right, y up, z backward; world up +Z)."""
    rng = np.random.RandomState(seed)
    el = np.deg2rad(rng.uniform(0.0, 60.0, n))
    az = rng.uniform(0.0, 2 * np.pi, n)
    out = np.zeros((n, 4, 4), dtype=np.float32)
    for i in range(n):
        c = radius * np.array([np.cos(el[i]) * np.cos(az[i]), np.cos(el[i]) * np.sin(az[i]), np.sin(el[i])])
        fwd = -c / np.linalg.norm(c)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        out[i, :3, 0], out[i, :3, 1], out[i, :3, 2], out[i, :3, 3] = right, up, -fwd, c
        out[i, 3, 3] = 1.0
    return torch.from_numpy(out)


def synthetic_rays(n_rays: int, H: int = 800, W: int = 800, n_views: int = 100, seed: int = 0,
                   with_time: bool = False):
    """Random training rays drawn like ``sample_random_rays`` (src/dataset.py:147-171):
    dirs ((i-W/2)/f, -(j-H/2)/f, -1) rotated by c2w and normalised; origin = camera centre."""
    gen = torch.Generator().manual_seed(seed)
    poses = synthetic_poses(n_views, seed)
    focal = 0.5 * W / np.tan(0.5 * CAMERA_ANGLE_X)
    img = torch.randint(0, n_views, (n_rays,), generator=gen)
    py = torch.randint(0, H, (n_rays,), generator=gen)
    px = torch.randint(0, W, (n_rays,), generator=gen)
    c2w = poses[img]
    dirs = torch.stack([(px - W * 0.5) / focal, -(py - H * 0.5) / focal, -torch.ones_like(px)], dim=-1).float()
    rd = torch.bmm(c2w[:, :3, :3], dirs.unsqueeze(-1)).squeeze(-1)
    rd = rd / torch.norm(rd, dim=-1, keepdim=True)
    ro = c2w[:, :3, 3].contiguous()
    target = torch.rand(n_rays, 4, generator=gen)
    if with_time:
        times = (img.float() / max(n_views - 1, 1)).unsqueeze(-1)
        return ro, rd.contiguous(), target, times
    return ro, rd.contiguous(), target


def sample_rays_host(poses: Tensor, rgba8: np.ndarray, times: Optional[Tensor], img_idx: Tensor, pix_y: Tensor,
                     pix_x: Tensor, H: int, W: int, focal: float, scene_scale: float = 1.0):
    """Training-ray batch for given pixel picks, as BlenderDataset / DynamicDataset.sample_random_rays builds it
    (src/dataset.py:147-171, :268-294): pinhole directions ((x - W/2)/f, -(y - H/2)/f, -1) rotated by c2w[:3,:3] and
    normalised, origin c2w[:3,3] * scene_scale, RGBA target uint8 / 255, per-frame time."""
    c2w = poses[img_idx]
    dirs = torch.stack([(pix_x - W * 0.5) / focal, -(pix_y - H * 0.5) / focal, -torch.ones_like(pix_x)], dim=-1)
    rays_d = torch.bmm(c2w[:, :3, :3], dirs.unsqueeze(-1)).squeeze(-1)
    rays_o = c2w[:, :3, 3]
    if scene_scale != 1.0:
        rays_o = rays_o * scene_scale
    target = torch.from_numpy(rgba8[img_idx.numpy(), pix_y.numpy(), pix_x.numpy()].astype(np.float32) / 255.0)
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    t_out = times[img_idx].unsqueeze(-1) if times is not None else None
    return rays_o, rays_d, target, t_out


def ball_occupancy(R: int, bound: float, radius: float = 0.75) -> Tensor:
    """Analytic occupancy: voxel active iff its corner point
    lies inside a ball of ``radius`` at the origin."""
    p = grid_corner_points(R, bound)
    return (p.norm(dim=-1) < radius).reshape(R, R, R)
