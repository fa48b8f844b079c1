"""NeuralField: per-mode wiring of encoders and decoders (reference src/core.py:9-363).

Same constructor (flat config dict), same ``forward(x, d=None, t=None)`` with the
per-mode return arity, same sub-module attribute names (run.py reaches into
``representation``, ``canonical_repr``, ``deform_grid_*``, ``deformation_grid``,
``deform_decoder``, ``time_modulation`` ... directly).  ``part1_fourier`` (2-D image
fitting) is outside the ray-marching hot path and is rejected.
"""
import torch
from torch import nn

import b2n

from .decoders import (DeformationNetwork, HashDeformationDecoder, InstantNeRFDecoder, NeRFDecoder,
                       TimeModulationNetwork)
from .embeddings import FourierRepresentation, HashRepresentation

_TRI_ANCHORS = (0.0, 0.5, 1.0)      # tri-grid time anchors (reference core.py:313-315)
_TRI_BANDWIDTH = 0.5


def _hash_from_cfg(cfg, prefix, defaults, bound_default):
    g = cfg.get
    return HashRepresentation(
        n_levels=g(prefix + "n_levels", defaults[0]),
        n_features_per_level=g(prefix + "n_features_per_level", 2),
        log2_hashmap_size=g(prefix + "log2_hashmap_size", defaults[1]),
        base_resolution=g(prefix + "base_resolution", 16),
        per_level_scale=g(prefix + "per_level_scale", 1.5),
        bound=g("scene_bound", bound_default),
    )


def _nerf_decoder_from_cfg(cfg, pos_dim, dir_dim):
    g = cfg.get
    return NeRFDecoder(pos_dim=pos_dim, dir_dim=dir_dim, hidden_dim=g("hidden_dim", 256),
                       num_layers=g("num_layers", 8), skip_layer=g("skip_layer", 4), view_dim=g("view_dim", 128))


class NeuralField(nn.Module):
    def __init__(self, config):
        super().__init__()
        g = config.get
        self.mode = config["mode"]
        self.use_coord_noise = g("use_coord_noise", False)
        self.coord_noise_std = g("coord_noise_std", 0.005)
        self.time_noise_std = g("time_noise_std", 0.02)
        mode = self.mode
        if mode == "part1_fourier":
            raise NotImplementedError("part1_fourier (2-D image fitting) is outside the ray-marching hot path")

        if mode == "part2_nerf":
            use_pe = g("use_positional_encoding", True)
            L = g("L_embed", 0) if use_pe else 0
            self.representation = FourierRepresentation(input_dim=3, L=L, use_encoding=use_pe)
            use_dir = g("use_viewdirs", True)
            self.dir_representation = FourierRepresentation(
                input_dim=3, L=g("L_embed_dir", 4) if use_dir else 0, use_encoding=use_dir)
            self.decoder = _nerf_decoder_from_cfg(config, self.representation.out_dim, self.dir_representation.out_dim)
        elif mode == "part2_instant":
            self.representation = _hash_from_cfg(config, "", (16, 19), 1.0)
            self.dir_representation = FourierRepresentation(input_dim=3, L=g("L_embed_dir", 4), use_encoding=True)
            self.decoder = InstantNeRFDecoder(pos_dim=self.representation.out_dim,
                                              dir_dim=self.dir_representation.out_dim, hidden_dim=g("hidden_dim", 64))
        elif mode == "part3":
            self.dir_representation = FourierRepresentation(input_dim=3, L=g("L_embed_dir", 4), use_encoding=True)
            self.time_encoder = FourierRepresentation(input_dim=1, L=g("L_embed_time", 10), use_encoding=True)
            self.pos_encoder_for_deform = FourierRepresentation(input_dim=3, L=g("L_embed", 10), use_encoding=True)
            self.deform_net = DeformationNetwork(pos_dim=self.pos_encoder_for_deform.out_dim,
                                                 time_dim=self.time_encoder.out_dim,
                                                 hidden_dim=g("deform_hidden_dim", 128),
                                                 num_layers=g("deform_num_layers", 4))
            if g("canonical_type", "nerf") == "instant":
                self.canonical_repr = _hash_from_cfg(config, "", (16, 19), 1.0)
                self.decoder = InstantNeRFDecoder(pos_dim=self.canonical_repr.out_dim + self.time_encoder.out_dim,
                                                  dir_dim=self.dir_representation.out_dim,
                                                  hidden_dim=g("hidden_dim", 64))
            else:
                self.canonical_repr = FourierRepresentation(input_dim=3, L=g("L_embed_canon", 10), use_encoding=True)
                self.decoder = _nerf_decoder_from_cfg(config, self.canonical_repr.out_dim + self.time_encoder.out_dim,
                                                      self.dir_representation.out_dim)
            self.direct_time_conditioning = g("direct_time_conditioning", False)
            if self.direct_time_conditioning:
                self.pos_encoder_direct = FourierRepresentation(input_dim=3, L=g("L_embed", 10), use_encoding=True)
                self.decoder_direct = _nerf_decoder_from_cfg(
                    config, self.pos_encoder_direct.out_dim + self.time_encoder.out_dim, self.dir_representation.out_dim)
        elif mode == "part4":
            self.dir_representation = FourierRepresentation(input_dim=3, L=g("L_embed_dir", 4), use_encoding=True)
            self.time_encoder = FourierRepresentation(input_dim=1, L=g("L_embed_time", 10), use_encoding=True)
            tm_dim = g("time_modulation_dim", 64)
            self.time_modulation = TimeModulationNetwork(time_dim=self.time_encoder.out_dim, output_dim=tm_dim,
                                                         hidden_dim=tm_dim, num_layers=g("time_modulation_layers", 2))
            self.deform_grid_start = _hash_from_cfg(config, "deform_", (14, 19), 1.5)
            self.deform_grid_mid = _hash_from_cfg(config, "deform_", (14, 19), 1.5)
            self.deform_grid_end = _hash_from_cfg(config, "deform_", (14, 19), 1.5)
            with torch.no_grad():          # the three grids start identical (fixed init seed): de-correlate mid/end
                for grid in (self.deform_grid_mid, self.deform_grid_end):
                    grid.encoding.params.add_(torch.randn_like(grid.encoding.params) * 1e-4)
            self.deformation_grid = self.deform_grid_start      # legacy alias kept by the reference (core.py:199)
            self.deform_decoder = HashDeformationDecoder(hash_dim=self.deform_grid_start.out_dim, time_mod_dim=tm_dim,
                                                         hidden_dim=g("deform_hidden_dim", 64))
            self.canonical_repr = _hash_from_cfg(config, "", (16, 19), 1.5)
            self.decoder = InstantNeRFDecoder(pos_dim=self.canonical_repr.out_dim + self.time_encoder.out_dim,
                                              dir_dim=self.dir_representation.out_dim, hidden_dim=g("hidden_dim", 64))
        else:
            raise ValueError(f"unknown mode {mode!r}")

    # ------------------------------------------------------------------ helpers
    def _augment(self, x, t):
        """train-time input noise of the deformation branch (reference core.py:254-262, :289-294)."""
        xd, td = x, t
        if self.training and self.use_coord_noise:
            if self.coord_noise_std > 0:
                xd = x + torch.randn_like(x) * self.coord_noise_std
            if self.time_noise_std > 0:
                td = torch.clamp(t + torch.randn_like(t) * self.time_noise_std, 0.0, 1.0)
        return xd, td

    @staticmethod
    def _tri_weights(t):
        ws = [torch.clamp(1.0 - torch.abs(t - a) / _TRI_BANDWIDTH, 0.0, 1.0) for a in _TRI_ANCHORS]
        total = ws[0] + ws[1] + ws[2] + 1e-8
        return [w / total for w in ws]

    def _tri_blend(self, xd, td):
        """sum_i w_i(t) * deform_grid_i(x) (reference core.py:308-335): one kernel each way when the three grids share
        their geometry (they always do: one config block builds all three), otherwise the module-by-module form"""
        grids = (self.deform_grid_start, self.deform_grid_mid, self.deform_grid_end)
        g0 = grids[0].encoding.geometry
        same = all(g.encoding.geometry.levels == g0.levels and g.bound == grids[0].bound for g in grids[1:])
        if xd.is_cuda and same and g0.n_features == 2 and not xd.requires_grad and not td.requires_grad:
            return b2n.ops.hash_tri_blend(xd, td, [g.encoding.params for g in grids], g0, float(grids[0].bound))
        w0, w1, w2 = self._tri_weights(td)
        return w0 * grids[0](xd) + w1 * grids[1](xd) + w2 * grids[2](xd)

    # ------------------------------------------------------------------ forward
    def forward(self, x, d=None, t=None):
        mode = self.mode
        if mode == "part3":
            if t is None:
                raise ValueError("Part 3 requires time input 't'.")
            if getattr(self, "direct_time_conditioning", False):
                h = torch.cat([self.pos_encoder_direct(x), self.time_encoder(t)], dim=-1)
                rgb, sigma = self.decoder_direct(h, self.dir_representation(d))
                return rgb, sigma, torch.zeros_like(x)
            xd, td = self._augment(x, t)
            feat_t = self.time_encoder(td)
            delta_x = self.deform_net(self.pos_encoder_for_deform(xd), feat_t)
            feat_can = self.canonical_repr(x + delta_x)
            rgb, sigma = self._decode(torch.cat([feat_can, feat_t], dim=-1), d)
            return rgb, sigma, delta_x
        if mode == "part4":
            if t is None:
                raise ValueError("Part 4 requires time input 't'.")
            xd, td = self._augment(x, t)
            feat_t = self.time_encoder(td)
            time_mod = self.time_modulation(feat_t)
            blend = self._tri_blend(xd, td)
            delta_x = self.deform_decoder(blend, time_mod)
            feat_can = self.canonical_repr(x + delta_x)
            rgb, sigma = self._decode(torch.cat([feat_can, feat_t], dim=-1), d)
            return rgb, sigma, delta_x
        if d is None:
            raise ValueError(f"{mode} requires view directions.")
        return self._decode(self.representation(x), d)

    def density(self, x, t=None):
        """sigma [P,1] only.  ``DensityGrid.update`` (reference src/renderer.py:108-116) evaluates the whole model on
        the R^3 lattice with zero view directions and keeps sigma; with a hash-grid / 64-wide decoder the colour
        branch (direction features + 2 of the 3 fused layers' worth of work) is skipped here.  In the
        16-bit mode sigma comes from the SAME fused decoder kernel as in ``forward`` with its colour network switched
        off (b2n.instant_sigma: bit-identical sigma, so occupancy decisions agree with what is rendered); in fp32 mode
        from the same layer-by-layer kernels as ``forward``.  The 256-wide NeRFDecoder has no separable density branch:
        it runs ``forward`` and drops rgb."""
        mode = self.mode
        dec = self.decoder
        instant = isinstance(dec, InstantNeRFDecoder)
        fusable = instant and x.is_cuda and dec.can_fuse(self.dir_representation)
        if mode == "part2_instant" and instant:
            feat = self.representation(x)
            if fusable:
                return b2n.instant_sigma(feat, dec.sigma_net.params, dec.sigma_net.input_pad_value)
            return dec.density(feat)
        if mode in ("part3", "part4") and instant and not getattr(self, "direct_time_conditioning", False):
            if t is None:
                raise ValueError(f"{mode} requires time input 't'.")
            xd, td = self._augment(x, t)
            feat_t = self.time_encoder(td)
            if mode == "part3":
                delta_x = self.deform_net(self.pos_encoder_for_deform(xd), feat_t)
            else:
                delta_x = self.deform_decoder(self._tri_blend(xd, td), self.time_modulation(feat_t))
            feat_can = self.canonical_repr(x + delta_x)
            if fusable:
                return b2n.instant_sigma(torch.cat([feat_can, feat_t], dim=-1), dec.sigma_net.params,
                                         dec.sigma_net.input_pad_value)
            return dec.density(feat_can, feat_t)
        out = self.forward(x, torch.zeros_like(x), t) if mode in ("part3", "part4") else self.forward(x, torch.zeros_like(x))
        return out[1]

    def _decode(self, feat, d):
        """decoder(feat, gamma(d)); the 64-wide Instant decoder runs fused on raw directions in bf16 mode."""
        dec = self.decoder
        if isinstance(dec, InstantNeRFDecoder) and dec.can_fuse(self.dir_representation):
            return dec.forward_fused(feat, d, self.dir_representation)
        return dec(feat, self.dir_representation(d))
