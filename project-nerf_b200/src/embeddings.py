"""Encoders of the hot path on B200 kernels.

FourierRepresentation  <- reference src/embeddings.py:6-36   (b2n_pe_fwd / b2n_pe_bwd)
HashRepresentation     <- reference src/embeddings.py:39-93  (b2n_hash_fwd / b2n_hash_bwd;
                          replaces the tinycudann HashGrid the reference imports)
"""
import torch
from torch import nn

import b2n
from .abstract import BaseRepresentation


class FourierRepresentation(BaseRepresentation):
    """gamma(x) = [x, sin(x f_k pi), cos(x f_k pi)]_k, f_k = 2^k; ``freq_bands`` is a
    persistent buffer exactly like the reference's (it is part of the checkpoint format)."""

    def __init__(self, input_dim=2, L=10, use_encoding=True):
        super().__init__()
        self.input_dim, self.L, self.use_encoding = input_dim, L, use_encoding
        active = bool(use_encoding and L > 0)
        bands = 2.0 ** torch.linspace(0.0, L - 1, steps=L) if active else torch.empty(0)
        self.register_buffer("freq_bands", bands)
        self._out_dim = input_dim + 2 * input_dim * L if active else input_dim

    def forward(self, x):
        if not self.use_encoding or self.L == 0:
            return x
        return b2n.fourier_encode(x, self.freq_bands)

    @property
    def out_dim(self):
        return self._out_dim


class HashGridEncoding(nn.Module):
    """Stand-in for ``tcnn.Encoding(3, {"otype": "HashGrid", ...})``: owns the flat fp32
    ``params`` vector (level-major, entry-major, feature-minor) that run.py reaches into
    for TV losses and perturbation (run.py:614,1116,1849; src/core.py:192-196) and exposes
    ``n_output_dims``.  Calling it encodes unit-cube coordinates like tcnn does."""

    def __init__(self, n_input_dims, encoding_config, seed=1337):
        super().__init__()
        c = encoding_config
        if n_input_dims != 3 or c.get("otype", "HashGrid") != "HashGrid":
            raise ValueError("only the 3-D HashGrid encoding is on the hot path")
        self.geometry = b2n.HashGeometry(c["n_levels"], c["base_resolution"], c["per_level_scale"],
                                         c["log2_hashmap_size"], c["n_features_per_level"])
        self.n_input_dims = 3
        self.n_output_dims = self.geometry.out_dim
        gen = torch.Generator().manual_seed(seed)      # fixed seed: equal-shaped grids start identical, as upstream
        self.params = nn.Parameter((torch.rand(self.geometry.n_params, generator=gen) * 2 - 1) * 1e-4)

    def forward(self, x01):
        return b2n.hash_encode(x01, self.params, self.geometry, 0.0)


class HashRepresentation(BaseRepresentation):
    """World coords in [-bound, bound]^3 -> clamp to the unit cube -> multiresolution hash
    features.  Normalisation, clamp, 16 x 8 gathers and the trilinear blend are one kernel."""

    def __init__(self, n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
                 per_level_scale=1.5, bound=1.0):
        super().__init__()
        self.bound = bound
        self.encoding = HashGridEncoding(3, {
            "otype": "HashGrid", "n_levels": n_levels, "n_features_per_level": n_features_per_level,
            "log2_hashmap_size": log2_hashmap_size, "base_resolution": base_resolution,
            "per_level_scale": per_level_scale})
        self._out_dim = self.encoding.n_output_dims

    def forward(self, x):
        return b2n.hash_encode(x, self.encoding.params, self.encoding.geometry, float(self.bound))

    @property
    def out_dim(self):
        return self._out_dim
