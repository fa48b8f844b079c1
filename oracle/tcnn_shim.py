"""Stand-in ``tinycudann`` module built on the oracle's restatement.

TEST INFRASTRUCTURE ONLY.  ``install()`` registers a fake ``tinycudann`` in
``sys.modules`` so that the reference's own files (src/embeddings.py:57-73,
src/decoders.py:107-134, :281-295) import and run verbatim on CPU; only the
library arithmetic is restated (oracle.nerf_oracle.hash_encode / fused_mlp,
fp32).  Used by tests/golden/make_golden.py in the build container and by the
``--impl reference`` CPU timing when the reference tree is present.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn

from . import nerf_oracle as O


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        c = encoding_config
        if c.get("otype") != "HashGrid" or n_input_dims != 3:
            raise NotImplementedError("shim covers the 3-D HashGrid encoding only")
        self.levels = O.hash_level_table(c["n_levels"], c["base_resolution"], c["per_level_scale"],
                                         c["log2_hashmap_size"])
        self.n_feat = c["n_features_per_level"]
        self.n_input_dims = n_input_dims
        self.n_output_dims = len(self.levels) * self.n_feat
        gen = torch.Generator().manual_seed(seed)          # fixed seed -> identical grids, cf. src/core.py:191-196
        n = O.hash_table_entries(self.levels) * self.n_feat
        self.params = nn.Parameter((torch.rand(n, generator=gen) * 2 - 1) * 1e-4)

    def forward(self, x):
        return O.hash_encode(x, self.params.float(), self.levels, self.n_feat)


# what the padded input columns of a Network hold: 0.0 (the oracle's choice) or 1.0 (the other plausible
# upstream behaviour; tests/golden/make_golden.py writes a checkpoint fixture under it for b2n.checkpoint)
INPUT_PAD_VALUE = 0.0


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        c = network_config
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.n_neurons, self.n_hidden = c["n_neurons"], c["n_hidden_layers"]
        if c.get("activation", "ReLU") != "ReLU":
            raise NotImplementedError
        self.out_act = c.get("output_activation", "None")
        gen = torch.Generator().manual_seed(seed)
        self.params = nn.Parameter(O._fused_init(n_input_dims, n_output_dims, self.n_neurons, self.n_hidden, gen))

    def forward(self, x):
        return O.fused_mlp(x, self.params.float(), self.n_input_dims, self.n_output_dims, self.n_neurons,
                           self.n_hidden, self.out_act, pad_value=INPUT_PAD_VALUE)


def install():
    mod = types.ModuleType("tinycudann")
    mod.Encoding, mod.Network = Encoding, Network
    mod.__b2n_shim__ = True
    sys.modules["tinycudann"] = mod
    return mod
