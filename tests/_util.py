"""Shared helpers for the test-suite (fixture loading, tolerances)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out, sd, grads, gsum, ghead = {}, {}, {}, {}, {}
    for k in z.files:
        v = z[k]
        if k == "cfg":
            out[k] = json.loads(bytes(v).decode())
            continue
        special = k.startswith(("sd::", "grad::", "gradsum::", "gradhead::"))
        t = torch.from_numpy(np.array(v)) if (v.shape != () or special) else v.item()
        if k.startswith("sd::"):
            sd[k[4:]] = t
        elif k.startswith("grad::"):
            grads[k[6:]] = t
        elif k.startswith("gradsum::"):
            gsum[k[9:]] = t
        elif k.startswith("gradhead::"):
            ghead[k[10:]] = t
        else:
            out[k] = t
    out["sd"], out["grads"], out["gradsum"], out["gradhead"] = sd, grads, gsum, ghead
    return out


# achieved parity figures of a test session: name -> value (conftest.py prints them in the terminal summary and dumps
# them to gpurun_out/parity_maxima.json so that every tolerance in the suite can be read next to what was measured)
ACHIEVED = {}


def record(name, value):
    value = float(value)
    ACHIEVED[name] = max(ACHIEVED.get(name, 0.0), value)
    return value


def rel_l2(a, b):
    """||a-b|| / ||b||  (Frobenius)"""
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rel_err(a, b):
    """max |a-b| / (max|b| + tiny): the '1e-4 relative' of the north star, per tensor."""
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item() if b.numel() else 0.0


def full_nerf_state_dict(seed=31):
    """Rebuild the reference's 8x256 NeRFDecoder init bit-identically without the
    reference: same seed, nn.Linear modules created in src/decoders.py:50-66 order."""
    torch.manual_seed(seed)
    sd = {"representation.freq_bands": 2.0 ** torch.linspace(0.0, 9, steps=10),
          "dir_representation.freq_bands": 2.0 ** torch.linspace(0.0, 3, steps=4)}

    def lin(name, i, o):
        m = torch.nn.Linear(i, o)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach(), m.bias.detach()
    for i in range(8):
        lin(f"decoder.pts_layers.{i}", 63 if i == 0 else (319 if i == 4 else 256), 256)
    lin("decoder.sigma_layer", 256, 1)
    lin("decoder.feature_layer", 256, 256)
    lin("decoder.view_layer", 283, 128)
    lin("decoder.rgb_layer", 128, 3)
    return sd
