#!/usr/bin/env python
"""Workload for an ncu capture of k_mlp256 (training forward + backward chain at C1 size):
    ncu --set full --clock-control none --import-source on -k regex:k_mlp256 --launch-skip 4 --launch-count 2 \
        -o gpurun_out/mlp256 python tools/ncu_mlp256.py [pair]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "project-nerf_b200")]
import torch  # noqa: E402

import b2n  # noqa: E402
from src.core import NeuralField  # noqa: E402

pair = int(sys.argv[1]) if len(sys.argv) > 1 else 1
b2n._lib.lib.b2n_debug_mlp256_set_pair(pair)
torch.manual_seed(0)
model = NeuralField(dict(mode="part2_nerf", L_embed=10, L_embed_dir=4)).cuda().train()
P = 262144
xe = torch.randn(P, 63, device="cuda").requires_grad_(True)
de = torch.randn(P, 27, device="cuda")
for _ in range(4):
    r, s = model.decoder(xe, de)
    (r.sum() + s.sum()).backward()
torch.cuda.synchronize()
print("done")
