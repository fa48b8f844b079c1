#!/usr/bin/env python
"""Micro-timings of single entry points (CUDA events, L2-cold inputs) -- development aid.
    python tools/kbench.py mlp256 [P]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "project-nerf_b200")]
import torch  # noqa: E402

import b2n  # noqa: E402
from b2n import ops  # noqa: E402
from src.core import NeuralField  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def mlp256(P):
    torch.manual_seed(0)
    model = NeuralField(dict(mode="part2_nerf", L_embed=10, L_embed_dir=4)).cuda().eval()
    xe = torch.randn(P, 63, device="cuda")
    de = torch.randn(P, 27, device="cuda")
    flops = 2.0 * P * 593408
    for save in (False, True):
        med, best = timeit(lambda: ops.nerf_mlp_forward(model.decoder, xe, de, save=save))
        print(f"mlp256 fwd P={P} save={save}: median {med:.3f} ms best {best:.3f} ms -> {flops / best / 1e9:.1f} TFLOP/s "
              f"({100 * flops / best / 1e9 / 1391.5:.1f} % of sustained bf16 peak)")
    with torch.no_grad():
        b2n.set_mlp_precision("fp32")
        med, best = timeit(lambda: model.decoder(xe, de), n=3, warm=1)
        print(f"fp32 layer-wise path: {best:.3f} ms -> {flops / best / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "mlp256"
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    {"mlp256": mlp256}[what](P)
