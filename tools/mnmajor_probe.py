#!/usr/bin/env python
"""Establishes the tcgen05 MN-major smem-descriptor encoding empirically (development aid for a tensor-core weight-gradient
kernel).  `python tools/mnmajor_probe.py` runs each candidate (lbo, sbo, k-advance, major bits) in its own process (a wrong
stride can fault) and reports which reproduce A^T B."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "project-nerf_b200")]

CANDIDATES = [(8192, 1024, 2048, 0x18000), (1024, 8192, 2048, 0x18000), (8192, 1024, 256, 0x18000), (8192, 1024, 32, 0x18000),
              (1024, 8192, 32, 0x18000), (8192, 128, 2048, 0x18000), (128, 8192, 2048, 0x18000), (8192, 1024, 2048, 0),
              (16, 1024, 2048, 0x18000), (8192, 2048, 2048, 0x18000), (2048, 8192, 2048, 0x18000), (1024, 1024, 2048, 0x18000)]


def one(lbo, sbo, kadv, extra):
    import torch
    from b2n._lib import call, ptr, stream
    torch.manual_seed(0)
    A = torch.randn(64, 128, device="cuda").to(torch.bfloat16)
    B = torch.randn(64, 128, device="cuda").to(torch.bfloat16)
    ref = A.float().t() @ B.float()
    D = torch.zeros(128, 128, device="cuda")
    call("b2n_debug_mnmajor_probe", ptr(A), ptr(B), ptr(D), lbo, sbo, kadv, extra, stream())
    torch.cuda.synchronize()
    err = float((D - ref).abs().max() / ref.abs().max())
    # also the error against B^T A (swapped roles) and against the K-major reading, to recognise near misses
    print("rel err %.3e  (vs transposed result %.3e)  lbo=%d sbo=%d kadv=%d idesc_extra=0x%x"
          % (err, float((D - ref.t()).abs().max() / ref.abs().max()), lbo, sbo, kadv, extra), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 5:
        one(*[int(v, 0) for v in sys.argv[1:]])
    else:
        for c in CANDIDATES:
            r = subprocess.run([sys.executable, __file__] + [str(v) for v in c], capture_output=True, text=True, timeout=120)
            out = [l for l in r.stdout.splitlines() if l.startswith("rel err")]
            print(out[0] if out else "FAULT/ERROR   lbo=%d sbo=%d kadv=%d idesc_extra=0x%x" % c, flush=True)
