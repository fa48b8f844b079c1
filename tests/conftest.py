import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "project-nerf_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pytest_terminal_summary(terminalreporter):
    """achieved parity figures (tests/_util.record): printed, and dumped for profiles/ when run on the GPU box"""
    import json
    from _util import ACHIEVED
    if not ACHIEVED:
        return
    terminalreporter.write_sep("-", "achieved parity figures (max over the cases of each check)")
    for k in sorted(ACHIEVED):
        terminalreporter.write_line(f"{k:70s} {ACHIEVED[k]:.3e}")
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_maxima.json"), "w") as f:
            json.dump(ACHIEVED, f, indent=1, sort_keys=True)
    except OSError:
        pass
