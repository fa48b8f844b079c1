"""Host-side helpers run.py imports from ``src.utils`` (reference src/utils.py): PSNR, the
OOM-halving render wrapper, experiment naming and a thin TensorBoard logger.  No hot-path
arithmetic lives here."""
import math
import os
import warnings
from datetime import datetime

import torch

warnings.filterwarnings("ignore", category=DeprecationWarning)
warnings.filterwarnings("ignore", category=FutureWarning)
os.environ.setdefault("TF_CPP_MIN_LOG_LEVEL", "3")


def compute_psnr(mse):
    """10 log10(1 / mse) for images in [0, 1]."""
    return 10.0 * math.log10(1.0 / mse)


def compute_psnr_torch(pred, target):
    return compute_psnr(torch.mean((pred - target) ** 2).item())


def render_image_safe(render_fn, model, rays_o, rays_d, near, far, n_samples, chunk, white_bkgd):
    """Calls ``render_fn`` and, on a CUDA OOM, halves ``chunk`` (down to 1024) and retries.
    Allocation stays in torch, so our kernels surface OOM as torch.cuda.OutOfMemoryError too."""
    size = int(chunk)
    while True:
        try:
            return render_fn(model=model, rays_o=rays_o, rays_d=rays_d, near=near, far=far, n_samples=n_samples,
                             chunk=size, white_bkgd=white_bkgd)
        except torch.cuda.OutOfMemoryError:
            if not torch.cuda.is_available() or size <= 1024:
                raise
            torch.cuda.empty_cache()
            size = max(size // 2, 1024)
            print(f">>> CUDA OOM, reducing render chunk to {size}")


def get_exp_name(cfg):
    return cfg.get("exp_name", datetime.now().strftime("%Y%m%d_%H%M%S"))


class TensorBoardLogger:
    """SummaryWriter wrapper that silently disables itself when tensorboard is not installed."""

    def __init__(self, log_dir):
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer, self.enabled = SummaryWriter(log_dir), True
        except ImportError:
            print("!!! TensorBoard not installed, logging disabled")
            self.writer, self.enabled = None, False

    def log_scalar(self, tag, value, step):
        if self.enabled:
            self.writer.add_scalar(tag, value, step)

    def log_scalars(self, main_tag, tag_scalar_dict, step):
        if self.enabled:
            self.writer.add_scalars(main_tag, tag_scalar_dict, step)

    def close(self):
        if self.enabled:
            self.writer.close()
