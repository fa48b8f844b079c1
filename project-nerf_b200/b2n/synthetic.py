"""Synthetic NeRF-Synthetic-shaped inputs for benchmarks and smoke tests.

Host-side data generation only (numpy/torch CPU); nothing here computes the hot path.
Scene shape follows the reference's loaders: 800x800 views, camera_angle_x of the Blender
sets, pinhole rays built as in src/dataset.py:84-96,147-171, look-at poses on the upper
hemisphere as in run.py:1394-1417.

The same generators exist once more in oracle/nerf_oracle.py (synthetic_poses / synthetic_rays / ball_occupancy) ON PURPOSE:
the product may not import oracle/, and the oracle -- which also feeds the CPU reference arm of bench.py -- may not depend on
the product.  tests/test_synthetic.py pins the two copies to each other bit for bit.
"""
import numpy as np
import torch

CAMERA_ANGLE_X = 0.6911112070083618


def hemisphere_poses(n, seed=0, radius=4.0311):
    rng = np.random.RandomState(seed)
    el = np.deg2rad(rng.uniform(0.0, 60.0, n))
    az = rng.uniform(0.0, 2 * np.pi, n)
    c = radius * np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1)
    fwd = -c / np.linalg.norm(c, axis=-1, keepdims=True)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right, axis=-1, keepdims=True)
    up = np.cross(right, fwd)
    m = np.zeros((n, 4, 4), dtype=np.float32)
    m[:, :3, 0], m[:, :3, 1], m[:, :3, 2], m[:, :3, 3], m[:, 3, 3] = right, up, -fwd, c, 1.0
    return torch.from_numpy(m)


def random_rays(n_rays, H=800, W=800, n_views=100, seed=0, with_time=False):
    """(rays_o [B,3], rays_d [B,3] normalised, target_rgba [B,4][, times [B,1]]) on the CPU."""
    gen = torch.Generator().manual_seed(seed)
    poses = hemisphere_poses(n_views, seed)
    focal = 0.5 * W / np.tan(0.5 * CAMERA_ANGLE_X)
    img = torch.randint(0, n_views, (n_rays,), generator=gen)
    py = torch.randint(0, H, (n_rays,), generator=gen)
    px = torch.randint(0, W, (n_rays,), generator=gen)
    c2w = poses[img]
    dirs = torch.stack([(px - W * 0.5) / focal, -(py - H * 0.5) / focal, -torch.ones_like(px)], -1).float()
    rd = torch.bmm(c2w[:, :3, :3], dirs.unsqueeze(-1)).squeeze(-1)
    rd = (rd / rd.norm(dim=-1, keepdim=True)).contiguous()
    ro = c2w[:, :3, 3].contiguous()
    target = torch.rand(n_rays, 4, generator=gen)
    if with_time:
        return ro, rd, target, (img.float() / max(n_views - 1, 1)).unsqueeze(-1)
    return ro, rd, target


def image_rays(pose, H=800, W=800):
    """All rays of one view (src/dataset.py:101-122): [H*W,3] origins and unit directions."""
    focal = 0.5 * W / np.tan(0.5 * CAMERA_ANGLE_X)
    j, i = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1).float().reshape(-1, 3)
    rd = dirs @ pose[:3, :3].T
    rd = rd / rd.norm(dim=-1, keepdim=True)
    return pose[:3, 3].expand_as(rd).contiguous(), rd.contiguous()


def ball_occupancy(R, bound, radius=0.75):
    """Analytic occupancy: voxel active iff its corner point lies inside a ball at the origin."""
    ax = torch.linspace(-bound, bound, R)
    p = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1)
    return p.norm(dim=-1) < radius
