"""Synthetic scene builders shared by the golden generator, the tests, smoke() and bench.py.

Nothing here computes the hot path; it only makes inputs:
  * energise_     - rescales default-initialised NeRF weights so that densities, colours and flows
                    actually vary along a ray (SURVEY.md §8c caveat: default init is near-constant).
  * synthetic_rays - NDC-like rays of the kind datasets/ray_utils.py:62-106 emits.
  * FakeDataset   - the five attributes rendering.py:190-200 reads from ``kwargs['dataset']``.
"""
import math

import numpy as np
import torch


@torch.no_grad()
def energise_(module, trunk_gain=2.0, sigma_gain=30.0, head_gain=3.0):
    """In-place, deterministic rescale of a module that uses the reference's parameter names."""
    for name, p in module.named_parameters():
        if not name.endswith("weight"):
            if name.startswith("static_sigma") or name.startswith("transient_sigma"):
                p.add_(0.5)
            continue
        if "_xyz_encoding_" in name or "dir_encoding" in name:
            p.mul_(trunk_gain)
        elif name.startswith("static_sigma") or name.startswith("transient_sigma"):
            p.mul_(sigma_gain)
        elif "_rgb" in name or "_flow_" in name:
            p.mul_(head_gain)
    return module


def synthetic_rays(n_rays, n_frames=30, seed=0, img_wh=None, frame=None):
    """(rays (R,6) fp32, ts (R,) int64).  o=(x,y,-1), d=(dx,dy,2) so o+d*z spans the NDC cube."""
    g = torch.Generator().manual_seed(seed)
    if img_wh is not None:
        w, h = img_wh
        assert n_rays == w * h
        ys, xs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
        x, y = xs.reshape(-1), ys.reshape(-1)
        dx = 0.1 * torch.sin(3.1 * x + 0.7) * torch.cos(2.3 * y)
        dy = 0.1 * torch.cos(2.7 * x) * torch.sin(1.9 * y + 0.3)
    else:
        x = torch.rand(n_rays, generator=g) * 2 - 1
        y = torch.rand(n_rays, generator=g) * 2 - 1
        dx = 0.1 * torch.randn(n_rays, generator=g)
        dy = 0.1 * torch.randn(n_rays, generator=g)
    o = torch.stack([x, y, -torch.ones_like(x)], -1)
    d = torch.stack([dx, dy, 2 * torch.ones_like(x)], -1)
    rays = torch.cat([o, d], -1).float().contiguous()
    if frame is None:
        ts = torch.randint(0, n_frames, (n_rays,), generator=g, dtype=torch.int64)
    else:
        ts = torch.full((n_rays,), int(frame), dtype=torch.int64)
    return rays, ts


class FakeDataset:
    """Pinhole cameras looking down -z from around the origin; some NDC points fall outside."""

    def __init__(self, n_frames=6, n_cams=2, img_wh=(32, 18), seed=0):
        rng = np.random.RandomState(seed)
        w, h = img_wh
        self.img_wh = img_wh
        self.N_frames = n_frames
        self.cam_train = list(range(n_cams))
        f = 0.9 * w
        self.Ks = [torch.tensor([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]], dtype=torch.float32)
                   for _ in range(n_cams)]
        poses = []
        for _ in range(n_cams * n_frames):
            ang = rng.uniform(-0.25, 0.25, size=3)
            cx, sx = math.cos(ang[0]), math.sin(ang[0])
            cy, sy = math.cos(ang[1]), math.sin(ang[1])
            cz, sz = math.cos(ang[2]), math.sin(ang[2])
            Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
            Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
            Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
            R = Rz @ Ry @ Rx
            t = rng.uniform(-0.4, 0.4, size=(3, 1))
            poses.append(np.concatenate([R, t], 1).astype(np.float32))
        self.poses = np.stack(poses, 0)
