"""Seeded synthetic inputs for benches and tests: ray batches shaped like the reference's datasets produce.
No dataset is available offline; these mimic datasets/satellite.py:18-65 + :218-227 (RPC rays normalised to the scene cube,
one sun direction per image) and datasets/blender.py:115-149 (pinhole rays, near 2 / far 6)."""
import math

import torch


def synthetic_sat_rays(n_rays, n_images=17, seed=0, dtype=torch.float32):
    """Rays shaped like datasets/satellite.py:18-65 + :218-227 output: (R,11) = o,d,near,far,sun; ts (R,)."""
    g = torch.Generator().manual_seed(seed)
    ts = torch.randint(0, n_images, (n_rays,), generator=g)
    inc = torch.deg2rad(5 + 30 * torch.rand(n_images, generator=g))
    azv = 2 * math.pi * torch.rand(n_images, generator=g)
    view = torch.stack([torch.sin(inc) * torch.cos(azv), torch.sin(inc) * torch.sin(azv), -torch.cos(inc)], -1)
    d = view[ts] + 1e-3 * torch.randn(n_rays, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    o = torch.cat([2 * torch.rand(n_rays, 2, generator=g) - 1, torch.ones(n_rays, 1)], -1)
    near = torch.zeros(n_rays, 1)
    far = 0.3 + 0.3 * torch.rand(n_rays, 1, generator=g)
    el = torch.deg2rad(30 + 40 * torch.rand(n_images, generator=g))
    az = torch.deg2rad(90 + 110 * torch.rand(n_images, generator=g))
    sun = torch.stack([torch.sin(az) * torch.cos(el), torch.cos(az) * torch.cos(el), torch.sin(el)], -1)[ts]
    return torch.cat([o, d, near, far, sun], -1).to(dtype), ts


def synthetic_blender_rays(n_rays, seed=0, dtype=torch.float32):
    """Rays shaped like datasets/blender.py:115-149: pinhole camera on a radius-4 sphere; (R,8), near 2, far 6."""
    g = torch.Generator().manual_seed(seed)
    th = 2 * math.pi * torch.rand(1, generator=g)
    ph = torch.deg2rad(20 + 40 * torch.rand(1, generator=g))
    cam = 4 * torch.tensor([torch.cos(th) * torch.cos(ph), torch.sin(th) * torch.cos(ph), torch.sin(ph)])
    fwd = -cam / cam.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0])); right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    px = (400 * torch.rand(n_rays, 2, generator=g) - 200) / 555.5
    d = fwd[None] + px[:, :1] * right[None] + px[:, 1:] * up[None]
    d = d / d.norm(dim=-1, keepdim=True)
    o = cam[None].expand(n_rays, 3)
    return torch.cat([o, d, torch.full((n_rays, 1), 2.0), torch.full((n_rays, 1), 6.0)], -1).to(dtype)
