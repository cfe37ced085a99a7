"""Synthetic inputs of the benchmark configurations (BASELINE.json configs): datasets are not available
offline, so rays come from random poses.  Host-side helpers only (torch CPU / Python floats); the rays themselves are
produced by the device generator (ops.generate_rays) or drawn with a seeded CPU generator."""
from __future__ import annotations

import math

import torch


def generic_rays(B, seed, device=None, pin=False):
    """'Generic' rays: origins, directions ~ N(0,I) (directions not normalised, like pinhole directions,
    dataset.py:123), viewdirs = d/|d|, radii 1e-3, near 0.1, far 10, plus uniform target pixels."""
    from mipnerf360_b200.intern.ray import Rays
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(B, 3, generator=g)
    d = torch.randn(B, 3, generator=g)
    rays = Rays(o, d, d / d.norm(dim=-1, keepdim=True), torch.full((B, 1), 1e-3), torch.full((B, 1), 0.1),
                torch.full((B, 1), 10.0))
    pixels = torch.rand(B, 3, generator=g)
    if pin:
        rays = Rays(*[r.pin_memory() for r in rays])
        pixels = pixels.pin_memory()
    if device is not None:
        rays = Rays(*[r.to(device) for r in rays])
        pixels = pixels.to(device)
    return rays, pixels


def look_at(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """Camera-to-world matrix [4,4] of a camera at `eye` looking at `target` (pose.py:101-110 convention: the
    camera looks along -z, +y is up)."""
    eye, target, up = (torch.tensor(v, dtype=torch.float64) for v in (eye, target, up))
    z = eye - target
    z = z / z.norm()
    x = torch.linalg.cross(up, z)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    c2w = torch.eye(4, dtype=torch.float64)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, eye
    return c2w.float()


def perturbed_identity(seed, max_deg=5.0, max_shift=0.1):
    """Small random rotation (<= max_deg about a random axis) and translation (<= max_shift) of the identity pose."""
    g = torch.Generator().manual_seed(seed)
    axis = torch.randn(3, generator=g, dtype=torch.float64)
    axis = axis / axis.norm()
    ang = math.radians(max_deg) * float(torch.rand((), generator=g))
    K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]], dtype=torch.float64)
    R = torch.eye(3, dtype=torch.float64) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)
    c2w = torch.eye(4, dtype=torch.float64)
    c2w[:3, :3] = R
    c2w[:3, 3] = (torch.rand(3, generator=g, dtype=torch.float64) * 2 - 1) * max_shift
    return c2w.float()


# name -> dict(height, width, focal, c2w, near, far, ndc): BASELINE.json configs[2] and configs[3]
def llff_case(height=756, width=1008, seed=0):
    """LLFF-shaped frame (4032x3024 / factor 4): forward-facing pose, NDC rays, near 0.05 (the loader's near = 0 collapses
    every sample through g(0 + 1e-6)), far 1."""
    return dict(name=f"llff_{width}x{height}_ndc", height=height, width=width, focal=0.82 * width,
                c2w=perturbed_identity(seed), near=0.05, far=1.0, ndc=True)


def garden_case(height=3286, width=4946, angle=0.0):
    """Garden-shaped unbounded frame: camera on a circle of radius 4 at height 1.5 looking at the origin, no NDC,
    near 0.2, far 1e3, so most samples lie outside the unit ball (contraction-heavy)."""
    eye = (4.0 * math.cos(angle), 4.0 * math.sin(angle), 1.5)
    return dict(name=f"garden_{width}x{height}_unbounded", height=height, width=width, focal=0.8 * width,
                c2w=look_at(eye), near=0.2, far=1e3, ndc=False)
