"""CPU oracle for ray generation (the step immediately before the hot path).

TEST INFRASTRUCTURE ONLY (see mip360_oracle.py).  NumPy restatement of the reference's pinhole ray generator
(dataset.py:109-145), the LLFF NDC variant (dataset.py:364-387) and convert_to_ndc (intern/ray.py:59-79), pinned
against outputs of the literal reference classes (tests/golden/make_golden_raygen.py -> raygen_golden.npz).

Reference quirk kept on purpose: the "dx" of the pinhole radii is taken along axis 1 of the [n_img, h, w, 3]
array, i.e. between vertically adjacent pixels (rows), although the comment in the source says x-axis.
"""
import numpy as np


def pinhole_rays(cam_to_world, h, w, focal, near, far):
    """dataset.py:109-145.  cam_to_world [n,>=3,4] -> dict of [n,h,w,c] float32 arrays."""
    c2w = np.asarray(cam_to_world, dtype=np.float32)
    x, y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing="xy")
    cam = np.stack([(x - w * 0.5 + 0.5) / focal, -(y - h * 0.5 + 0.5) / focal, -np.ones_like(x)], axis=-1)
    directions = (cam[None, ..., None, :] * c2w[:, None, None, :3, :3]).sum(axis=-1)
    origins = np.broadcast_to(c2w[:, None, None, :3, -1], directions.shape)
    viewdirs = directions / np.linalg.norm(directions, axis=-1, keepdims=True)
    dx = np.sqrt(np.sum((directions[:, :-1, :, :] - directions[:, 1:, :, :]) ** 2, -1))
    dx = np.concatenate([dx, dx[:, -2:-1, :]], 1)
    radii = dx[..., None] * 2 / np.sqrt(12)
    ones = np.ones_like(origins[..., :1])
    return dict(origins=origins, directions=directions, viewdirs=viewdirs, radii=radii, near=ones * near, far=ones * far)


def convert_to_ndc(origins, directions, focal, w, h, near=1.0):
    """intern/ray.py:59-79."""
    t = -(near + origins[..., 2]) / (directions[..., 2] + 1e-15)
    origins = origins + t[..., None] * directions
    dx, dy, dz = tuple(np.moveaxis(directions, -1, 0))
    ox, oy, oz = tuple(np.moveaxis(origins, -1, 0))
    o0 = -((2 * focal) / w) * (ox / (oz + 1e-15))
    o1 = -((2 * focal) / h) * (oy / (oz + 1e-15))
    o2 = 1 + 2 * near / (oz + 1e-15)
    d0 = -((2 * focal) / w) * (dx / (dz + 1e-15) - ox / (oz + 1e-15))
    d1 = -((2 * focal) / h) * (dy / (dz + 1e-15) - oy / (oz + 1e-15))
    d2 = -2 * near / (oz + 1e-15)
    return np.stack([o0, o1, o2], -1), np.stack([d0, d1, d2], -1)


def llff_ndc_rays(cam_to_world, h, w, focal, near, far):
    """dataset.py:364-387: pinhole rays -> NDC origins/directions, radii from NDC-origin neighbours in both axes."""
    r = pinhole_rays(cam_to_world, h, w, focal, near, far)
    o, d = convert_to_ndc(r["origins"], r["directions"], focal, w, h)
    dx = np.sqrt(np.sum((o[:, :-1, :, :] - o[:, 1:, :, :]) ** 2, -1))
    dx = np.concatenate([dx, dx[:, -2:-1, :]], 1)
    dy = np.sqrt(np.sum((o[:, :, :-1, :] - o[:, :, 1:, :]) ** 2, -1))
    dy = np.concatenate([dy, dy[:, :, -2:-1]], 2)
    radii = (0.5 * (dx + dy))[..., None] * 2 / np.sqrt(12)
    ones = np.ones_like(o[..., :1])
    return dict(origins=o, directions=d, viewdirs=r["viewdirs"], radii=radii, near=ones * near, far=ones * far)


def flatten(rays):
    """dataset.py:147-152: [n,h,w,c] -> [n*h*w, c] float32."""
    return {k: np.asarray(v, dtype=np.float32).reshape(-1, v.shape[-1]) for k, v in rays.items()}
