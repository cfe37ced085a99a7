"""intern/pose.py:112-212 — depth / normal visualisation of a rendered frame, on the device.

Only the two entry points test.py:52-56 and video.py:40-43 call are mirrored (`visualize_depth`,
`visualize_normals`, plus the `sinebow` colour map they default to); the camera-path helpers of pose.py are host-side
setup code outside the hot path.  Inputs may be NumPy arrays (as `render_image` returns them) or CUDA tensors; the
result comes back in the same kind.  `as_uint8=True` returns `to8b(...)` of the picture directly (3 B/pixel).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def _to_device(x):
    if x is None:
        return None, False
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=torch.float32), False
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda(), True


def _back(t, as_numpy):
    return t.cpu().numpy() if as_numpy else t


def sinebow(h):
    """pose.py:123-126: a cyclic and uniform colour map, sin(pi (k/6 - h))^2 for k = 3, 5, 7."""
    h = torch.as_tensor(h, dtype=torch.float32)
    f = lambda x: torch.sin(torch.pi * x) ** 2  # noqa: E731
    return torch.stack([f(3 / 6 - h), f(5 / 6 - h), f(7 / 6 - h)], -1)


def turbo_lut(n=256):
    """Default colour table when modulus == 0.  The reference asks matplotlib for 'turbo' (pose.py:204): when
    matplotlib is importable its table is used as is.  It is not part of this image, so otherwise the table is rebuilt
    from the published 5th-order polynomial fit of turbo, which is visually close in the interior but deviates by up
    to about 0.13 (blue channel) at the dark end.  Pass `colormap=` (a table or a matplotlib Colormap) for exact colours."""
    try:
        import matplotlib
        return np.asarray(matplotlib.colormaps["turbo"](np.arange(n) if n == 256 else np.linspace(0, 1, n)),
                          dtype=np.float32)[:, :3]
    except Exception:
        pass
    x = np.arange(n, dtype=np.float64) / (n - 1)
    v = np.stack([np.ones_like(x), x, x ** 2, x ** 3, x ** 4, x ** 5], 0)
    coef = np.array([[0.13572138, 4.61539260, -42.66032258, 132.13108234, -152.94239396, 59.28637943],
                     [0.09140261, 2.19418839, 4.84296658, -14.18503333, 4.27729857, 2.82956604],
                     [0.10667330, 12.64194608, -60.58204836, 110.36276771, -89.90310912, 27.34824973]])
    return np.clip(coef @ v, 0.0, 1.0).T.astype(np.float32)


def _colour_table(colormap):
    """None | 'sinebow' | the sinebow function -> None (evaluated on the device); tables and listed colour maps ->
    host fp32 table [n,3]."""
    if colormap is None or colormap is sinebow or (isinstance(colormap, str) and colormap == "sinebow"):
        return None
    if hasattr(colormap, "colors"):  # matplotlib ListedColormap
        colormap = colormap.colors
    elif hasattr(colormap, "N") and callable(colormap):  # any matplotlib Colormap: integer input indexes its table
        colormap = colormap(np.arange(colormap.N))
    if callable(colormap):
        raise TypeError("visualize_depth: colormap must be a colour table [n,3], a matplotlib Colormap or sinebow; "
                        "arbitrary Python callables cannot run on the device and there is no host fallback")
    table = np.asarray(colormap.cpu() if isinstance(colormap, torch.Tensor) else colormap, dtype=np.float32)
    if table.ndim != 2 or table.shape[1] < 3 or table.shape[0] < 1:
        raise ValueError(f"visualize_depth: colour table must be [n >= 1, >= 3], got {table.shape}")
    return np.ascontiguousarray(table[:, :3])


def visualize_normals(depth, acc, scaling=None, as_uint8=False):
    """pose.py:128-147.  As in the reference, the picture is only produced when `scaling` is None (the whole body,
    including the return, sits under that branch); a given scaling returns None."""
    if scaling is not None:
        return None
    d, as_numpy = _to_device(depth)
    a, _ = _to_device(acc)
    return _back(ops.visualize_normals(d, a, as_uint8=as_uint8), as_numpy)


def visualize_depth(depth, acc=None, near=None, far=None, ignore_frac=0, curve_fn="neg_log", modulus=0, colormap=None,
                    as_uint8=False):
    """pose.py:149-212.  curve_fn names one of the curves the docstring of the reference suggests: 'neg_log'
    (-log(x + eps), the default), 'identity', 'inverse' (1/(x + eps)), 'log'."""
    if callable(curve_fn) or curve_fn not in ops.CURVES:
        raise TypeError(f"visualize_depth: curve_fn must be one of {sorted(ops.CURVES)} (device curves; no host fallback)")
    d, as_numpy = _to_device(depth)
    a, _ = _to_device(acc)
    table = _colour_table(colormap)  # None = sinebow, evaluated on the device
    if colormap is None and not modulus > 0:
        table = turbo_lut()  # pose.py:204: turbo unless the depth is wrapped
    if table is not None:
        table = torch.as_tensor(table).cuda()
    out = ops.visualize_depth(d, a, near, far, ignore_frac, curve_fn, modulus, table, as_uint8=as_uint8)
    return _back(out, as_numpy)
