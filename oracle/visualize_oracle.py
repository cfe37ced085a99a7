"""CPU restatement of the reference's depth / normal visualisation (intern/pose.py:112-212).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing in the product path).  Pinned against the literal
reference functions by tests/golden/make_golden_visualize.py -> tests/golden/visualize_golden.npz.

Differences from the literal code, all deliberate and checked by the golden test:
  * the 3x3 convolutions are written out as stencils (scipy.signal.convolve2d(mode='same'), utils.py:9-10);
  * the automatic near/far planes use an exact fp64 cumulative sum of acc over the sorted depths instead of the
    reference's sequential float32 cumsum (pose.py:184-186) — same pixels selected except at float32 knife edges;
  * colour maps are tables [n,3] indexed like a matplotlib listed colour map (trunc(value * n), 1.0 -> n-1), or sinebow.
"""
import numpy as np

EPS = np.finfo(np.float32).eps


def conv_same_3x3(z, k):
    """scipy.signal.convolve2d(z, k, mode='same') for a 3x3 kernel: true convolution, zero fill."""
    H, W = z.shape
    zp = np.zeros((H + 2, W + 2), dtype=np.float64)
    zp[1:-1, 1:-1] = z
    out = np.zeros((H, W), dtype=np.float64)
    for i in range(3):
        for j in range(3):
            out += k[i, j] * zp[2 - i:2 - i + H, 2 - j:2 - j + W]
    return out


def depth_to_normals(depth):
    """pose.py:112-121."""
    f_blur = np.array([1, 2, 1]) / 4
    f_edge = np.array([-1, 0, 1]) / 2
    dy = conv_same_3x3(depth, f_blur[None, :] * f_edge[:, None])
    dx = conv_same_3x3(depth, f_blur[:, None] * f_edge[None, :])
    inv_denom = 1 / np.sqrt(1 + dx ** 2 + dy ** 2)
    return np.stack([dx * inv_denom, dy * inv_denom, inv_denom], -1)


def sinebow(h):
    """pose.py:123-126."""
    f = lambda x: np.sin(np.pi * x) ** 2  # noqa: E731
    return np.stack([f(3 / 6 - h), f(5 / 6 - h), f(7 / 6 - h)], -1)


def normals_scaling(depth):
    """pose.py:130-136.  As in the reference, the depth variance is NumPy's float32 pairwise reduction (relative error
    about 1e-7); the device path accumulates it in fp64."""
    mask = ~np.isnan(depth)
    x, y = np.meshgrid(np.arange(depth.shape[1]), np.arange(depth.shape[0]), indexing="xy")
    xy_var = (np.var(x[mask]) + np.var(y[mask])) / 2
    z_var = np.var(depth[mask])
    return np.sqrt(xy_var / z_var)


def visualize_normals(depth, acc):
    """pose.py:128-147 (scaling=None branch)."""
    with np.errstate(all="ignore"):
        normals = depth_to_normals(normals_scaling(depth) * depth.astype(np.float64))
        vis = np.isnan(normals) + np.nan_to_num((normals + 1) / 2)
        if acc is not None:
            vis = vis * acc[:, :, None] + (1 - acc)[:, :, None]
    return vis


def auto_planes(depth, acc, ignore_frac):
    """pose.py:180-189: depth values spanning the middle (1 - 2 ignore_frac) of the accumulated acc."""
    sortidx = np.argsort(depth.reshape(-1), kind="stable")
    depth_sorted = depth.reshape(-1)[sortidx]
    w = np.rint(np.clip(acc.reshape(-1)[sortidx].astype(np.float64), 0, 1024) * 2.0 ** 24)  # exact integer weights
    cum = np.cumsum(w)
    keep = depth_sorted[(cum >= cum[-1] * ignore_frac) & (cum <= cum[-1] * (1 - ignore_frac))]
    return keep[0], keep[-1]


CURVES = {
    "neg_log": lambda x: -np.log(x + EPS),
    "identity": lambda x: x,
    "inverse": lambda x: 1 / (x + EPS),
    "log": lambda x: np.log(x + EPS),
}


def visualize_depth(depth, acc=None, near=None, far=None, ignore_frac=0, curve="neg_log", modulus=0, lut=None):
    """pose.py:149-212 in float32 like the reference; lut None -> sinebow."""
    depth = depth.astype(np.float32)
    acc = np.ones_like(depth) if acc is None else acc.astype(np.float32)
    acc = np.where(np.isnan(depth), np.zeros_like(acc), acc)
    with np.errstate(all="ignore"):
        if not near or not far:
            lo, hi = auto_planes(depth, acc, ignore_frac)
            near = near or lo - EPS
            far = far or hi + EPS
        fn = CURVES[curve]
        d, cn, cf = fn(depth), fn(np.float32(near)), fn(np.float32(far))
        if modulus > 0:
            value = np.mod(d, np.float32(modulus)) / np.float32(modulus)
        else:
            value = np.nan_to_num(np.clip((d - np.minimum(cn, cf)) / np.abs(cf - cn), 0, 1))
        if lut is None:
            vis = sinebow(value)
        else:
            lut = np.asarray(lut, dtype=np.float32)
            idx = np.where(np.isnan(value), 0, value * np.float32(len(lut))).astype(np.int64)
            vis = lut[np.clip(idx, 0, len(lut) - 1)][..., :3]
        vis = vis * acc[:, :, None] + (1 - acc)[:, :, None]
    return vis.astype(np.float32)
