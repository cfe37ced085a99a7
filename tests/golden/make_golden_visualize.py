"""Generate tests/golden/visualize_golden.npz by running the LITERAL reference visualisers
(intern/pose.py visualize_normals / visualize_depth) on small synthetic frames.

Build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_visualize.py
matplotlib is not installed here; pose.py imports matplotlib.cm at module level but only touches it for the default
'turbo' colour map, so the module is stubbed and every depth case passes an explicit colour map (a listed table with
matplotlib's float-indexing rule, or the reference's own sinebow).
"""
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.cm"):
    try:
        __import__(name)
    except Exception:
        sys.modules[name] = types.ModuleType(name)
import matplotlib  # noqa: E402

matplotlib.cm = sys.modules["matplotlib.cm"]
from intern import pose as ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "visualize_golden.npz")
G = {}
rng = np.random.default_rng(7)


def frame(h, w, nan_pixels=0, lo=2.0, hi=6.0):
    yy, xx = np.mgrid[0:h, 0:w]
    depth = lo + (hi - lo) * (0.5 + 0.35 * np.sin(xx / 3.0) * np.cos(yy / 4.0)) + 0.05 * rng.normal(size=(h, w))
    depth = depth.astype(np.float32)
    acc = np.clip(rng.uniform(-0.1, 1.2, size=(h, w)), 0, 1).astype(np.float32)
    for _ in range(nan_pixels):
        depth[rng.integers(h), rng.integers(w)] = np.nan
    return depth, acc


def table_map(lut):
    """A listed colour map called with floats, as matplotlib evaluates it (colors.py Colormap.__call__)."""
    def f(value):
        xa = np.array(value, copy=True)
        xa = xa * len(lut)
        xa[xa == len(lut)] = len(lut) - 1
        return lut[np.clip(xa.astype(int), 0, len(lut) - 1)]
    return f


lut = rng.uniform(size=(256, 4)).astype(np.float32)
G["lut"] = lut

# normals
for tag, (h, w, nans, use_acc) in {"normals_a": (23, 31, 0, True), "normals_nan": (17, 40, 3, True),
                                   "normals_noacc": (9, 12, 0, False)}.items():
    depth, acc = frame(h, w, nans)
    G[f"{tag}/depth"], G[f"{tag}/acc"] = depth, acc
    G[f"{tag}/use_acc"] = np.array(use_acc)
    with np.errstate(all="ignore"):
        G[f"{tag}/vis"] = ref.visualize_normals(depth.copy(), acc.copy() if use_acc else None)

# depth
cases = {
    "depth_planes": dict(h=23, w=31, nans=0, near=2.0, far=6.0, ignore_frac=0, modulus=0, cmap="lut"),
    "depth_planes_nan": dict(h=16, w=19, nans=2, near=2.0, far=6.0, ignore_frac=0, modulus=0, cmap="lut"),
    "depth_auto_near": dict(h=20, w=27, nans=0, near=0, far=7.0, ignore_frac=0, modulus=0, cmap="lut"),  # LLFF: near = 0
    "depth_auto_both": dict(h=25, w=33, nans=0, near=None, far=None, ignore_frac=0.1, modulus=0, cmap="lut"),
    "depth_auto_frac_sinebow": dict(h=21, w=22, nans=0, near=None, far=None, ignore_frac=0.05, modulus=0, cmap="sinebow"),
    "depth_modulus": dict(h=18, w=29, nans=1, near=2.0, far=6.0, ignore_frac=0, modulus=0.25, cmap=None),
}
for tag, c in cases.items():
    depth, acc = frame(c["h"], c["w"], c["nans"])
    cmap = table_map(lut) if c["cmap"] == "lut" else ref.sinebow if c["cmap"] == "sinebow" else None
    with np.errstate(all="ignore"):
        vis = ref.visualize_depth(depth.copy(), acc.copy(), c["near"], c["far"], ignore_frac=c["ignore_frac"],
                                  modulus=c["modulus"], colormap=cmap)
    G[f"{tag}/depth"], G[f"{tag}/acc"], G[f"{tag}/vis"] = depth, acc, np.asarray(vis)
    G[f"{tag}/params"] = np.array([np.nan if c["near"] is None else c["near"], np.nan if c["far"] is None else c["far"],
                                   c["ignore_frac"], c["modulus"]], dtype=np.float64)
    G[f"{tag}/cmap"] = np.array(c["cmap"] or "sinebow")

np.savez_compressed(OUT, **G)
print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(G), "arrays")
