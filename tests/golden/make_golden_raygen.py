"""Generate tests/golden/raygen_golden.npz by running the LITERAL reference ray generators
(dataset.py NeRFDataset.generate_rays / LLFF.generate_rays, intern/ray.py convert_to_ndc).

Build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_raygen.py
matplotlib / imageio are not installed here and are only used by plotting helpers; they are stubbed so that
dataset.py imports.  No dataset files are read: the class instances are created without __init__ and given the
attributes generate_rays() needs.
"""
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "imageio"):
    try:
        __import__(name)
    except Exception:
        sys.modules[name] = types.ModuleType(name)
import matplotlib  # noqa: E402

matplotlib.pyplot = sys.modules["matplotlib.pyplot"]
matplotlib.cm = sys.modules["matplotlib.cm"]
import dataset as ref_ds  # noqa: E402
from intern import ray as ref_ray  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raygen_golden.npz")
G = {}
rng = np.random.default_rng(0)


def random_pose(n):
    out = []
    for _ in range(n):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        out.append(np.concatenate([q, rng.normal(size=(3, 1))], 1))
    return np.stack(out).astype(np.float32)


for tag, cls, h, w, focal, near, far in [("pinhole", ref_ds.NeRFDataset, 6, 8, 9.5, 2.0, 6.0),
                                         ("pinhole_wide", ref_ds.NeRFDataset, 5, 37, 30.25, 0.1, 10.0),
                                         ("llff_ndc", ref_ds.LLFF, 7, 9, 11.0, 0.0, 1.0)]:
    obj = object.__new__(cls)
    c2w = random_pose(2)
    if tag == "llff_ndc":  # forward-facing cameras behind the z = -1 plane, as the NDC mapping assumes
        c2w[:, :3, :3] = np.eye(3, dtype=np.float32) + 0.05 * rng.normal(size=(2, 3, 3)).astype(np.float32)
        c2w[:, :3, 3] = 0.1 * rng.normal(size=(2, 3)).astype(np.float32)
    obj.h, obj.w, obj.focal, obj.near, obj.far, obj.cam_to_world = h, w, focal, near, far, c2w
    obj.generate_rays()
    G[f"{tag}/c2w"] = c2w
    G[f"{tag}/hwf"] = np.array([h, w, focal, near, far], dtype=np.float64)
    for k in obj.rays._fields:
        G[f"{tag}/{k}"] = np.asarray(getattr(obj.rays, k))

o = rng.normal(size=(4, 5, 3)).astype(np.float32) - np.array([0, 0, 3], np.float32)
d = rng.normal(size=(4, 5, 3)).astype(np.float32)
no, nd = ref_ray.convert_to_ndc(o, d, 12.5, 20, 10)
G["ndc/origins"], G["ndc/directions"], G["ndc/out_origins"], G["ndc/out_directions"] = o, d, no, nd
np.savez_compressed(OUT, **G)
print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")
