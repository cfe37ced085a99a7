import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
GOLDEN_R2 = os.path.join(ROOT, "tests", "golden", "reference_golden_r2.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_sessionstart(session):
    """GPU box without the prebuilt library (e.g. a fresh checkout): build it once; nvcc is part of the image."""
    lib = os.path.join(ROOT, "mipnerf360_b200", "lib", "libmip360_b200.so")
    if torch.cuda.is_available() and not os.path.exists(lib):
        from mipnerf360_b200 import build
        build.build()


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Access to the committed outputs of the literal reference (tests/golden/make_golden.py)."""

    def __init__(self, path=GOLDEN):
        self._z = np.load(path, allow_pickle=False)

    def case(self, name, device="cpu"):
        out = {}
        pre = name + "/"
        for k in self._z.files:
            if k.startswith(pre):
                v = self._z[k]
                if v.dtype.kind in "fiub" and v.dtype.kind != "U":
                    t = torch.from_numpy(np.array(v))
                    out[k[len(pre):]] = t.to(device) if t.ndim > 0 else t
                else:
                    out[k[len(pre):]] = v
        return out


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def golden2():
    """Second fixture set (tests/golden/make_golden_r2.py): render_image, to8b, the train.py loop, rare branches."""
    return Golden(GOLDEN_R2)


def rays_from(case, device="cpu"):
    from oracle.mip360_oracle import Rays
    return Rays(*[case[k].to(device) for k in Rays._fields])


# ---------------------------------------------------------------------------------------------------
# measured parity margins: tests call parity_record(name, value); the session writes the worst value per name to
# gpurun_out/parity.json (copied to profiles/rNN_parity.json).  Tolerances in the tests are <= 2x these records.
# ---------------------------------------------------------------------------------------------------
_PARITY = {}


def parity_record(name, value):
    v = float(value)
    _PARITY[name] = max(_PARITY.get(name, 0.0), v)
    return v


def pytest_sessionfinish(session, exitstatus):
    if _PARITY and torch.cuda.is_available():
        import json
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        json.dump(dict(sorted(_PARITY.items())), open(os.path.join(out, "parity.json"), "w"), indent=1)
