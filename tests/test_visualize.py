"""Depth / normal visualisation (intern/pose.py:112-212).

CPU: oracle/visualize_oracle.py against the outputs of the literal reference (tests/golden/visualize_golden.npz).
GPU: the CUDA path (through the C ABI) against the same golden outputs, against the oracle on larger frames, and
property checks at render sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import visualize_oracle as V

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "visualize_golden.npz"))
NORMAL_CASES = ["normals_a", "normals_nan", "normals_noacc"]
DEPTH_CASES = ["depth_planes", "depth_planes_nan", "depth_auto_near", "depth_auto_both", "depth_auto_frac_sinebow",
               "depth_modulus"]
# float32 pipeline (log, sin, normalisation): a few ulp of [0,1] quantities; normals are fp64 on both sides
TOL_DEPTH = 2e-5
TOL_NORMALS = 1e-6


def _depth_args(tag):
    near, far, frac, modulus = Z[tag + "/params"]
    cmap = str(Z[tag + "/cmap"])
    return dict(near=None if np.isnan(near) else float(near), far=None if np.isnan(far) else float(far),
                ignore_frac=float(frac), modulus=float(modulus)), cmap


def _u8(x):
    return (255 * np.clip(np.nan_to_num(x), 0, 1)).astype(np.uint8)


def _close_u8(a, b, frac=2e-3):
    """uint8 pictures: equal up to one level on at most `frac` of the values (truncation at integer boundaries)."""
    diff = np.abs(a.astype(np.int32) - b.astype(np.int32))
    assert diff.max() <= 1, diff.max()
    assert (diff > 0).mean() <= frac, (diff > 0).mean()


@pytest.mark.parametrize("tag", NORMAL_CASES)
def test_oracle_normals_equal_reference(tag):
    acc = Z[tag + "/acc"] if bool(Z[tag + "/use_acc"]) else None
    vis = V.visualize_normals(Z[tag + "/depth"], acc)
    np.testing.assert_allclose(vis, Z[tag + "/vis"], rtol=0, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("tag", DEPTH_CASES)
def test_oracle_depth_equals_reference(tag):
    kw, cmap = _depth_args(tag)
    vis = V.visualize_depth(Z[tag + "/depth"], Z[tag + "/acc"], lut=Z["lut"] if cmap == "lut" else None, **kw)
    np.testing.assert_allclose(vis, Z[tag + "/vis"], rtol=0, atol=1e-6, equal_nan=True)


def test_colour_tables_host_logic():
    """Host side of visualize_depth: what is accepted as a colour map and what the default table looks like."""
    from mipnerf360_b200.intern import pose as P
    lut = P.turbo_lut()
    assert lut.shape == (256, 3) and lut.dtype == np.float32 and lut.min() >= 0 and lut.max() <= 1
    assert lut[230, 0] > 2 * lut[230, 2] and lut[25, 2] > 2 * lut[25, 0]          # red at the near end, blue at the far end
    assert P._colour_table(None) is None and P._colour_table("sinebow") is None and P._colour_table(P.sinebow) is None
    rgba = np.random.default_rng(0).uniform(size=(17, 4)).astype(np.float32)
    for given in (rgba, rgba.tolist(), torch.as_tensor(rgba)):
        assert np.array_equal(P._colour_table(given), rgba[:, :3])

    class Listed:                       # duck-typed matplotlib ListedColormap
        colors = rgba

    class Segmented:                    # duck-typed matplotlib Colormap: integer input indexes its table
        N = 17

        def __call__(self, idx):
            return rgba[idx]

    assert np.array_equal(P._colour_table(Listed()), rgba[:, :3])
    assert np.array_equal(P._colour_table(Segmented()), rgba[:, :3])
    with pytest.raises(TypeError):
        P._colour_table(lambda v: v)
    with pytest.raises(ValueError):
        P._colour_table(np.zeros((5, 2)))
    # the reference's sinebow (pose.py:123-126)
    h = np.linspace(0, 1, 11, dtype=np.float32)
    np.testing.assert_allclose(P.sinebow(h).numpy(), V.sinebow(h), atol=1e-6)
    with pytest.raises(TypeError):
        P.visualize_depth(np.ones((4, 4), np.float32), None, 1.0, 2.0, curve_fn=lambda x: x)
    if not torch.cuda.is_available():   # no device: the product path fails loudly, there is no host fallback
        with pytest.raises(Exception):
            P.visualize_depth(np.ones((4, 4), np.float32), None, 1.0, 2.0)


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def pose():
    from mipnerf360_b200.intern import pose as P
    return P


@pytest.mark.gpu
@pytest.mark.parametrize("tag", NORMAL_CASES)
def test_gpu_normals_vs_reference(pose, tag):
    acc = Z[tag + "/acc"] if bool(Z[tag + "/use_acc"]) else None
    vis = pose.visualize_normals(Z[tag + "/depth"], acc)
    assert isinstance(vis, np.ndarray) and vis.shape == Z[tag + "/vis"].shape
    np.testing.assert_allclose(vis, Z[tag + "/vis"], rtol=0, atol=TOL_NORMALS, equal_nan=True)
    _close_u8(pose.visualize_normals(Z[tag + "/depth"], acc, as_uint8=True), _u8(Z[tag + "/vis"]))
    assert pose.visualize_normals(Z[tag + "/depth"], acc, scaling=1.0) is None  # the reference's quirk (pose.py:130)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", DEPTH_CASES)
def test_gpu_depth_vs_reference(pose, tag):
    kw, cmap = _depth_args(tag)
    colormap = Z["lut"] if cmap == "lut" else ("sinebow" if kw["modulus"] == 0 else None)
    vis = pose.visualize_depth(Z[tag + "/depth"], Z[tag + "/acc"], colormap=colormap, **kw)
    np.testing.assert_allclose(vis, Z[tag + "/vis"], rtol=0, atol=TOL_DEPTH, equal_nan=True)
    _close_u8(pose.visualize_depth(Z[tag + "/depth"], Z[tag + "/acc"], colormap=colormap, as_uint8=True, **kw),
              _u8(Z[tag + "/vis"]), frac=5e-3)


@pytest.mark.gpu
def test_gpu_visualisers_vs_oracle_at_frame_size(pose):
    """A 756 x 1008 frame (BASELINE render config) with NaN pixels and partial accumulation."""
    from mipnerf360_b200 import ops
    rng = np.random.default_rng(3)
    h, w = 756, 1008
    yy, xx = np.mgrid[0:h, 0:w]
    depth = (3 + np.sin(xx / 40.0) * np.cos(yy / 55.0) + 0.02 * rng.normal(size=(h, w))).astype(np.float32)
    depth[rng.integers(h, size=50), rng.integers(w, size=50)] = np.nan
    acc = np.clip(rng.uniform(-0.2, 1.3, size=(h, w)), 0, 1).astype(np.float32)
    # normals: scaling statistics and picture
    stats = ops.normals_scaling(torch.as_tensor(depth).cuda()).cpu().numpy()
    assert stats[0] == np.count_nonzero(~np.isnan(depth))
    np.testing.assert_allclose(stats[7], V.normals_scaling(depth), rtol=2e-6)  # the reference's variance is float32
    np.testing.assert_allclose(pose.visualize_normals(depth, acc), V.visualize_normals(depth, acc), rtol=0, atol=TOL_NORMALS)
    # depth: automatic planes by weighted quantile == sort-based oracle, several ignore fractions
    for frac in (0.0, 0.01, 0.1, 0.3):
        rngd = ops.depth_range(torch.as_tensor(depth).cuda(), torch.as_tensor(acc).cuda(), None, None, frac).cpu().numpy()
        a = np.where(np.isnan(depth), 0, acc)
        lo, hi = V.auto_planes(depth, a, frac)
        expect = np.array([lo - V.EPS, hi + V.EPS], dtype=np.float32)
        assert np.array_equal(rngd, expect, equal_nan=True), (frac, rngd, expect)
    lut = rng.uniform(size=(256, 3)).astype(np.float32)
    d_ok = np.nan_to_num(depth, nan=3.0)  # NaN depths make far = NaN at ignore_frac = 0 (np.argsort puts NaN last)
    for kw in (dict(near=2.0, far=4.5), dict(near=None, far=None, ignore_frac=0.05), dict(near=0, far=4.5),
               dict(near=2.0, far=4.5, modulus=0.2), dict(near=None, far=None, curve="identity")):
        okw = dict(kw)
        gkw = dict(kw)
        if "curve" in gkw:
            gkw["curve_fn"] = gkw.pop("curve")
        use_lut = "modulus" not in kw
        ref = V.visualize_depth(d_ok, acc, lut=lut if use_lut else None, **okw)
        got = pose.visualize_depth(d_ok, acc, colormap=lut if use_lut else None, **gkw)
        if use_lut:
            # a table lookup amplifies a 1-ulp difference in `value` to a different entry on a few pixels
            bad = (np.abs(got - ref) > TOL_DEPTH).any(-1).mean()
            assert bad < 2e-3, (kw, bad)
        else:
            np.testing.assert_allclose(got, ref, rtol=0, atol=TOL_DEPTH)
    # NaN depth with fully automatic planes: the reference's far plane becomes NaN -> every value maps to table[0]
    got = pose.visualize_depth(depth, acc, colormap=lut)
    a = np.where(np.isnan(depth), 0, acc)[..., None]
    np.testing.assert_allclose(got, lut[0] * a + (1 - a), rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_gpu_visualisers_properties_and_errors(pose):
    from mipnerf360_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(5)
    depth = torch.rand(300, 500, device="cuda", generator=g) * 4 + 2
    acc = torch.rand(300, 500, device="cuda", generator=g)
    # tensors in -> tensors out, on the device
    out = pose.visualize_depth(depth, acc, 2.0, 6.0)
    assert isinstance(out, torch.Tensor) and out.is_cuda and out.shape == (300, 500, 3)
    # default turbo table: inside [0,1], near is red-ish, far is blue-ish (docstring of the reference, pose.py:165-167)
    ones = torch.ones(1, 2, device="cuda")
    rgb = pose.visualize_depth(torch.tensor([[6 / 3 ** 0.9, 6 / 3 ** 0.1]], device="cuda"), ones, 2.0, 6.0)  # values 0.9, 0.1
    assert rgb[0, 0, 0] > 2 * rgb[0, 0, 2] and rgb[0, 1, 2] > 2 * rgb[0, 1, 0]
    assert float(out.min()) >= 0 and float(out.max()) <= 1
    # acc = 0 -> white, whatever the depth
    white = pose.visualize_depth(depth, torch.zeros_like(acc), 2.0, 6.0)
    assert torch.equal(white, torch.ones_like(white))
    white = pose.visualize_normals(depth, torch.zeros_like(acc))
    assert torch.equal(white, torch.ones_like(white))
    # a planar depth ramp has constant normals away from the border
    yy, xx = torch.meshgrid(torch.arange(64.0, device="cuda"), torch.arange(96.0, device="cuda"), indexing="ij")
    n = pose.visualize_normals(0.5 * xx + 0.25 * yy, None)
    inner = n[1:-1, 1:-1].reshape(-1, 3)
    assert float((inner - inner[0]).abs().max()) < 1e-6
    # uint8 output equals to8b of the float output
    from mipnerf360_b200 import ops
    assert torch.equal(pose.visualize_depth(depth, acc, 2.0, 6.0, as_uint8=True), ops.to8b(out))
    # buffers that are contiguous but not 16-byte aligned take the scalar-load path: same pictures
    flat_d, flat_a = torch.cat([depth.flatten(), depth.new_zeros(4)]), torch.cat([acc.flatten(), acc.new_zeros(4)])
    d1 = flat_d.roll(1)[1:1 + depth.numel()].view_as(depth)
    a1 = flat_a.roll(1)[1:1 + acc.numel()].view_as(acc)
    assert d1.data_ptr() % 16 == 4 and torch.equal(d1, depth)
    assert torch.equal(pose.visualize_depth(d1, a1, 2.0, 6.0), out)
    assert torch.equal(ops.depth_range(d1, a1, None, None, 0.1), ops.depth_range(depth, acc, None, None, 0.1))
    x = torch.randn(1001, device="cuda", generator=g)
    assert torch.equal(ops.to8b(x.roll(1)[1:]), ops.to8b(x[:-1].clone()))
    with pytest.raises(TypeError):
        pose.visualize_depth(depth, acc, 2.0, 6.0, curve_fn=lambda x: x)
    with pytest.raises(TypeError):
        pose.visualize_depth(depth, acc, 2.0, 6.0, colormap=lambda v: v)
    with pytest.raises(_lib.Mip360Error):
        pose.visualize_depth(depth, acc, None, None, ignore_frac=0.7)
