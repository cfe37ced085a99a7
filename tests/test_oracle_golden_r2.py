"""Pins the CPU oracle against the second set of literal-reference outputs (tests/golden/reference_golden_r2.npz,
made by tests/golden/make_golden_r2.py): to8b, render_image, the train.py loop with AdamW + lr_decay, the
non-default branches (stable=False, diag=True, plain positional encoding, other view-direction degrees) and rays
with more than 128 samples.  CPU only."""
import numpy as np
import torch

from conftest import rays_from
from oracle import mip360_oracle as O

RT = dict(rtol=2e-6, atol=1e-7)


def close(a, b, **kw):
    torch.testing.assert_close(a, b, **{**RT, **kw})


def cov_close(cov, ref, rel=3e-6):
    scale = ref.flatten(-2).norm(dim=-1)[..., None, None]
    assert ((cov - ref).abs() <= rel * scale + 1e-30).all()


def sd_of(golden2, case):
    return golden2.case(case)


def test_to8b(golden2):
    c = golden2.case("to8b")
    for x, y in (("x2", "y2"), ("x3", "y3"), ("k", "yk")):
        assert np.array_equal(O.to8b(c[x].numpy()), c[y].numpy())


def test_render_image(golden2):
    c = golden2.case("render_image")
    sd = sd_of(golden2, "render_image_sd")
    rgb8, rgb, dist, acc = O.render_image(sd, rays_from(c), int(c["H"]), int(c["W"]), int(c["N"]), chunks=int(c["chunks"]))
    close(rgb, c["rgb_float"], rtol=1e-5, atol=1e-6)
    close(dist, c["dists"], rtol=1e-5, atol=1e-6)
    close(acc, c["accs"], rtol=1e-5, atol=1e-6)
    assert np.abs(rgb8.astype(int) - c["rgb8"].numpy().astype(int)).max() <= 1  # truncation of values 1e-6 apart
    # each chunk is its own batch for the contraction norm (App. A1): one big chunk gives a different image
    _, rgb_one, _, _ = O.render_image(sd, rays_from(c), int(c["H"]), int(c["W"]), int(c["N"]), chunks=10**6)
    assert (rgb_one - c["rgb_float"]).abs().max() > 1e-5


def test_train_loop(golden2):
    c = golden2.case("train_loop")
    sd0 = sd_of(golden2, "train_loop_sd0")
    params, log = O.train_loop(sd0, rays_from(c), c["pixels"], int(c["N"]), iterations=3)
    ref_log = c["log"].numpy()
    log = np.array(log)
    # the reference drifts near/far by 1e-6 per forward inside an iteration (App. A4); the oracle is pure
    np.testing.assert_allclose(log[:, :5], ref_log[:, :5], rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(log[:, 5], ref_log[:, 5], rtol=1e-12)  # learning rate after 3, 6, 9 scheduler steps
    sd3 = sd_of(golden2, "train_loop_sd3")
    lr = 2e-4  # warm-up learning rate; an AdamW step moves a weight by at most ~lr
    for k in sd3:
        d = (params[k] - sd3[k]).abs()
        moved = (sd3[k] - sd0[k]).abs().max()
        assert moved > 0, k
        assert float(d.max()) <= 2.5 * lr and float(d.mean()) <= 0.05 * lr, (k, float(d.max()), float(d.mean()))
    o = golden2.case("train_loop_optim")
    assert sorted(set(o["steps"].tolist())) == [3.0, 6.0]  # nerf parameters stepped 3 times, proposal ones 6 (2 per iteration)


def test_frustum_branches(golden2):
    for stable in (True, False):
        c = golden2.case(f"frustum_stable{int(stable)}")
        mean, cov = O.conical_frustum_to_gaussian(c["d"], c["t0"], c["t1"], c["radii"], stable=stable)
        close(mean, c["mean"], rtol=1e-5 if not stable else 2e-6)
        cov_close(cov, c["cov"], rel=3e-6 if stable else 2e-4)  # the unstable formula cancels catastrophically by design
    c = golden2.case("gaussian_to_xyz")
    mean, cov = O.gaussian_to_xyz(c["d"], c["t_mean"], c["t_var"], c["r_var"], diag=True)
    close(mean, c["mean_diag"], rtol=0, atol=0)
    close(cov, c["cov_diag"])
    mean, cov = O.gaussian_to_xyz(c["d"], c["t_mean"], c["t_var"], c["r_var"])
    close(cov, c["cov_full"])
    close(torch.diagonal(c["cov_full"], dim1=-2, dim2=-1), c["cov_diag"])
    c = golden2.case("pos_enc_plain")
    close(O.pos_enc(c["mean"]), c["enc"])


def test_viewdir_degrees(golden2):
    for lo, hi in ((0, 4), (1, 3), (0, 6), (2, 3)):
        c = golden2.case(f"viewdir_{lo}_{hi}")
        enc = O.viewdir_enc(c["viewdirs"], lo, hi)
        assert enc.shape[-1] == 4 * (hi - lo)
        close(enc, c["enc"], rtol=1e-5, atol=2e-6)
    c = golden2.case("model_vd13")
    sd = sd_of(golden2, "model_vd13_sd")
    assert sd["prop_net.model.0.weight"].shape[1] == 42 + 8
    rgb, dist, acc = O.model_forward(sd, rays_from(c), int(c["N"]), False, viewdir_deg=(1, 3))
    close(rgb, c["fwd_rgb"], rtol=1e-5, atol=1e-6)
    close(dist, c["fwd_dist"], rtol=1e-5, atol=1e-6)
    close(acc, c["fwd_acc"], rtol=1e-5, atol=1e-6)


def test_more_than_128_samples(golden2):
    c = golden2.case("n150_sample")
    rays = rays_from(c)
    t, (mean, cov) = O.sample_along_rays(rays.origins, rays.directions, rays.radii, int(c["N"]), rays.near, rays.far, True,
                                         c["t_rand"])
    close(t, c["t_vals"], rtol=0, atol=0)
    close(mean, c["means"])
    cov_close(cov, c["covs"])
    c = golden2.case("n150_pdf")
    close(O.sorted_piecewise_constant_pdf(c["bins"], c["weights"], int(c["M"]), True, c["jitter"]), c["samples"], rtol=0, atol=0)
    c = golden2.case("n150_resample")
    rays = rays_from(c)
    t, (mean, cov) = O.resample_along_rays(rays.origins, rays.directions, rays.radii, c["t_in"], c["weights"], True, 0.01,
                                           c["jitter"])
    close(t, c["t_vals"], rtol=0, atol=0)
    close(mean, c["means"])
    cov_close(cov, c["covs"])
    c = golden2.case("n150_render")
    comp, dist, acc, w = O.volumetric_rendering(c["rgb"], c["density"], c["t_vals"], c["dirs"], True)
    for a, b in ((comp, "comp_rgb"), (dist, "distance"), (acc, "acc"), (w, "weights")):
        close(a, c[b], rtol=1e-5, atol=1e-6)
    c = golden2.case("n150_interlevel")
    close(O.bounds(c["t_fine"], c["w_fine"], c["t_coarse"]), c["bounds"], rtol=1e-5)
    close(O.Loss_prop(c["t_fine"], c["w_fine"], c["t_coarse"], c["w_coarse"]), c["Loss_prop"], rtol=1e-5)
    c = golden2.case("n150_distortion")
    close(O.loss_dist(c["s_vals"], c["weights"]), c["loss"], rtol=2e-5)


def test_literal_contract_loop_equals_closed_form():
    """The per-sample autograd-Jacobian loop of parameterization.py:64-83 (restated as written) against the closed form
    the oracle and the kernels use, including samples whose Jacobian is not the identity."""
    g = torch.Generator().manual_seed(0)
    mean = torch.randn(3, 5, 3, generator=g) * 0.4
    mean[0, 0] = torch.tensor([30.0, -10.0, 5.0])  # dominates the global norm: stays outside the unit ball after scaling
    A = torch.randn(3, 5, 3, 3, generator=g) * 0.2
    cov = A @ A.transpose(-1, -2)
    m_lit, c_lit = O.gaussian_contract_literal(mean.clone(), cov.clone())
    m_cf, c_cf = O.gaussian_contract(mean, cov)
    assert (m_cf.norm(dim=-1) > 1).any()
    close(m_cf, m_lit, rtol=0, atol=0)
    cov_close(c_cf, c_lit, rel=3e-6)
