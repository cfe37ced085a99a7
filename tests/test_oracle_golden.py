"""Pins the CPU oracle against outputs of the literal reference (tests/golden/reference_golden.npz).

CPU only.  Tolerances: the oracle differs from the reference only by (a) the closed-form contraction
Jacobian vs. autograd's and (b) summation order in vectorised reductions, so fp32 agreement is at the
few-ulp level; stated per test.
"""
import torch

from conftest import rays_from
from oracle import mip360_oracle as O

RT = dict(rtol=2e-6, atol=1e-7)


def close(a, b, **kw):
    kw = {**RT, **kw}
    torch.testing.assert_close(a, b, **kw)


def test_scalar_helpers(golden):
    c = golden.case("g")
    close(O.g(c["x"]), c["out"], rtol=0, atol=0)
    c = golden.case("t_to_s")
    s, _ = O.t_to_s(c["t"], c["near"], c["far"])
    close(s, c["out"], rtol=0, atol=0)
    c = golden.case("s_to_t")
    close(O.s_to_t(c["s"], c["near"], c["far"]), c["out"], rtol=0, atol=0)


def test_sample_along_rays(golden):
    for name in ("sample_det", "sample_rand", "sample_rand_n64", "sample_jac", "sample_small"):
        c = golden.case(name)
        rays = rays_from(c)
        t, (mean, cov) = O.sample_along_rays(rays.origins, rays.directions, rays.radii, int(c["N"]), rays.near,
                                             rays.far, bool(c["randomized"]), c["t_rand"])
        close(t, c["t_vals"], rtol=0, atol=0)
        close(mean, c["means"])
        # covariance entries span many orders of magnitude inside one 3x3 block; tolerance relative to the block norm
        scale = c["covs"].flatten(-2).norm(dim=-1)[..., None, None]
        assert ((cov - c["covs"]).abs() <= 3e-6 * scale + 1e-30).all(), name
    # the Jacobian case really has J != I
    c = golden.case("sample_jac")
    rays = rays_from(c)
    t_mean, _, _ = O.frustum_moments(c["t_vals"][..., :-1], c["t_vals"][..., 1:], rays.radii)
    mean_c, _ = O.gaussian_contract(rays.directions[:, None, :] * t_mean[..., None], torch.zeros(3, 4, 3, 3))
    assert (mean_c.norm(dim=-1) > 1).any()


def test_piecewise_constant_pdf(golden):
    for name in ("pdf_det", "pdf_rand", "pdf_rand_n64", "pdf_tiny"):
        c = golden.case(name)
        out = O.sorted_piecewise_constant_pdf(c["bins"], c["weights"], int(c["M"]), bool(c["randomized"]), c["jitter"])
        close(out, c["samples"], rtol=0, atol=0)  # bit-equal (App. B2)


def test_resample_along_rays(golden):
    for name in ("resample_det", "resample_rand"):
        c = golden.case(name)
        rays = rays_from(c)
        t, (mean, cov) = O.resample_along_rays(rays.origins, rays.directions, rays.radii, c["t_in"], c["weights"],
                                               bool(c["randomized"]), 0.01, c["jitter"])
        close(t, c["t_vals"], rtol=0, atol=0)
        close(mean, c["means"])
        scale = c["covs"].flatten(-2).norm(dim=-1)[..., None, None]
        assert ((cov - c["covs"]).abs() <= 3e-6 * scale + 1e-30).all()


def test_encodings(golden):
    c = golden.case("ipe")
    close(O.integrated_pos_enc(c["mean"], c["cov"]), c["enc"], rtol=1e-5, atol=1e-6)
    c = golden.case("viewdir")
    close(O.viewdir_enc(c["viewdirs"]), c["enc"], rtol=0, atol=0)


def test_compositing(golden):
    for wb in (0, 1):
        c = golden.case(f"render_wb{wb}")
        rgb, dist, acc, w = O.volumetric_rendering(c["rgb"], c["density"], c["t_vals"], c["dirs"], bool(wb))
        close(w, c["weights"], rtol=0, atol=0)
        close(rgb, c["comp_rgb"], rtol=0, atol=0)
        close(dist, c["distance"], rtol=0, atol=0)
        close(acc, c["acc"], rtol=0, atol=0)
    c = golden.case("density_to_weight")
    close(O.density_to_weight(c["t_vals"], c["density"][..., 0], c["dirs"]), c["weights"], rtol=0, atol=0)


def test_losses(golden):
    c = golden.case("interlevel")
    b = O.bounds(c["t_fine"], c["w_fine"], c["t_coarse"])
    close(b, c["bounds"], rtol=1e-6)
    close(O.loss_prop(c["w_coarse"], c["bounds"]), c["loss_prop"], rtol=1e-6)
    close(O.Loss_prop(c["t_fine"], c["w_fine"], c["t_coarse"], c["w_coarse"]), c["Loss_prop"], rtol=1e-6)
    c = golden.case("distortion")
    close(O.loss_dist(c["s_vals"], c["weights"]), c["loss"], rtol=2e-6)
    close(O.loss_dist_quadratic(c["s_vals"], c["weights"]), c["loss"], rtol=2e-6)
    c = golden.case("loss_nerf")
    ln, psnr = O.Loss_nerf(c["input"], c["target"])
    close(ln, c["loss"], rtol=0, atol=0)
    close(psnr, c["psnr"], rtol=0, atol=0)


def _sd(golden):
    return golden.case("state_dict")


def test_model_forward_and_grads(golden):
    sd = _sd(golden)
    for randomized in (0, 1):
        c = golden.case(f"model_rand{randomized}")
        rays = rays_from(c)
        N = int(c["N"])
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        t_hat, w_hat = O.prop_forward(params, rays, N, bool(randomized), t_rand=c["t_rand"])
        close(t_hat, c["t_hat"], rtol=0, atol=0)
        close(w_hat, c["w_hat"], rtol=1e-5)
        rgb, dist, acc, t_f, w_f, s_f = O.nerf_forward(params, rays, c["t_hat"], c["w_hat"], bool(randomized),
                                                       jitter=c["jitter"])
        close(t_f, c["t_fine"], rtol=0, atol=0)
        close(w_f, c["w_fine"], rtol=1e-5)
        close(s_f, c["s_fine"], rtol=0, atol=0)
        close(rgb, c["rgb"], rtol=1e-5)
        close(dist, c["dist"], rtol=1e-5)
        close(acc, c["acc"], rtol=1e-5)
        lp = O.Loss_prop(t_f.detach(), w_f.detach(), t_hat, w_hat)
        close(lp, c["loss_prop"], rtol=1e-5)
        ln, psnr = O.Loss_nerf(rgb, c["pixels"])
        ld = O.loss_dist(s_f, w_f)
        close(ln, c["loss_nerf"], rtol=1e-5)
        close(ld, c["loss_dist"], rtol=1e-5)
        names_p = [k for k in params if k.startswith("prop_net")]
        names_n = [k for k in params if k.startswith("nerf_net")]
        gp = torch.autograd.grad(lp, [params[k] for k in names_p], retain_graph=True)
        gn = torch.autograd.grad(ln + 0.01 * ld, [params[k] for k in names_n])
        for k, gr in list(zip(names_p, gp)) + list(zip(names_n, gn)):
            ref = c["grad." + k]
            assert (gr - ref).abs().max() <= 2e-5 * ref.abs().max() + 1e-9, k
        out = O.model_forward(sd, rays, N, bool(randomized), t_rand=c["t_rand"], jitter=c["jitter"])
        close(out[0], c["fwd_rgb"], rtol=1e-5)
        close(out[1], c["fwd_dist"], rtol=1e-5)
        close(out[2], c["fwd_acc"], rtol=1e-5)


def test_fp64_truth_is_close_to_fp32(golden):
    """The same oracle in fp64 agrees with the fp32 reference outputs to fp32 rounding."""
    c = golden.case("sample_rand_n64")
    rays = rays_from(c)
    t, (mean, cov) = O.sample_along_rays(*[x.double() for x in (rays.origins, rays.directions, rays.radii)],
                                         int(c["N"]), rays.near.double(), rays.far.double(), True, c["t_rand"].double())
    close(t.float(), c["t_vals"], rtol=1e-5)
    close(mean.float(), c["means"], rtol=1e-5, atol=1e-6)


def test_init_state_dict_matches_reference_layout(golden):
    keys = [str(k) for k in golden.case("state_dict_full_shapes")["keys"]]
    sd = O.init_state_dict(seed=0)
    assert list(sd.keys()) == keys
    assert sum(v.numel() for v in sd.values()) == 7_624_453
    small = O.init_state_dict(16, 32, seed=0)
    ref = _sd(golden)
    for k in ref:
        torch.testing.assert_close(small[k], ref[k], rtol=0, atol=0)


def test_raygen_oracle_equals_reference():
    """oracle/raygen_oracle.py against the literal dataset.py / ray.py generators (tests/golden/raygen_golden.npz)."""
    import os
    import numpy as np
    from oracle import raygen_oracle as R
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "raygen_golden.npz"))
    for tag, fn in (("pinhole", R.pinhole_rays), ("pinhole_wide", R.pinhole_rays), ("llff_ndc", R.llff_ndc_rays)):
        h, w, focal, near, far = z[tag + "/hwf"]
        out = fn(z[tag + "/c2w"], int(h), int(w), float(focal), float(near), float(far))
        for k, v in out.items():
            assert np.array_equal(v, z[f"{tag}/{k}"]), (tag, k)
    o, d = R.convert_to_ndc(z["ndc/origins"], z["ndc/directions"], 12.5, 20, 10)
    assert np.array_equal(o, z["ndc/out_origins"]) and np.array_equal(d, z["ndc/out_directions"])
