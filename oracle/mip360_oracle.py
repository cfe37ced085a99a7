"""CPU oracle for the per-ray hot path of zhangkai0425/mipnerf360.

TEST INFRASTRUCTURE ONLY.  This file is a vectorised restatement, in plain torch CPU
ops, of the reference's algorithm.  It is the checker for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
it.  Nothing under ``mipnerf360_b200/`` imports it and the product path has no CPU
fallback.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced by importing
``/root/reference`` in the build container (``tests/golden/make_golden.py``) and
committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function below against them.

Every function cites the reference file:line it follows (paths relative to
``/root/reference``).  All functions are pure (the reference mutates several inputs in
place; where a single reference call observes its own mutation the
arithmetic is replicated explicitly) and dtype-generic, so the same code yields an fp64
"truth" when fed doubles.
"""
from __future__ import annotations

import math
from collections import namedtuple

import torch
import torch.nn.functional as F

Rays = namedtuple("Rays", ("origins", "directions", "viewdirs", "radii", "near", "far"))  # intern/ray.py:6

G_EPS = 1e-6

# 21 unit directions of intern/encoding.py:9-30 (icosahedron-derived basis), row order = column order of the encoding
_A, _B, _C, _D, _E = 0.8506508, 0.5257311, 0.809017, 0.5, 0.309017
IPE_BASIS = (
    (_A, 0, _B), (_C, _D, _E), (_B, _A, 0), (1, 0, 0), (_C, _D, -_E), (_A, 0, -_B), (_E, _C, -_D),
    (0, _B, -_A), (_D, _E, -_C), (0, 1, 0), (-_B, _A, 0), (-_E, _C, -_D), (0, _B, _A), (-_E, _C, _D),
    (_E, _C, _D), (_D, _E, _C), (_D, -_E, _C), (0, 0, 1), (-_D, _E, _C), (-_C, _D, _E), (-_C, _D, -_E),
)


def basis(dtype=torch.float32, device="cpu"):
    return torch.tensor(IPE_BASIS, dtype=dtype, device=device)


# ----------------------------------------------------------------------------------------------
# intern/parameterization.py
# ----------------------------------------------------------------------------------------------
def g(x):
    """intern/parameterization.py:15-21 — 1/(x+1e-6); pure (the reference adds eps in place)."""
    return 1.0 / (x + G_EPS)


def t_to_s(t_vals, near, far):
    """intern/parameterization.py:5-8.

    One reference call evaluates g(near) twice and each evaluation shifts ``near`` in place,
    so the numerator sees near+1e-6 and the denominator near+2e-6.  Returned alongside is the
    shifted t (the reference's ``t_vals`` argument after the call)."""
    t1 = t_vals + G_EPS
    n1 = near + G_EPS
    f1 = far + G_EPS
    n2 = n1 + G_EPS
    s = (1.0 / t1 - 1.0 / n1) / (1.0 / f1 - 1.0 / n2)
    return s, t1


def s_to_t(s_vals, near, far):
    """intern/parameterization.py:10-13 (g(far) first, then g(near), then the outer g)."""
    return g(s_vals * g(far) + (1 - s_vals) * g(near))


def contract(x):
    """intern/parameterization.py:23-29 — norm over the WHOLE tensor (App. A1)."""
    n = torch.linalg.vector_norm(x)
    if n <= 1:
        return x
    return (2 - 1 / n) * (x / n)


def contract_jacobian(x):
    """Closed form of jacobian(contract, x) for 3-vectors x[..., 3], as evaluated at
    intern/parameterization.py:77-79 (per-point norm; identity inside the unit ball)."""
    m = torch.linalg.vector_norm(x, dim=-1)[..., None, None]
    eye = torch.eye(3, dtype=x.dtype, device=x.device).expand(x.shape[:-1] + (3, 3))
    outer = x[..., :, None] * x[..., None, :]
    ms = torch.clamp(m, min=1.0)  # keeps the unused branch finite
    jac = (2 / ms - 1 / ms**2) * eye + (-2 / ms**3 + 2 / ms**4) * outer
    return torch.where(m <= 1, eye, jac)


def gaussian_to_xyz(d, t_mean, t_var, r_var, diag=False):
    """intern/parameterization.py:31-62: full covariance [B,N,3,3] (the branch the model uses) or, with diag=True,
    its diagonal [B,N,3] (:49-53)."""
    mean = d[..., None, :] * t_mean[..., None]
    d_mag_sq = torch.clamp(torch.sum(d**2, dim=-1, keepdim=True), min=1e-10)
    if diag:
        d_outer_diag = d**2
        null_outer_diag = 1 - d_outer_diag / d_mag_sq
        return mean, t_var[..., None] * d_outer_diag[..., None, :] + r_var[..., None] * null_outer_diag[..., None, :]
    d_outer = d[..., :, None] * d[..., None, :]
    eye = torch.eye(3, dtype=d.dtype, device=d.device)
    null_outer = eye - d[..., :, None] * (d / d_mag_sq)[..., None, :]
    cov = t_var[..., None, None] * d_outer[..., None, :, :] + r_var[..., None, None] * null_outer[..., None, :, :]
    return mean, cov


def gaussian_contract(mean, cov, norm_sq=None):
    """intern/parameterization.py:64-83 with the B*N autograd-Jacobian loop in closed form.

    ``norm_sq`` overrides the global Frobenius norm^2 (used to check a kernel given the same scalar)."""
    if norm_sq is None:
        n = torch.linalg.vector_norm(mean)
    else:
        n = torch.sqrt(torch.as_tensor(norm_sq, dtype=mean.dtype))
    mean_c = mean if n <= 1 else (2 - 1 / n) * (mean / n)
    jac = contract_jacobian(mean_c)
    cov_c = torch.matmul(torch.matmul(jac, cov), jac.transpose(-1, -2))
    return mean_c, cov_c


def gaussian_contract_literal(mean, cov):
    """intern/parameterization.py:64-83 AS WRITTEN: one torch.autograd.functional.jacobian call per sample in a double
    Python loop (the reference's dominant cost, SURVEY §6: 0.15-0.18 ms per call).  Used to time the literal algorithm
    (bench.py's CPU baseline) and to pin the closed form of `gaussian_contract` against it; O(B*N) Python iterations."""
    from torch.autograd.functional import jacobian
    mean = contract(mean)
    jf = torch.zeros((mean.shape[0], mean.shape[1], 3, 3), dtype=mean.dtype)
    for b in range(mean.shape[0]):
        for n in range(mean.shape[1]):
            jf[b][n] = jacobian(contract, mean[b][n])
    cov = torch.matmul(jf, cov)
    cov = torch.matmul(cov, jf.transpose(2, 3))
    return mean, cov


LITERAL_CONTRACT = False  # set by bench.py to time the reference's per-sample Jacobian loop instead of the closed form


def frustum_moments(t0, t1, radii, stable=True):
    """intern/parameterization.py:100-113: the stable formula of the mip-NeRF paper, or the original one."""
    if not stable:
        t_mean = (3 * (t1**4 - t0**4)) / (4 * (t1**3 - t0**3))
        r_var = radii**2 * (3 / 20 * (t1**5 - t0**5) / (t1**3 - t0**3))
        t_mosq = 3 / 5 * (t1**5 - t0**5) / (t1**3 - t0**3)
        return t_mean, t_mosq - t_mean**2, r_var
    mu = (t0 + t1) / 2
    hw = (t1 - t0) / 2
    t_mean = mu + (2 * mu * hw**2) / (3 * mu**2 + hw**2)
    t_var = (hw**2) / 3 - (4 / 15) * ((hw**4 * (12 * mu**2 - hw**2)) / (3 * mu**2 + hw**2) ** 2)
    r_var = radii**2 * ((mu**2) / 4 + (5 / 12) * hw**2 - 4 / 15 * (hw**4) / (3 * mu**2 + hw**2))
    return t_mean, t_var, r_var


def conical_frustum_to_gaussian(d, t0, t1, radii, norm_sq=None, stable=True):
    """intern/parameterization.py:85-117 (diag=False; with diag=True the reference itself fails inside
    gaussian_contract: a [B,N,3] diagonal cannot be multiplied by the [B,N,3,3] Jacobians)."""
    t_mean, t_var, r_var = frustum_moments(t0, t1, radii, stable)
    mean, cov = gaussian_to_xyz(d, t_mean, t_var, r_var)
    if LITERAL_CONTRACT and norm_sq is None:
        with torch.no_grad():  # Jf is filled from detached jacobian() results in the reference too (:76-79)
            return gaussian_contract_literal(mean, cov)
    return gaussian_contract(mean, cov, norm_sq)


def frustum_norm_sq(t_vals, directions, radii=None):
    """Sum over all rays/samples of |d * t_mean|^2 = the squared Frobenius norm that
    contract() sees at intern/parameterization.py:75 (accumulated in fp64)."""
    t_mean, _, _ = frustum_moments(t_vals[..., :-1].double(), t_vals[..., 1:].double(), 0.0)
    mean = directions.double()[..., None, :] * t_mean[..., None]
    return float((mean.to(t_vals.dtype).double() ** 2).sum())


def para_rays(t_vals, origins, directions, radii, norm_sq=None):
    """intern/parameterization.py:119-136 — origins are added AFTER the contraction (App. A2)."""
    means, covs = conical_frustum_to_gaussian(directions, t_vals[..., :-1], t_vals[..., 1:], radii, norm_sq)
    return means + origins[..., None, :], covs


# ----------------------------------------------------------------------------------------------
# intern/ray.py
# ----------------------------------------------------------------------------------------------
def level0_t_vals(near, far, num_samples, randomized, t_rand=None):
    """intern/ray.py:100-111.  ``t_rand`` = the torch.rand(B, N+1) draw of ray.py:106."""
    s = torch.linspace(0.0, 1, num_samples + 1, dtype=near.dtype, device=near.device)
    t = g(s * g(far) + (1 - s) * g(near))
    if randomized:
        mids = 0.5 * (t[..., 1:] + t[..., :-1])
        upper = torch.cat([mids, t[..., -1:]], -1)
        lower = torch.cat([t[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(near.shape[0], num_samples + 1, dtype=near.dtype, device=near.device)
        t = lower + (upper - lower) * t_rand
    else:
        t = t.expand(near.shape[0], num_samples + 1).clone()
    return t


def sample_along_rays(origins, directions, radii, num_samples, near, far, randomized, t_rand=None):
    """intern/ray.py:81-116."""
    t = level0_t_vals(near, far, num_samples, randomized, t_rand)
    return t, para_rays(t, origins, directions, radii)


def blur_weights(weights, resample_padding):
    """intern/ray.py:137-142."""
    w_pad = torch.cat([weights[..., :1], weights, weights[..., -1:]], dim=-1)
    w_max = torch.maximum(w_pad[..., :-1], w_pad[..., 1:])
    return 0.5 * (w_max[..., :-1] + w_max[..., 1:]) + resample_padding


def pdf_to_cdf(weights):
    """intern/ray.py:15-27: pad so the sum is >= 1e-5, normalise, cumsum, clamp, bracket with 0 and 1."""
    eps = 1e-5
    wsum = torch.sum(weights, dim=-1, keepdim=True)
    padding = torch.clamp(eps - wsum, min=0)
    w = weights + padding / weights.shape[-1]
    wsum = wsum + padding
    pdf = w / wsum
    cdf = torch.clamp(torch.cumsum(pdf[..., :-1], dim=-1), max=1)
    zeros = torch.zeros(cdf.shape[:-1] + (1,), dtype=cdf.dtype, device=cdf.device)  # N = 1: cdf is empty here
    return torch.cat([zeros, cdf, zeros + 1], dim=-1)


def pdf_uniforms(batch, num_samples, randomized, dtype=torch.float32, jitter=None, device=None):
    """intern/ray.py:29-39.  ``jitter`` = the uniform_(0, 1/M - eps) draw of ray.py:33.
    The stratum offset is doubled in the source (App. A5) and reproduced here."""
    eps32 = torch.finfo(torch.float32).eps
    device = jitter.device if jitter is not None else (device or "cpu")
    if randomized:
        s = 1 / num_samples
        u = (torch.arange(num_samples, device=device) * s).to(dtype)[None, :]
        if jitter is None:
            jitter = torch.empty(batch, num_samples, dtype=dtype, device=device).uniform_(to=(s - eps32))
        u = u + u + jitter
        return torch.minimum(u, torch.full_like(u, 1.0 - eps32))
    u = torch.linspace(0.0, 1.0 - eps32, num_samples, dtype=dtype, device=device)
    return u.expand(batch, num_samples)


def invert_cdf(bins, cdf, u):
    """intern/ray.py:41-56 via searchsorted (App. B2).  Returns (samples, i0): i0 = index of the last
    cdf knot <= u, the "bin index" that must be bit-exact."""
    n = cdf.shape[-1] - 1
    i0 = torch.searchsorted(cdf.contiguous(), u.contiguous(), right=True) - 1
    i0 = torch.clamp(i0, 0, n)
    i1 = torch.clamp(i0 + 1, max=n)
    c0, c1 = torch.gather(cdf, -1, i0), torch.gather(cdf, -1, i1)
    b0, b1 = torch.gather(bins, -1, i0), torch.gather(bins, -1, i1)
    # reference: x1 = min over knots with cdf > u; if none (cannot happen since u < 1 = cdf[-1]) it is the last knot
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    return b0 + t * (b1 - b0), i0


def sorted_piecewise_constant_pdf(bins, weights, num_samples, randomized=True, jitter=None, return_aux=False):
    """intern/ray.py:12-57."""
    cdf = pdf_to_cdf(weights)
    u = pdf_uniforms(weights.shape[0], num_samples, randomized, weights.dtype, jitter, device=weights.device)
    samples, i0 = invert_cdf(bins, cdf, u)
    if return_aux:
        return samples, cdf, u, i0
    return samples


def resample_t_vals(t_vals, weights, randomized, resample_padding, jitter=None):
    """intern/ray.py:136-149 (the no_grad block)."""
    w = blur_weights(weights.detach(), resample_padding)
    return sorted_piecewise_constant_pdf(t_vals.detach(), w, t_vals.shape[-1], randomized, jitter)


def resample_along_rays(origins, directions, radii, t_vals, weights, randomized, resample_padding, jitter=None):
    """intern/ray.py:118-153."""
    new_t = resample_t_vals(t_vals, weights, randomized, resample_padding, jitter)
    return new_t, para_rays(new_t, origins, directions, radii)


def density_to_weight(t_vals, density, dirs):
    """model.py:59-78 (density[..., 0] squeezed by the caller here: density is [B, N])."""
    delta = (t_vals[..., 1:] - t_vals[..., :-1]) * torch.linalg.norm(dirs[..., None, :], dim=-1)
    dd = density * delta
    alpha = 1 - torch.exp(-dd)
    trans = torch.exp(-torch.cat([torch.zeros_like(dd[..., :1]), torch.cumsum(dd[..., :-1], dim=-1)], dim=-1))
    return alpha * trans


def volumetric_rendering(rgb, density, t_vals, dirs, white_bkgd):
    """intern/ray.py:155-191.  rgb [B,N,3], density [B,N,1]."""
    t_mids = 0.5 * (t_vals[..., :-1] + t_vals[..., 1:])
    weights = density_to_weight(t_vals, density[..., 0], dirs)
    comp_rgb = (weights[..., None] * rgb).sum(dim=-2)
    acc = weights.sum(dim=-1)
    distance = (weights * t_mids).sum(dim=-1) / acc
    distance = torch.clamp(torch.nan_to_num(distance), t_vals[:, 0], t_vals[:, -1])
    if white_bkgd:
        comp_rgb = comp_rgb + (1.0 - acc[..., None])
    return comp_rgb, distance, acc, weights


# ----------------------------------------------------------------------------------------------
# intern/encoding.py
# ----------------------------------------------------------------------------------------------
def integrated_pos_enc(mean, cov):
    """intern/encoding.py:33-56 — single-scale IPE on 21 directions (App. A3, B3)."""
    P = basis(mean.dtype, mean.device)
    gamma = torch.matmul(mean, P.T)
    sigma = torch.einsum("kc,...cd,kd->...k", P, cov, P)
    damp = torch.exp(-0.5 * sigma)
    return torch.cat((damp * torch.sin(gamma), damp * torch.cos(gamma)), -1)


def pos_enc(mean):
    """intern/encoding.py:57-60 — the branch without a covariance: plain sin / cos of the 21 projections."""
    gamma = torch.matmul(mean, basis(mean.dtype, mean.device).T)
    return torch.cat((torch.sin(gamma), torch.cos(gamma)), -1)


def viewdir_enc(viewdirs, min_deg=0, max_deg=4):
    """intern/encoding.py:69-90 — arccos(z), arctan(y/(x+1e-6)) (App. A10)."""
    scales = torch.tensor([2.0**i for i in range(min_deg, max_deg)], dtype=viewdirs.dtype, device=viewdirs.device)
    x, y, z = viewdirs[..., 0:1], viewdirs[..., 1:2], viewdirs[..., 2:3]
    theta = scales * torch.arccos(z)
    phi = scales * torch.arctan(y / (x + 1e-6))
    return torch.cat((torch.sin(theta), torch.cos(theta), torch.sin(phi), torch.cos(phi)), -1)


def mlp_input(mean, cov, viewdirs, min_deg=0, max_deg=4):
    """model.py:85-88 / 173-176: [B,N,42] IPE ++ [B,4*(max-min)] view-dir encoding repeated along N (App. A12)."""
    enc = integrated_pos_enc(mean, cov)
    vd = viewdir_enc(viewdirs, min_deg, max_deg)[:, None, :].expand(-1, enc.shape[1], -1)
    return torch.cat((enc, vd), -1)


# ----------------------------------------------------------------------------------------------
# intern/distillation.py, intern/regularization.py, intern/loss.py
# ----------------------------------------------------------------------------------------------
def bounds_per_ray(t_fine, w_fine, t_coarse):
    """Per-ray b[r,i] = sum_j w_fine[r,j] * [t0_j <= R_i and t1_j >= L_i] (closed intervals),
    the per-ray term of intern/distillation.py:25-29 (App. B5)."""
    t0, t1 = t_fine[..., :-1], t_fine[..., 1:]
    L, R = t_coarse[..., :-1], t_coarse[..., 1:]
    overlap = ~((t0[:, None, :] > R[:, :, None]) | (t1[:, None, :] < L[:, :, None]))  # [B, i, j]
    return (overlap * w_fine[:, None, :]).sum(-1)


def bounds(t_fine, w_fine, t_coarse):
    """intern/distillation.py:4-33 — the boolean [B,N] mask indexes ALL rays, so the bound for coarse
    interval i is the batch total, broadcast to every ray (App. A6).  Detached."""
    b = bounds_per_ray(t_fine, w_fine, t_coarse).sum(0, keepdim=True)
    return b.expand_as(w_fine).detach()


def loss_prop(coarse_weights, bnd):
    """intern/distillation.py:35-51."""
    r = F.relu(bnd - coarse_weights)
    return torch.sum(r * r / (coarse_weights + 1e-6)) / bnd.shape[0]


def Loss_prop(t, w, t_hat, w_hat):
    """intern/loss.py:6-21."""
    return loss_prop(w_hat, bounds(t, w, t_hat))


def loss_dist_per_ray(s_vals, weights):
    """Per-ray value of intern/regularization.py:14-17 in O(N) form (App. A9, B4); the reference's
    scalar is the sum over rays."""
    m = 0.5 * (s_vals[..., :-1] + s_vals[..., 1:])
    w_excl = torch.cumsum(weights, -1) - weights
    wm_excl = torch.cumsum(weights * m, -1) - weights * m
    # sum_{i,j} w_i w_j |m_i - m_j| with m non-decreasing = 2 * sum_i w_i (m_i W_<i - (wm)_<i)
    inter = 2 * (weights * (m * w_excl - wm_excl)).sum(-1)
    intra = (weights**2 * (s_vals[..., 1:] - s_vals[..., :-1])).sum(-1) / 3
    return inter + intra


def loss_dist_quadratic(s_vals, weights):
    """intern/regularization.py:3-19 as the literal double sum (vectorised), valid for unsorted s too."""
    m = 0.5 * (s_vals[..., :-1] + s_vals[..., 1:])
    inter = (weights[..., :, None] * weights[..., None, :] * (m[..., :, None] - m[..., None, :]).abs()).sum((-1, -2))
    intra = (weights**2 * (s_vals[..., 1:] - s_vals[..., :-1])).sum(-1) / 3
    return (inter + intra).sum()


def loss_dist(s_vals, weights):
    return loss_dist_per_ray(s_vals, weights).sum()


def mse_to_psnr(mse):
    """intern/loss.py:56-58."""
    return -10.0 * torch.log10(mse)


def Loss_nerf(inp, target):
    """intern/loss.py:23-40: mse summed over channels / batch; loss = -psnr + 30."""
    mse = ((inp[..., :3] - target[..., :3]) ** 2).sum() / inp.shape[0]
    psnr = mse_to_psnr(mse)
    return -psnr + 30, psnr


# ----------------------------------------------------------------------------------------------
# model.py — functional restatement over a reference state_dict
# ----------------------------------------------------------------------------------------------
def init_state_dict(hidden_proposal=256, hidden_nerf=1024, input_size=58, seed=0, dtype=torch.float32):
    """Random-init weights drawn the way the reference draws them: nn.Linear modules created in the
    order of model.py:43-53 / 131-158, then kaiming_uniform_ on every weight (model.py:8-12), prop_net
    first, under torch.manual_seed(seed).  Biases keep nn.Linear's default initialiser."""
    torch.manual_seed(seed)
    sd = {}

    def build(names_dims):
        mods = [(n, torch.nn.Linear(a, b)) for n, a, b in names_dims]
        for _, m in mods:
            torch.nn.init.kaiming_uniform_(m.weight)
        for n, m in mods:
            sd[n + ".weight"], sd[n + ".bias"] = m.weight.detach().to(dtype), m.bias.detach().to(dtype)

    hp, hn = hidden_proposal, hidden_nerf
    build([(f"prop_net.model.{2 * i}", a, b)
           for i, (a, b) in enumerate([(input_size, hp), (hp, hp), (hp, hp), (hp, hp), (hp, 1)])])
    build([("nerf_net.model.0", input_size, hn)] + [(f"nerf_net.model.{2 * i}", hn, hn) for i in range(1, 8)]
          + [("nerf_net.final_density.0", hn, 1), ("nerf_net.final_color.0", hn, 3)])
    return sd


def prop_mlp(sd, x):
    """model.py:43-53: 58->256 ReLU x3, Sigmoid on the 4th, Linear(256,1)."""
    for i in range(4):
        x = F.linear(x, sd[f"prop_net.model.{2 * i}.weight"], sd[f"prop_net.model.{2 * i}.bias"])
        x = torch.relu(x) if i < 3 else torch.sigmoid(x)
    return F.linear(x, sd["prop_net.model.8.weight"], sd["prop_net.model.8.bias"])


def nerf_mlp(sd, x):
    """model.py:131-158: 8x1024 trunk (ReLU x7, Sigmoid), Sigmoid(Linear(1024,1)), Sigmoid(Linear(1024,3))."""
    for i in range(8):
        x = F.linear(x, sd[f"nerf_net.model.{2 * i}.weight"], sd[f"nerf_net.model.{2 * i}.bias"])
        x = torch.relu(x) if i < 7 else torch.sigmoid(x)
    raw_density = torch.sigmoid(F.linear(x, sd["nerf_net.final_density.0.weight"], sd["nerf_net.final_density.0.bias"]))
    raw_rgb = torch.sigmoid(F.linear(x, sd["nerf_net.final_color.0.weight"], sd["nerf_net.final_color.0.bias"]))
    return raw_density, raw_rgb


def prop_forward(sd, rays, num_samples, randomized, density_bias=-1.0, t_rand=None, viewdir_deg=(0, 4)):
    """model.py:80-94."""
    t, (mean, cov) = sample_along_rays(rays.origins, rays.directions, rays.radii, num_samples,
                                       rays.near, rays.far, randomized, t_rand)
    raw = prop_mlp(sd, mlp_input(mean, cov, rays.viewdirs, *viewdir_deg))
    density = F.softplus(raw + density_bias)
    return t, density_to_weight(t, density[..., 0], rays.directions)


def nerf_forward(sd, rays, t_vals, coarse_weights, randomized, density_bias=-1.0, rgb_padding=0.001,
                 resample_padding=0.01, white_bkgd=False, jitter=None, viewdir_deg=(0, 4)):
    """model.py:163-200.  Returns (rgb, dist, acc, t_vals(+1e-6 as the reference returns it), weights, s_vals)."""
    t, (mean, cov) = resample_along_rays(rays.origins, rays.directions, rays.radii, t_vals, coarse_weights,
                                         randomized, resample_padding, jitter)
    raw_density, raw_rgb = nerf_mlp(sd, mlp_input(mean, cov, rays.viewdirs, *viewdir_deg))
    rgb = raw_rgb * (1 + 2 * rgb_padding) - rgb_padding
    density = F.softplus(raw_density + density_bias)
    comp_rgb, distance, acc, weights = volumetric_rendering(rgb, density, t, rays.directions, white_bkgd)
    s_vals, t_shift = t_to_s(t, rays.near, rays.far)
    return comp_rgb, distance, acc, t_shift, weights, s_vals


def model_forward(sd, rays, num_samples, randomized, t_rand=None, jitter=None, **kw):
    """model.py:247-252."""
    t_hat, w_hat = prop_forward(sd, rays, num_samples, randomized, kw.get("density_bias", -1.0), t_rand,
                                viewdir_deg=kw.get("viewdir_deg", (0, 4)))
    return nerf_forward(sd, rays, t_hat, w_hat, randomized, jitter=jitter, **kw)[:3]


def to8b(img):
    """intern/utils.py:17-21 (NumPy): (255 * clip(nan_to_num(x), 0, 1)).astype(uint8), any rank."""
    import numpy as np
    return (255 * np.clip(np.nan_to_num(np.asarray(img)), 0, 1)).astype(np.uint8)


def render_image(sd, rays, height, width, num_samples, chunks=4096, randomized=False, **kw):
    """model.py:254-274: the chunk loop (each chunk is its own batch for the batch-global contraction norm,
    App. A1), to8b on the colours.  Returns (uint8 [h,w,3], float rgb [h,w,3], dist [h,w], acc [h,w])."""
    n = rays[0].shape[0]
    outs = []
    with torch.no_grad():
        for i in range(0, n, chunks):
            outs.append(model_forward(sd, Rays(*[r[i:i + chunks] for r in rays]), num_samples, randomized, **kw))
    rgb = torch.cat([o[0] for o in outs]).reshape(height, width, 3)
    dist = torch.cat([o[1] for o in outs]).reshape(height, width)
    acc = torch.cat([o[2] for o in outs]).reshape(height, width)
    return to8b(rgb.numpy()), rgb, dist, acc


def lr_at(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1.0):
    """intern/scheduler.py:13-23 at scheduler step `step`."""
    if lr_delay_steps > 0:
        delay = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
    else:
        delay = 1.0
    t = min(max(step / max_steps, 0), 1)
    return delay * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


def train_loop(sd, rays, pixels, num_samples, iterations, randomized=False, dist_weight=0.01, weight_decay=1e-5,
               lr_init=2e-3, lr_final=2e-5, max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.1):
    """train.py:38-82: AdamW over ALL parameters (those without a gradient are skipped), lr_decay stepped after every
    optimiser step, 2 proposal sub-steps + 1 NeRF sub-step per iteration.  Returns (final params, log rows
    [loss_prop, loss_nerf, loss_dist, loss_all, psnr, lr])."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=lr_init, weight_decay=weight_decay)
    cfg = dict(lr_init=lr_init, lr_final=lr_final, max_steps=max_steps, lr_delay_steps=lr_delay_steps,
               lr_delay_mult=lr_delay_mult)
    sched_step = 0

    def set_lr():
        for gch in opt.param_groups:
            gch["lr"] = lr_at(sched_step, **cfg)

    set_lr()  # _LRScheduler.__init__ performs an initial step: lr = get_lr() at last_epoch 0
    log = []
    for _ in range(iterations):
        for _ in range(2):
            t_hat, w_hat = prop_forward(params, rays, num_samples, randomized)
            out = nerf_forward(params, rays, t_hat, w_hat, randomized)
            lp = Loss_prop(out[3].detach(), out[4].detach(), t_hat, w_hat)
            opt.zero_grad()
            lp.backward()
            opt.step()
            sched_step += 1
            set_lr()
        t_hat, w_hat = prop_forward(params, rays, num_samples, randomized)
        rgb, _, _, _, w, s_vals = nerf_forward(params, rays, t_hat.detach(), w_hat.detach(), randomized)
        ln, psnr = Loss_nerf(rgb, pixels)
        ld = loss_dist(s_vals, w)
        la = ln + dist_weight * ld
        opt.zero_grad()
        la.backward()
        opt.step()
        sched_step += 1
        set_lr()
        log.append([float(lp), float(ln), float(ld), float(la), float(psnr), lr_at(sched_step, **cfg)])
    return {k: v.detach() for k, v in params.items()}, log


def train_iteration_grads(sd, rays, pixels, num_samples, randomized=True, dist_weight=0.01):
    """One nerf sub-step and one prop sub-step of train.py:53-82 with autograd on CPU (no optimiser):
    returns (loss_prop, loss_all, grads dict).  Used for gradient parity and as the CPU baseline body."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    t_hat, w_hat = prop_forward(params, rays, num_samples, randomized)
    out = nerf_forward(params, rays, t_hat, w_hat, randomized)
    lp = Loss_prop(out[3].detach(), out[4].detach(), t_hat, w_hat)
    gp = torch.autograd.grad(lp, [params[k] for k in params if k.startswith("prop_net")])
    t_hat, w_hat = prop_forward(params, rays, num_samples, randomized)
    rgb, _, _, _, w, s = nerf_forward(params, rays, t_hat.detach(), w_hat.detach(), randomized)
    ln, _ = Loss_nerf(rgb, pixels)
    la = ln + dist_weight * loss_dist(s, w)
    gn = torch.autograd.grad(la, [params[k] for k in params if k.startswith("nerf_net")])
    grads = dict(zip([k for k in params if k.startswith("prop_net")], gp))
    grads.update(zip([k for k in params if k.startswith("nerf_net")], gn))
    return lp.detach(), la.detach(), grads
