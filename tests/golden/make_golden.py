"""Generate tests/golden/reference_golden.npz by running the LITERAL reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every array is an input to, or an output of, an unmodified reference function.  Inputs are
cloned before each call because the reference mutates them.  Random draws
made inside the reference are captured by re-seeding and repeating the same first draw.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import model as ref_model  # noqa: E402
from intern import distillation as ref_dist  # noqa: E402
from intern import encoding as ref_enc  # noqa: E402
from intern import loss as ref_loss  # noqa: E402
from intern import parameterization as ref_par  # noqa: E402
from intern import ray as ref_ray  # noqa: E402
from intern import regularization as ref_reg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")
G = {}


def put(case, **arrays):
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        G[f"{case}/{k}"] = np.asarray(v)


def make_rays(B, seed, near=0.1, far=10.0, far_first=None):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(B, 3, generator=g)
    d = torch.randn(B, 3, generator=g)
    v = d / d.norm(dim=-1, keepdim=True)
    r = torch.full((B, 1), 1e-3) * (1 + torch.rand(B, 1, generator=g))
    n = torch.full((B, 1), near)
    f = torch.full((B, 1), far)
    if far_first is not None:
        f[0, 0] = far_first
    return ref_ray.Rays(o, d, v, r, n, f)


def clone_rays(rays):
    return ref_ray.Rays(*[x.clone() for x in rays])


def rays_dict(rays):
    return {k: getattr(rays, k) for k in rays._fields}


# ---- scalar helpers: g, t_to_s, s_to_t ------------------------------------------------------
torch.manual_seed(1)
t = torch.rand(5, 9).cumsum(-1) + 0.1
near, far = torch.full((5, 1), 0.1), torch.full((5, 1), 10.0)
put("g", x=t, out=ref_par.g(t.clone()))
put("t_to_s", t=t, near=near, far=far, out=ref_par.t_to_s(t.clone(), near.clone(), far.clone()))
s = torch.linspace(0, 1, 9)[None].repeat(5, 1)
put("s_to_t", s=s, near=near, far=far, out=ref_par.s_to_t(s.clone(), near.clone(), far.clone()))

# ---- sample_along_rays (literal Jacobian loop) ----------------------------------------------
for name, B, N, randomized, kw in [
    ("sample_det", 5, 8, False, {}),
    ("sample_rand", 6, 8, True, {}),
    ("sample_rand_n64", 3, 64, True, {}),
    # one sample holds most of the batch norm => Jacobian != I for it (App. A2)
    ("sample_jac", 3, 4, False, dict(far=1.0, far_first=1e3)),
    # tiny scene: global norm <= 1 => contract() is the identity branch
    ("sample_small", 2, 4, False, dict(near=0.01, far=0.05)),
]:
    rays = make_rays(B, 10 + B, **kw)
    if name == "sample_small":
        rays = rays._replace(directions=rays.directions * 0.05)
    torch.manual_seed(7)
    t_rand = torch.rand(B, N + 1)
    torch.manual_seed(7)
    r = clone_rays(rays)
    t_vals, (means, covs) = ref_ray.sample_along_rays(r.origins, r.directions, r.radii, N, r.near, r.far, randomized)
    put(name, N=N, randomized=randomized, t_rand=t_rand, t_vals=t_vals, means=means, covs=covs, **rays_dict(rays))

# ---- sorted_piecewise_constant_pdf ----------------------------------------------------------
for name, B, N, randomized, tiny in [("pdf_det", 7, 16, False, False), ("pdf_rand", 7, 16, True, False),
                                     ("pdf_rand_n64", 4, 64, True, False), ("pdf_tiny", 3, 8, True, True)]:
    gen = torch.Generator().manual_seed(100 + N)
    bins = (torch.rand(B, N + 1, generator=gen) * 0.5).cumsum(-1) + 0.1
    w = torch.rand(B, N, generator=gen) ** 3
    if tiny:
        w = w * 1e-8  # exercises the eps padding of ray.py:15-19
        w[1] = 0.0
    M = N + 1
    torch.manual_seed(3)
    jitter = torch.empty(B, M).uniform_(to=(1 / M - torch.finfo(torch.float32).eps))
    torch.manual_seed(3)
    out = ref_ray.sorted_piecewise_constant_pdf(bins.clone(), w.clone(), M, randomized)
    put(name, bins=bins, weights=w, M=M, randomized=randomized, jitter=jitter, samples=out)

# ---- resample_along_rays ---------------------------------------------------------------------
for name, B, N, randomized in [("resample_det", 4, 8, False), ("resample_rand", 4, 8, True)]:
    rays = make_rays(B, 21)
    gen = torch.Generator().manual_seed(5)
    r = clone_rays(rays)
    t0, _ = ref_ray.sample_along_rays(r.origins, r.directions, r.radii, N, r.near, r.far, False)
    w = torch.rand(B, N, generator=gen) ** 2 * 0.2
    M = N + 1
    torch.manual_seed(11)
    jitter = torch.empty(B, M).uniform_(to=(1 / M - torch.finfo(torch.float32).eps))
    torch.manual_seed(11)
    r = clone_rays(rays)
    new_t, (means, covs) = ref_ray.resample_along_rays(r.origins, r.directions, r.radii, t0.clone(), w.clone(),
                                                      randomized, 0.01)
    put(name, t_in=t0, weights=w, randomized=randomized, jitter=jitter, t_vals=new_t, means=means, covs=covs,
        **rays_dict(rays))

# ---- encodings -------------------------------------------------------------------------------
gen = torch.Generator().manual_seed(9)
mean = torch.randn(4, 6, 3, generator=gen)
A = torch.randn(4, 6, 3, 3, generator=gen) * 0.3
cov = A @ A.transpose(-1, -2)
put("ipe", mean=mean, cov=cov, enc=ref_enc.PositionalEncoding()(mean.clone(), cov.clone()))
vd = torch.randn(9, 3, generator=gen)
vd = vd / vd.norm(dim=-1, keepdim=True)
put("viewdir", viewdirs=vd, enc=ref_enc.ViewdirectionEncoding(0, 4)(vd.clone()))

# ---- compositing -----------------------------------------------------------------------------
B, N = 6, 12
rays = make_rays(B, 31)
tv = (torch.rand(B, N + 1, generator=gen) * 0.4).cumsum(-1) + 0.1
tv[2, 5:] = tv[2, 5]  # collapsed tail (App. A5): zero-width intervals
rgb = torch.rand(B, N, 3, generator=gen)
dens = torch.rand(B, N, 1, generator=gen) * 3
dens[4] = 0.0  # empty ray: acc = 0 -> nan_to_num path of ray.py:187
for wb in (False, True):
    c, dist, acc, w = ref_ray.volumetric_rendering(rgb.clone(), dens.clone(), tv.clone(), rays.directions.clone(), wb)
    put(f"render_wb{int(wb)}", rgb=rgb, density=dens, t_vals=tv, dirs=rays.directions, comp_rgb=c, distance=dist,
        acc=acc, weights=w)
pn = ref_model.prop_net(randomized=False, num_samples=N, hidden_proposal=8, device=torch.device("cpu"))
put("density_to_weight", t_vals=tv, density=dens, dirs=rays.directions,
    weights=pn.density_to_weight(tv.clone(), dens.clone(), rays.directions.clone()))

# ---- losses ----------------------------------------------------------------------------------
B, N = 5, 10
tf = (torch.rand(B, N + 1, generator=gen) * 0.3).cumsum(-1) + 0.1
tc = (torch.rand(B, N + 1, generator=gen) * 0.3).cumsum(-1) + 0.1
tf[1, 6:] = tf[1, 6]            # collapsed tail
tc[2, 3] = tf[2, 4]             # exact ties between coarse and fine knots
tc[2] = tc[2].sort().values
tc[3] = tf[3]                   # identical grids
wf = torch.rand(B, N, generator=gen) * 0.1
wc = torch.rand(B, N, generator=gen) * 0.1
bnd = ref_dist.bounds(tf.clone(), wf.clone(), tc.clone())
put("interlevel", t_fine=tf, w_fine=wf, t_coarse=tc, w_coarse=wc, bounds=bnd,
    loss_prop=ref_dist.loss_prop(wc.clone(), bnd.clone()),
    Loss_prop=ref_loss.Loss_prop(tf.clone(), wf.clone(), tc.clone(), wc.clone()))
sv = torch.rand(B, N + 1, generator=gen).cumsum(-1)
sv = sv / sv[:, -1:]
sv[0, 7:] = sv[0, 7]
put("distortion", s_vals=sv, weights=wf, loss=ref_reg.loss_dist(sv.clone(), wf.clone()))
a, b = torch.rand(8, 3, generator=gen), torch.rand(8, 3, generator=gen)
ln, psnr = ref_loss.Loss_nerf(a.clone(), b.clone())
put("loss_nerf", input=a, target=b, loss=ln, psnr=psnr)

# ---- whole model at tiny widths (fixture size), literal reference ----------------------------
HP, HN, N, B = 16, 32, 8, 4
for randomized in (False, True):
    torch.manual_seed(0)
    m = ref_model.mipNeRF360(randomized=randomized, num_samples=N, hidden_proposal=HP, hidden_nerf=HN,
                             device=torch.device("cpu"))
    tag = f"model_rand{int(randomized)}"
    sd = m.state_dict()
    if not randomized:
        for k, v in sd.items():
            put("state_dict", **{k: v})
        put("state_dict_full_shapes", keys=np.array(list(ref_model.mipNeRF360(
            num_samples=8, device=torch.device("cpu")).state_dict().keys())))
    rays = make_rays(B, 77)
    pixels = torch.rand(B, 3, generator=gen)
    torch.manual_seed(5)
    t_rand = torch.rand(B, N + 1)
    jitter = torch.empty(B, N + 1).uniform_(to=(1 / (N + 1) - torch.finfo(torch.float32).eps))
    torch.manual_seed(5)
    r = clone_rays(rays)
    t_hat, w_hat = m.prop_net.forward(r)
    r = clone_rays(rays)  # pristine near/far for the nerf call (the reference drifts them across calls, App. A4)
    rgb, dist, acc, t_f, w_f, s_f = m.nerf_net.forward(r, t_vals=t_hat, coarse_weights=w_hat)
    lp = ref_loss.Loss_prop(t_f.detach(), w_f.detach(), t_hat, w_hat)
    gp = torch.autograd.grad(lp, list(m.prop_net.parameters()), retain_graph=True)
    ln, psnr = ref_loss.Loss_nerf(rgb, pixels)
    ld = ref_loss.Loss_dist(s_f, w_f)
    la = ln + 0.01 * ld
    # train.py:68-71 detaches t_hat/w_hat before the nerf step; gradients w.r.t. nerf params are the same here
    gn = torch.autograd.grad(la, list(m.nerf_net.parameters()))
    put(tag, N=N, HP=HP, HN=HN, t_rand=t_rand, jitter=jitter, pixels=pixels, t_hat=t_hat, w_hat=w_hat, rgb=rgb,
        dist=dist, acc=acc, t_fine=t_f, w_fine=w_f, s_fine=s_f, loss_prop=lp, loss_nerf=ln, psnr=psnr, loss_dist=ld,
        loss_all=la, **rays_dict(rays))
    for (k, _), gr in zip(m.prop_net.named_parameters(), gp):
        put(tag, **{"grad.prop_net." + k: gr})
    for (k, _), gr in zip(m.nerf_net.named_parameters(), gn):
        put(tag, **{"grad.nerf_net." + k: gr})
    # full forward on one clone: only rgb/dist/acc are returned (model.py:247-252)
    torch.manual_seed(5)
    out = m(clone_rays(rays))
    put(tag, fwd_rgb=out[0], fwd_dist=out[1], fwd_acc=out[2])

np.savez_compressed(OUT, **G)
print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")
