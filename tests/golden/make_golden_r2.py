"""Generate tests/golden/reference_golden_r2.npz by running the LITERAL reference (second fixture set: render_image,
to8b, the train.py loop shape, the rarely used branches and sizes beyond one warp per ray).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_r2.py

Every array is an input to, or an output of, an unmodified reference function; inputs are cloned before each call
because the reference mutates them.
"""
import contextlib
import io
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
from intern import scheduler as ref_sched  # noqa: E402
from intern import utils as ref_utils  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_r2.npz")
G = {}
CPU = torch.device("cpu")


def put(case, **arrays):
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        G[f"{case}/{k}"] = np.array(v, copy=True)  # a copy: .numpy() aliases live parameters that the train loop updates


def make_rays(B, seed, near=0.1, far=10.0):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(B, 3, generator=g)
    d = torch.randn(B, 3, generator=g)
    v = d / d.norm(dim=-1, keepdim=True)
    r = torch.full((B, 1), 1e-3) * (1 + torch.rand(B, 1, generator=g))
    return ref_ray.Rays(o, d, v, r, torch.full((B, 1), near), torch.full((B, 1), far))


def clone_rays(rays):
    return ref_ray.Rays(*[x.clone() for x in rays])


def rays_dict(rays):
    return {k: getattr(rays, k) for k in rays._fields}


gen = torch.Generator().manual_seed(2024)

# ---- to8b (intern/utils.py:17-21): 2-D branch, recursive >= 3-D branch, NaN / inf / out-of-range values ----------
x2 = torch.rand(5, 7, generator=gen).numpy() * 1.4 - 0.2
x2[0, 0], x2[0, 1], x2[0, 2], x2[1, 0], x2[1, 1] = np.nan, np.inf, -np.inf, 1.0, 0.0
x2[2, :] = np.linspace(0, 1, 7, dtype=np.float32)
x3 = torch.rand(4, 6, 3, generator=gen).numpy() * 1.2 - 0.1
x3[1, 2, 0] = np.nan
k = np.arange(256, dtype=np.float32) / 255.0  # the values the truncating cast is most sensitive to
put("to8b", x2=x2.astype(np.float32), y2=ref_utils.to8b(x2.astype(np.float32)), x3=x3.astype(np.float32),
    y3=ref_utils.to8b(x3.astype(np.float32)), k=k.reshape(16, 16), yk=ref_utils.to8b(k.reshape(16, 16)))

# ---- render_image (model.py:254-274): literal chunk loop, ragged last chunk, deterministic sampling --------------
HP, HN, N = 16, 32, 8
torch.manual_seed(0)
m = ref_model.mipNeRF360(randomized=False, num_samples=N, hidden_proposal=HP, hidden_nerf=HN, device=CPU)
for kname, v in m.state_dict().items():
    put("render_image_sd", **{kname: v})
H, W, CH = 3, 4, 5
rays = make_rays(H * W, 55)
with contextlib.redirect_stdout(io.StringIO()):
    rgb8, dists, accs = m.render_image(clone_rays(rays), H, W, chunks=CH)
# the float image behind the uint8 one, chunk by chunk, for a tolerance-based comparison
outs = []
with torch.no_grad():
    for i in range(0, H * W, CH):
        outs.append(m(ref_ray.Rays(*[x[i:i + CH].clone() for x in rays])))
put("render_image", H=H, W=W, chunks=CH, N=N, HP=HP, HN=HN, rgb8=rgb8, dists=dists, accs=accs,
    rgb_float=torch.cat([o[0] for o in outs]).reshape(H, W, 3), **rays_dict(rays))

# ---- the loop of train.py:38-82: AdamW over model.parameters(), lr_decay, 3 iterations x 3 optimiser steps -------
torch.manual_seed(0)
m = ref_model.mipNeRF360(randomized=False, num_samples=N, hidden_proposal=HP, hidden_nerf=HN, device=CPU)
for kname, v in m.state_dict().items():
    put("train_loop_sd0", **{kname: v})
cfg = dict(lr_init=2e-3, lr_final=2e-5, max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.1)
optimizer = torch.optim.AdamW(m.parameters(), lr=cfg["lr_init"], weight_decay=1e-5)
scheduler = ref_sched.lr_decay(optimizer, **cfg)
m.train()
B = 8
rays = make_rays(B, 91)
pixels = torch.rand(B, 3, generator=gen)
log = []
for step in range(3):
    r = clone_rays(rays)  # a fresh batch object per iteration, as next(data) gives (the sub-steps then drift it, App. A4)
    for _ in range(2):
        t_hat, w_hat = m.prop_net.forward(r)
        _, _, _, t, w, _ = m.nerf_net.forward(r, t_vals=t_hat, coarse_weights=w_hat)
        t, w = t.detach(), w.detach()
        loss_prop = ref_loss.Loss_prop(t=t, w=w, t_hat=t_hat, w_hat=w_hat)
        optimizer.zero_grad()
        loss_prop.backward()
        optimizer.step()
        scheduler.step()
    t_hat, w_hat = m.prop_net.forward(r)
    t_hat, w_hat = t_hat.detach(), w_hat.detach()
    final_rgbs, _, _, _, fine_weights, s_vals = m.nerf_net.forward(r, t_vals=t_hat, coarse_weights=w_hat)
    loss_nerf, psnr = ref_loss.Loss_nerf(input=final_rgbs, target=pixels)
    loss_dist = ref_loss.Loss_dist(s_vals=s_vals, weights=fine_weights)
    loss_all = loss_nerf + 0.01 * loss_dist
    optimizer.zero_grad()
    loss_all.backward()
    optimizer.step()
    scheduler.step()
    log.append([float(loss_prop), float(loss_nerf), float(loss_dist), float(loss_all), float(psnr),
                float(scheduler.get_last_lr()[-1])])
put("train_loop", N=N, HP=HP, HN=HN, pixels=pixels, log=np.array(log, dtype=np.float64), **rays_dict(rays))
for kname, v in m.state_dict().items():
    put("train_loop_sd3", **{kname: v})
osd = optimizer.state_dict()
put("train_loop_optim", steps=np.array([float(osd["state"][i]["step"]) for i in sorted(osd["state"])]),
    n_params=len(osd["param_groups"][0]["params"]))
for i in sorted(osd["state"]):
    put("train_loop_optim", **{f"exp_avg.{i}": osd["state"][i]["exp_avg"], f"exp_avg_sq.{i}": osd["state"][i]["exp_avg_sq"]})

# ---- conical_frustum_to_gaussian called directly (separate t0 / t1), both formulas; gaussian_to_xyz(diag=True) ----
B, NN = 4, 6
d = torch.randn(B, 3, generator=gen)
t = (torch.rand(B, NN + 1, generator=gen) * 0.4).cumsum(-1) + 0.3
rad = torch.full((B, 1), 2e-2)
for stable in (True, False):
    mean, cov = ref_par.conical_frustum_to_gaussian(d.clone(), t[:, :-1].clone(), t[:, 1:].clone(), rad.clone(), False, stable)
    put(f"frustum_stable{int(stable)}", d=d, t0=t[:, :-1], t1=t[:, 1:], radii=rad, mean=mean, cov=cov)
t_mean, t_var, r_var = (torch.rand(B, NN, generator=gen) + 0.5, torch.rand(B, NN, generator=gen) * 0.1,
                        torch.rand(B, NN, generator=gen) * 0.01)
mean_d, cov_d = ref_par.gaussian_to_xyz(d.clone(), t_mean.clone(), t_var.clone(), r_var.clone(), diag=True)
mean_f, cov_f = ref_par.gaussian_to_xyz(d.clone(), t_mean.clone(), t_var.clone(), r_var.clone(), diag=False)
put("gaussian_to_xyz", d=d, t_mean=t_mean, t_var=t_var, r_var=r_var, mean_diag=mean_d, cov_diag=cov_d, mean_full=mean_f,
    cov_full=cov_f)
# PositionalEncoding without a covariance (encoding.py:57-60)
pe = ref_enc.PositionalEncoding()
xm = torch.randn(5, 3, generator=gen)
put("pos_enc_plain", mean=xm, enc=pe(xm.clone(), None))

# ---- view-direction encodings at other degrees (encoding.py:63-90) ------------------------------------------------
vd = torch.randn(7, 3, generator=gen)
vd = vd / vd.norm(dim=-1, keepdim=True)
for lo, hi in ((0, 4), (1, 3), (0, 6), (2, 3)):
    put(f"viewdir_{lo}_{hi}", viewdirs=vd, enc=ref_enc.ViewdirectionEncoding(lo, hi)(vd.clone()))

# ---- more than 128 samples per ray (beyond one warp x 4 intervals) -------------------------------------------------
N2 = 150
rays = make_rays(2, 13)
torch.manual_seed(7)
t_rand = torch.rand(2, N2 + 1)
torch.manual_seed(7)
r = clone_rays(rays)
t_vals, (means, covs) = ref_ray.sample_along_rays(r.origins, r.directions, r.radii, N2, r.near, r.far, True)
put("n150_sample", N=N2, t_rand=t_rand, t_vals=t_vals, means=means, covs=covs, **rays_dict(rays))
B = 3
bins = (torch.rand(B, N2 + 1, generator=gen) * 0.5).cumsum(-1) + 0.1
w = torch.rand(B, N2, generator=gen) ** 3
M = N2 + 1
torch.manual_seed(3)
jitter = torch.empty(B, M).uniform_(to=(1 / M - torch.finfo(torch.float32).eps))
torch.manual_seed(3)
put("n150_pdf", bins=bins, weights=w, M=M, jitter=jitter,
    samples=ref_ray.sorted_piecewise_constant_pdf(bins.clone(), w.clone(), M, True))
torch.manual_seed(3)
r = clone_rays(make_rays(B, 14))
new_t, (means, covs) = ref_ray.resample_along_rays(r.origins, r.directions, r.radii, bins.clone(), (w * 0.05).clone(), True, 0.01)
put("n150_resample", t_in=bins, weights=w * 0.05, jitter=jitter, t_vals=new_t, means=means, covs=covs, **rays_dict(make_rays(B, 14)))
rgb = torch.rand(B, N2, 3, generator=gen)
dens = torch.rand(B, N2, 1, generator=gen) * 3
dirs = torch.randn(B, 3, generator=gen)
c, dist, acc, wts = ref_ray.volumetric_rendering(rgb.clone(), dens.clone(), bins.clone(), dirs.clone(), True)
put("n150_render", rgb=rgb, density=dens, t_vals=bins, dirs=dirs, comp_rgb=c, distance=dist, acc=acc, weights=wts)
tc = (torch.rand(B, N2 + 1, generator=gen) * 0.5).cumsum(-1) + 0.1
wc = torch.rand(B, N2, generator=gen) * 0.02
wf = torch.rand(B, N2, generator=gen) * 0.02
put("n150_interlevel", t_fine=bins, w_fine=wf, t_coarse=tc, w_coarse=wc, bounds=ref_dist.bounds(bins.clone(), wf.clone(), tc.clone()),
    Loss_prop=ref_loss.Loss_prop(bins.clone(), wf.clone(), tc.clone(), wc.clone()))
sv = torch.rand(B, N2 + 1, generator=gen).cumsum(-1)
sv = sv / sv[:, -1:]
put("n150_distortion", s_vals=sv, weights=wf, loss=ref_reg.loss_dist(sv.clone(), wf.clone()))

# ---- whole model with other view-direction degrees (input width 42 + 4 * (max - min) = 50) ------------------------
torch.manual_seed(0)
m = ref_model.mipNeRF360(randomized=False, num_samples=N, hidden_proposal=HP, hidden_nerf=HN, viewdir_min_deg=1,
                         viewdir_max_deg=3, device=CPU)
for kname, v in m.state_dict().items():
    put("model_vd13_sd", **{kname: v})
rays = make_rays(4, 78)
out = m(clone_rays(rays))
put("model_vd13", N=N, HP=HP, HN=HN, fwd_rgb=out[0], fwd_dist=out[1], fwd_acc=out[2], **rays_dict(rays))

np.savez_compressed(OUT, **G)
print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")
