"""The caller of the hot path: one reference training iteration (train.py:51-82) and the optimiser around it.

train.py itself (data loading, TensorBoard, checkpoints) is out of scope; what matters to the
hot path is its call pattern, reproduced by `Trainer.step`:

    2 x [ prop fwd, nerf fwd (detached), Loss_prop, backward -> prop_net, AdamW step, scheduler step ]
    1 x [ prop fwd (detached), nerf fwd, Loss_nerf + dist_weight_decay * Loss_dist, backward -> nerf_net,
          AdamW step, scheduler step ]

Parameters of each net live in one flat fp32 buffer (the nn.Parameters are views), gradients in another:
AdamW is one fused kernel launch per net and the data-parallel gradient exchange is one
NCCL all-reduce per net over NVLink.  With world_size > 1 the reference's batch-coupled scalars (global
contraction norm, App. A1; batch-total proposal bounds, App. A6) are all-reduced too, so a sharded step
computes what the unsharded reference step computes on the concatenated batch.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from mipnerf360_b200 import ops
from mipnerf360_b200.intern.loss import Loss_dist, Loss_nerf, mse_to_psnr
from mipnerf360_b200.intern.ray import Rays


def lr_at(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1.0):
    """intern/scheduler.py:13-23 evaluated at scheduler step `step` (host scalar math)."""
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0), 1)
    return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


def sharded_loss_nerf(rgb, pixels, world, group=None):
    """loss.py:23-40 on a ray shard: the mse inside the log is the one of the concatenated batch, so that
    SUM-all-reduced gradients equal the unsharded gradients.  Pure torch + one scalar all-reduce."""
    sq = ((rgb[..., :3] - pixels[..., :3]) ** 2).sum()
    tot = sq.detach().clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    mse = (sq + (tot - sq.detach())) / (rgb.shape[0] * world)
    psnr = mse_to_psnr(mse)
    return -psnr + 30, psnr


class FlatAdamW:
    """torch.optim.AdamW semantics (train.py:38) over per-net flat buffers.  Like torch >= 2.0, parameters
    without a gradient are skipped: `step(names)` updates only the nets that were just back-propagated."""

    def __init__(self, groups, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.groups = {}
        self.on_step = {}
        for name, module in groups.items():
            params = list(module.parameters())
            n = sum(p.numel() for p in params)
            dev = params[0].device
            flat = torch.empty(n, device=dev, dtype=torch.float32)
            grad = torch.zeros(n, device=dev, dtype=torch.float32)
            off = 0
            for p in params:
                k = p.numel()
                flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = flat[off:off + k].view_as(p)
                p.grad = grad[off:off + k].view_as(p)
                off += k
            self.groups[name] = dict(params=params, flat=flat, grad=grad, m=torch.zeros_like(flat),
                                     v=torch.zeros_like(flat), step=0)

    def zero_grad(self):
        for g in self.groups.values():
            g["grad"].zero_()

    def step(self, names, lr=None):
        for name in names:
            g = self.groups[name]
            g["step"] += 1
            ops.adamw_step(g["flat"], g["grad"], g["m"], g["v"], self.lr if lr is None else lr, self.betas[0],
                           self.betas[1], self.eps, self.wd, g["step"])
            for cb in self.on_step.get(name, ()):
                cb()  # e.g. PackedMLP.invalidate: the kernel wrote through the flat buffer, not through autograd


class Trainer:
    def __init__(self, model, lr_init=2e-3, lr_final=2e-5, max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.1,
                 weight_decay=1e-5, dist_weight_decay=0.01):
        self.model = model
        self.sched = dict(lr_init=lr_init, lr_final=lr_final, max_steps=max_steps, lr_delay_steps=lr_delay_steps,
                          lr_delay_mult=lr_delay_mult)
        self.opt = FlatAdamW({"prop": model.prop_net, "nerf": model.nerf_net}, lr_init, weight_decay)
        self.opt.on_step = {"prop": [model.prop_net._packed.invalidate], "nerf": [model.nerf_net._packed.invalidate]}
        self.dist_weight_decay = dist_weight_decay
        self.sched_step = 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    # -- pieces ------------------------------------------------------------------------------------
    def _lr(self):
        return lr_at(self.sched_step, **self.sched)

    def _allreduce_grads(self, name):
        if self.world > 1:
            dist.all_reduce(self.opt.groups[name]["grad"], op=dist.ReduceOp.SUM)

    def _optim(self, name):
        self._allreduce_grads(name)
        self.opt.step([name], lr=self._lr())
        self.sched_step += 1  # scheduler.step() after every optimizer.step() (train.py:64,82; App. A11)

    def _loss_prop(self, t, w, t_hat, w_hat):
        """loss.py:18-19 with the bound total and the batch size taken over ALL ranks."""
        b = ops.bounds_per_ray(t, w, t_hat)
        total = ops.bounds_total(b)
        batch = float(w_hat.shape[0])
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM)
            batch *= self.world
        return ops.interlevel_loss(w_hat, bound_total=total, batch_div=batch)

    def _loss_nerf(self, rgb, pixels):
        if self.world == 1:
            return Loss_nerf(rgb, pixels)
        return sharded_loss_nerf(rgb, pixels, self.world)

    # -- one reference iteration -------------------------------------------------------------------
    def step(self, rays, pixels):
        """train.py:52-82 on device-resident rays/pixels.  Returns (loss_prop, loss_all, psnr) as device scalars."""
        m = self.model
        loss_prop = None
        for _ in range(2):
            t_hat, w_hat = m.prop_net.forward(rays)
            with torch.no_grad():  # train.py:55-57: the nerf outputs are detached before use
                _, _, _, t, w, _ = m.nerf_net.forward(rays, t_hat, w_hat)
            loss_prop = self._loss_prop(t, w, t_hat, w_hat)
            self.opt.zero_grad()
            loss_prop.backward()
            self._optim("prop")
        with torch.no_grad():
            t_hat, w_hat = m.prop_net.forward(rays)
        rgb, _, _, t, w, s = m.nerf_net.forward(rays, t_hat, w_hat)
        loss_nerf, psnr = self._loss_nerf(rgb, pixels)
        loss_dist = Loss_dist(s, w)
        loss_all = loss_nerf + self.dist_weight_decay * loss_dist
        self.opt.zero_grad()
        loss_all.backward()
        self._optim("nerf")
        return loss_prop.detach(), loss_all.detach(), psnr.detach()

    def step_host(self, rays_host, pixels_host):
        """Same, from pinned host buffers: H2D copy of the batch, the iteration, D2H read of the losses."""
        dev = next(self.model.parameters()).device
        rays = Rays(*[r.to(dev, non_blocking=True) for r in rays_host])
        pixels = pixels_host.to(dev, non_blocking=True)
        lp, la, psnr = self.step(rays, pixels)
        return torch.stack([lp, la, psnr]).cpu()
