"""The caller of the hot path: one reference training iteration (train.py:51-82) and the optimiser around it.

train.py itself (data loading, TensorBoard, checkpoints) is out of scope; what matters to the
hot path is its call pattern, reproduced by `Trainer.step`:

    2 x [ prop fwd, nerf fwd (detached), Loss_prop, backward -> prop_net, AdamW step, scheduler step ]
    1 x [ prop fwd (detached), nerf fwd, Loss_nerf + dist_weight_decay * Loss_dist, backward -> nerf_net,
          AdamW step, scheduler step ]

Parameters of each net live in one flat fp32 buffer (the nn.Parameters are views), gradients in another:
AdamW and the bf16 re-cast of the GEMM operands are ONE fused kernel launch per net.

Data parallel: one process per GPU, the ray batch is sharded, weights are replicated.  Gradients are
all-reduced over NCCL per layer bucket, last layer first, on a side stream, while the remaining dgrad / wgrad
GEMMs of the backward pass run.  The reference's batch-coupled scalars (global contraction norm, App. A1;
batch-total proposal bounds, App. A6; the squared error inside Loss_nerf's log) are all-reduced too, so a sharded
step computes what the unsharded reference step computes on the concatenated batch.
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.distributed as dist

from mipnerf360_b200 import _lib, ops
from mipnerf360_b200.intern.loss import Loss_dist, Loss_nerf, mse_to_psnr
from mipnerf360_b200.intern.ray import Rays


@contextlib.contextmanager
def _nvtx(name):
    """NVTX range around a phase of the iteration (visible in Nsight Systems / ncu --nvtx; free without a profiler)."""
    torch.cuda.nvtx.range_push("mip360/" + name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


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
    without a gradient are skipped: `step(names)` updates only the nets that were just back-propagated.

    groups: {name: nn.Module}.  The parameters of each module are re-pointed at views of one flat fp32 buffer
    (values preserved) and receive preallocated .grad views of a second one.  Every `mlp.PackedMLP` found on the
    modules (attribute `_packed`) is switched to direct gradient accumulation and is refreshed by `step` — the
    update kernel writes through raw pointers, which autograd's version counters do not see."""

    def __init__(self, groups, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.groups = {}
        for name, module in groups.items():
            params = list(module.parameters())
            n = sum(p.numel() for p in params)
            dev = params[0].device
            flat = torch.empty(n, device=dev, dtype=torch.float32)
            grad = torch.zeros(n, device=dev, dtype=torch.float32)
            off, offsets = 0, {}
            for p in params:
                k = p.numel()
                flat[off:off + k].copy_(p.data.reshape(-1))
                p.data = flat[off:off + k].view_as(p)
                p.grad = grad[off:off + k].view_as(p)
                offsets[id(p)] = (off, k)
                off += k
            packed = [m._packed for m in module.modules() if hasattr(m, "_packed")]
            for pk in packed:
                pk.direct_grad = True
            # one PackedMLP covering exactly the group's parameters: AdamW and the operand refresh are one launch
            fused = len(packed) == 1 and {id(p) for p in packed[0].params()} == {id(p) for p in params}
            self.groups[name] = dict(params=params, flat=flat, grad=grad, m=torch.zeros_like(flat),
                                     v=torch.zeros_like(flat), step=0, offsets=offsets, packed=packed, fused=fused)

    def span(self, name, params):
        """[lo, hi) of the flat buffer covered by `params` (which must be adjacent in it)."""
        offs = sorted(self.groups[name]["offsets"][id(p)] for p in params)
        for (a, n), (b, _) in zip(offs, offs[1:]):
            assert a + n == b, "parameters of one bucket must be contiguous in the flat buffer"
        return offs[0][0], offs[-1][0] + offs[-1][1]

    def zero_grad(self, names=None):
        for name in (names or self.groups):
            self.groups[name]["grad"].zero_()

    def step(self, names, lr=None, hyper_dev=None, zero_grad=False):
        """hyper_dev: device tensor [lr, 1 - beta1^step, sqrt(1 - beta2^step)] read by the kernel instead of the host
        values (CUDA-graph replays).  zero_grad: clear the gradients in the same pass (fused path only)."""
        for name in names:
            g = self.groups[name]
            g["step"] += 1
            lr_now = self.lr if lr is None else lr
            if g["fused"]:
                ops.adamw_pack_step(g["packed"][0], g["flat"], g["grad"], g["m"], g["v"], lr_now, self.betas[0],
                                    self.betas[1], self.eps, self.wd, g["step"], hyper_dev=hyper_dev, zero_grad=zero_grad)
                continue
            ops.adamw_step(g["flat"], g["grad"], g["m"], g["v"], lr_now, self.betas[0], self.betas[1], self.eps, self.wd,
                           g["step"])
            for pk in g["packed"]:
                pk.invalidate()

    # -- checkpoint round trip (train.py:39-41,98-103: optimizer.state_dict() <-> optim.pt) ----------------------
    def state_dict(self):
        """The layout torch.optim.AdamW(model.parameters()) produces for the same parameters in the same order, so
        optim_{step}.pt files are interchangeable with the reference's."""
        state, idx = {}, 0
        for g in self.groups.values():
            for p in g["params"]:
                off, k = g["offsets"][id(p)]
                if g["step"] > 0:
                    state[idx] = dict(step=torch.tensor(float(g["step"])),
                                      exp_avg=g["m"][off:off + k].view_as(p).clone(),
                                      exp_avg_sq=g["v"][off:off + k].view_as(p).clone())
                idx += 1
        group = dict(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.wd, amsgrad=False, maximize=False,
                     foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=True,
                     params=list(range(idx)))
        return dict(state=state, param_groups=[group])

    def load_state_dict(self, sd):
        pg = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.wd = pg["lr"], tuple(pg["betas"]), pg["eps"], pg["weight_decay"]
        idx = 0
        for g in self.groups.values():
            steps = set()
            for p in g["params"]:
                off, k = g["offsets"][id(p)]
                st = sd["state"].get(idx)
                if st is None:
                    g["m"][off:off + k].zero_()
                    g["v"][off:off + k].zero_()
                    steps.add(0)
                else:
                    g["m"][off:off + k].copy_(st["exp_avg"].reshape(-1))
                    g["v"][off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(st["step"]))
                idx += 1
            if len(steps) != 1:
                raise ValueError("FlatAdamW keeps one step count per net; the checkpoint has mixed counts " + str(steps))
            g["step"] = steps.pop()


class Trainer:
    """One reference training iteration per `step` (train.py:52-82), single GPU or data parallel.

    group: torch.distributed process group of the data-parallel ranks (default: the world when initialised).
    data_parallel=False: ignore torch.distributed (every rank trains on its own, whole, batch).
    overlap: bucket the gradient all-reduce per layer (last layer first) and issue each bucket on a side stream as soon
        as its wgrad has been enqueued, so that it overlaps the remaining dgrad / wgrad GEMMs.  Off by default: measured
        on 8 B200 at 2048 rays per rank (profiles/r02_dp_breakdown_n8.json) the bucketed form costs 0.35 ms per
        iteration against 0.18 ms for one flat 29.6 MB all-reduce after the backward pass — the persistent GEMM kernels
        occupy every SM, so NCCL's copy kernels slow them down by more than the exposed transfer costs over NVSwitch.
    graph: capture the whole iteration (3 forward/backward/optimiser sub-steps, collectives included) in ONE CUDA
        graph after `graph_warmup` eager iterations and replay it afterwards: ~330 kernel launches and the Python
        between them become one launch.  The learning rate and Adam's bias corrections of the three optimiser steps are
        read from a small device tensor the host refreshes before every replay; inputs are copied into static buffers.
        A new batch shape triggers a new capture."""

    def __init__(self, model, lr_init=2e-3, lr_final=2e-5, max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.1,
                 weight_decay=1e-5, dist_weight_decay=0.01, group=None, data_parallel=True, overlap=False,
                 fused_zero_grad=True, graph=False, graph_warmup=2):
        self.model = model
        self.sched = dict(lr_init=lr_init, lr_final=lr_final, max_steps=max_steps, lr_delay_steps=lr_delay_steps,
                          lr_delay_mult=lr_delay_mult)
        self.opt = FlatAdamW({"prop": model.prop_net, "nerf": model.nerf_net}, lr_init, weight_decay)
        self.dist_weight_decay = dist_weight_decay
        self.sched_step = 0
        # the fused AdamW launch clears the gradient buffer it has just consumed, so the next sub-step of that net
        # accumulates into zeros without a separate memset (set False to keep the gradients readable after a step)
        self.fused_zero_grad = bool(fused_zero_grad)
        self._grads_clean = {"prop": False, "nerf": False}
        # timing experiments only (scripts/dp_breakdown.py): leave out a class of collectives — results are then wrong
        self.debug_skip = set()  # subset of {"grads", "scalars"}
        self.use_graph = bool(graph)
        self.graph_warmup = int(graph_warmup)
        self._graphs = {}       # batch size -> dict(graph, rays, pixels, out, launches)
        self.replayed_launches = 0  # kernels of this library launched through graph replays (the C-side counter only sees captures)
        self._host_out, self._host_ev, self._host_i = None, None, 0   # step_host: pinned result slots
        self._eager_calls = 0
        self._hyper = None      # device [3 sub-steps, 3] = lr, 1 - beta1^t, sqrt(1 - beta2^t), read by the captured AdamW
        self._capture_substep = None
        self.group = group
        self.world = dist.get_world_size(group) if data_parallel and dist.is_available() and dist.is_initialized() else 1
        self.overlap = bool(overlap) and self.world > 1
        self._pending = []
        if self.world > 1:
            if group is None:
                self.group = dist.group.WORLD
            if self.overlap:
                self.comm_stream = torch.cuda.Stream()
                for name, net in (("prop", model.prop_net), ("nerf", model.nerf_net)):
                    net._packed.grad_hook = self._make_hook(name, net._packed)

    # -- gradient exchange ---------------------------------------------------------------------------------------
    def _make_hook(self, name, pk):
        """Bucket = the gradients completed by one wgrad: trunk layer l (plus the heads for the last trunk layer,
        whose wgrad runs first).  Issued from the backward pass right after that wgrad was enqueued."""
        L = len(pk.trunk)
        spans = []
        for l in range(L):
            ps = [pk.trunk[l][0].weight, pk.trunk[l][0].bias]
            if l == L - 1:
                for h in pk.heads:
                    ps += [h.weight, h.bias]
            spans.append(self.opt.span(name, ps))
        grad = self.opt.groups[name]["grad"]

        def hook(l):
            if "grads" in self.debug_skip:
                return
            lo, hi = spans[l]
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                dist.all_reduce(grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            self._pending.append(name)
        return hook

    def _finish_grads(self, name):
        if self.world == 1 or "grads" in self.debug_skip:
            return
        if self.overlap:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            self._pending.clear()
        else:
            dist.all_reduce(self.opt.groups[name]["grad"], op=dist.ReduceOp.SUM, group=self.group)

    @contextlib.contextmanager
    def _sharded_batch(self):
        """Inside a training sub-step the nets see one shard of a data-parallel batch and sum the batch-global
        contraction norm over the group (model._encode).  Outside (eval, render_image on any subset of the ranks)
        model() stays free of collectives."""
        nets = (self.model.prop_net, self.model.nerf_net)
        prev = [n.batch_group for n in nets]
        for n in nets:
            n.batch_group = self.group if self.world > 1 and "scalars" not in self.debug_skip else None
        try:
            yield
        finally:
            for n, g in zip(nets, prev):
                n.batch_group = g

    # -- pieces ------------------------------------------------------------------------------------
    def _lr(self):
        return lr_at(self.sched_step, **self.sched)

    def _zero_grad(self, name):
        """optimizer.zero_grad() of train.py:61,79 for the net about to be back-propagated."""
        if not self._grads_clean[name]:
            self.opt.zero_grad([name])
        self._grads_clean[name] = False

    def _optim(self, name):
        self._finish_grads(name)
        zero = self.fused_zero_grad and self.opt.groups[name]["fused"]
        hyper = None
        if self._capture_substep is not None:  # being captured: this step's scalars come from device memory at replay
            hyper = self._hyper[self._capture_substep]
            self._capture_substep += 1
        self.opt.step([name], lr=self._lr(), zero_grad=zero, hyper_dev=hyper)
        self._grads_clean[name] = zero
        self.sched_step += 1  # scheduler.step() after every optimizer.step() (train.py:64,82; App. A11)

    def _loss_prop(self, t, w, t_hat, w_hat):
        """loss.py:18-19 with the bound total and the batch size taken over ALL ranks."""
        total = ops.bounds_batch_total(t, w, t_hat)  # per-ray bounds and their batch totals in one launch
        batch = float(w_hat.shape[0])
        if self.world > 1:
            if "scalars" not in self.debug_skip:
                dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
            batch *= self.world
        return ops.interlevel_loss(w_hat, bound_total=total, batch_div=batch)

    def _loss_nerf(self, rgb, pixels):
        if self.world == 1 or "scalars" in self.debug_skip:
            return Loss_nerf(rgb, pixels)
        return sharded_loss_nerf(rgb, pixels, self.world, self.group)

    # -- one reference iteration -------------------------------------------------------------------
    def prop_substep(self, rays):
        """train.py:54-64: one proposal update.  Returns the loss (device scalar; data parallel: this shard's share of
        the global loss — the sum over ranks is the reference's value)."""
        m = self.model
        with _nvtx("prop_substep"):
            with self._sharded_batch(), _nvtx("forward"):
                t_hat, w_hat = m.prop_net.forward(rays)
                with torch.no_grad():  # train.py:55-57: the nerf outputs are detached before use
                    _, _, _, t, w, _ = m.nerf_net.forward(rays, t_hat, w_hat)
            with _nvtx("loss_prop"):
                loss_prop = self._loss_prop(t, w, t_hat, w_hat)
            self._zero_grad("prop")
            with _nvtx("backward"):
                loss_prop.backward()
            with _nvtx("allreduce+adamw"):
                self._optim("prop")
        return loss_prop.detach()

    def nerf_substep(self, rays, pixels):
        """train.py:68-82: the NeRF update.  Returns (loss_all, psnr) (device scalars).  Data parallel: psnr and the
        photometric part of loss_all are those of the whole batch, the distortion part is the local shard's sum."""
        m = self.model
        with _nvtx("nerf_substep"):
            with self._sharded_batch(), _nvtx("forward"):
                with torch.no_grad():
                    t_hat, w_hat = m.prop_net.forward(rays)
                rgb, _, _, t, w, s = m.nerf_net.forward(rays, t_hat, w_hat)
            with _nvtx("losses"):
                loss_nerf, psnr = self._loss_nerf(rgb, pixels)
                loss_dist = Loss_dist(s, w)
                loss_all = loss_nerf + self.dist_weight_decay * loss_dist
            self._zero_grad("nerf")
            with _nvtx("backward"):
                loss_all.backward()
            with _nvtx("allreduce+adamw"):
                self._optim("nerf")
        return loss_all.detach(), psnr.detach()

    def _step_eager(self, rays, pixels):
        ops.rng_advance(pixels.device)  # first node of a captured iteration: every replay draws fresh numbers
        loss_prop = None
        for _ in range(2):
            loss_prop = self.prop_substep(rays)
        loss_all, psnr = self.nerf_substep(rays, pixels)
        return loss_prop, loss_all, psnr

    # -- whole-iteration CUDA graph -------------------------------------------------------------------------------
    def _hyper_rows(self):
        """[lr, 1 - beta1^t, sqrt(1 - beta2^t)] of the next three optimiser steps (prop, prop, nerf)."""
        b1, b2 = self.opt.betas
        rows, tp, tn = [], self.opt.groups["prop"]["step"], self.opt.groups["nerf"]["step"]
        for i, t in enumerate((tp + 1, tp + 2, tn + 1)):
            rows.append([lr_at(self.sched_step + i, **self.sched), 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t)])
        return torch.tensor(rows, dtype=torch.float32)

    def _capture(self, rays, pixels):
        if not all(g["fused"] for g in self.opt.groups.values()):
            raise RuntimeError("graph=True needs the fused AdamW path (one PackedMLP per net)")
        dev = pixels.device
        # static inputs of the captured iteration: the ray fields, the pixels and the three sub-steps' optimiser scalars
        # are views of ONE flat buffer (each 64-byte aligned), so that a host batch arrives with a single copy
        fields = [r.detach() for r in rays] + [pixels.detach()]
        assert all(f.dtype == torch.float32 for f in fields)
        offs, off = [], 0
        for f in fields:
            offs.append(off)
            off += (f.numel() + 15) // 16 * 16
        flat = torch.zeros(off + 16, device=dev, dtype=torch.float32)
        views = [flat[o:o + f.numel()].view(f.shape) for o, f in zip(offs, fields)]
        for v, f in zip(views, fields):
            v.copy_(f)
        st = dict(rays=Rays(*views[:-1]), pixels=views[-1], hyper=flat[off:off + 9].view(3, 3), flat=flat, offs=offs,
                  host=None, graph=torch.cuda.CUDAGraph())
        self._hyper = st["hyper"]   # read by _optim while capturing: the AdamW launches take lr / bias corrections from here
        # host-side counters advance while capturing (no kernel runs); they are restored and advanced per replay instead
        saved = (self.sched_step, self.opt.groups["prop"]["step"], self.opt.groups["nerf"]["step"], dict(self._grads_clean))
        self._capture_substep = 0
        n0 = _lib.launch_count()
        try:
            with torch.cuda.graph(st["graph"], capture_error_mode="thread_local"):
                lp, la, psnr = self._step_eager(st["rays"], st["pixels"])
                st["out"] = torch.stack([lp, la, psnr])
        finally:
            self._capture_substep = None
        st["launches"] = _lib.launch_count() - n0  # kernels of this library inside one replay
        clean_after = dict(self._grads_clean)
        self.sched_step, self.opt.groups["prop"]["step"], self.opt.groups["nerf"]["step"], self._grads_clean = saved
        if self._grads_clean != clean_after or not all(clean_after.values()):
            # the captured iteration assumes the gradient buffers it finds are the ones it leaves behind (cleared by the
            # fused AdamW); bring the buffers to that state once
            self.opt.zero_grad()
            self._grads_clean = clean_after
        return st

    def _step_graph(self, rays, pixels, raw=False):
        key = int(pixels.shape[0])
        st = self._graphs.get(key)
        if st is None:
            if self._eager_calls < self.graph_warmup:
                self._eager_calls += 1
                return self._step_eager(rays, pixels)
            saved = (self.sched_step, self.opt.groups["prop"]["step"], self.opt.groups["nerf"]["step"], dict(self._grads_clean))
            try:
                st = self._graphs[key] = self._capture(rays, pixels)
            except RuntimeError as e:
                # capture refused (e.g. a tool that serialises streams, an allocator event on uncaptured work): nothing
                # has run, so restore the host-side counters and keep training eagerly
                import warnings
                warnings.warn(f"CUDA-graph capture of the training iteration failed ({e}); continuing without graphs")
                self.sched_step, self.opt.groups["prop"]["step"], self.opt.groups["nerf"]["step"], self._grads_clean = saved
                self._capture_substep = None
                self.use_graph = False
                self.opt.zero_grad()
                self._grads_clean = {"prop": True, "nerf": True}
                torch.cuda.synchronize()
                return self._step_eager(rays, pixels)
        fields = list(rays) + [pixels]
        if all(f.device.type == "cpu" for f in fields):
            # host batch: packed on the CPU into a pinned image of the flat buffer (two of them: the previous step's copy may
            # still be in flight), optimiser scalars included, then ONE host->device copy
            if st["host"] is None:
                st["host"] = [torch.zeros(st["flat"].numel(), dtype=torch.float32).pin_memory() for _ in range(2)]
                st["host_ev"] = [torch.cuda.Event() for _ in range(2)]
                st["host_i"] = 0
            i = st["host_i"] = 1 - st["host_i"]
            st["host_ev"][i].synchronize()      # the copy that last read this image has completed
            hbuf = st["host"][i]
            for o, f in zip(st["offs"], fields):
                hbuf[o:o + f.numel()].copy_(f.reshape(-1))
            n = st["flat"].numel() - 16
            hbuf[n:n + 9].copy_(self._hyper_rows().reshape(-1))
            st["flat"].copy_(hbuf, non_blocking=True)
            st["host_ev"][i].record()
        else:
            for dst, src in zip(st["rays"], rays):
                dst.copy_(src, non_blocking=True)
            st["pixels"].copy_(pixels, non_blocking=True)
            st["hyper"].copy_(self._hyper_rows())  # 36 bytes, staged copy from pageable memory
        st["graph"].replay()
        self.replayed_launches += st["launches"]
        self.sched_step += 3
        self.opt.groups["prop"]["step"] += 2
        self.opt.groups["nerf"]["step"] += 1
        if raw:
            return st["out"]   # the graph's static result tensor: valid until the next replay (stream-ordered reads only)
        out = st["out"].clone()
        return out[0], out[1], out[2]

    def step(self, rays, pixels):
        """train.py:52-82 on device-resident (or pinned host) rays/pixels.  Returns (loss_prop, loss_all, psnr) as
        device scalars."""
        if self.use_graph and _lib.PROFILE is None:  # instrumented runs (per-kernel events) stay eager
            return self._step_graph(rays, pixels)
        return self._step_eager(rays, pixels)

    def step_host(self, rays_host, pixels_host, wait=True):
        """Same, from pinned host buffers: H2D copy of the batch, the iteration, D2H read of the losses
        [loss_prop, loss_all, psnr].  wait=False returns a handle whose .result() blocks for this step's losses: a loop
        that enqueues step i+1 before asking for step i's result keeps the GPU busy across the read-back (the copies and
        the replay of i+1 are queued behind step i on the same stream)."""
        dev = next(self.model.parameters()).device
        if self.use_graph and int(pixels_host.shape[0]) in self._graphs and _lib.PROFILE is None:
            out = self._step_graph(rays_host, pixels_host, raw=True)  # one copy into the graph's static input buffer
        else:
            rays = Rays(*[r.to(dev, non_blocking=True) for r in rays_host])
            pixels = pixels_host.to(dev, non_blocking=True)
            out = torch.stack(self.step(rays, pixels))
        if self._host_out is None:
            self._host_out = [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._host_ev = [torch.cuda.Event() for _ in range(2)]
        i = self._host_i = (self._host_i + 1) % 2   # two slots: the previous step's result may not have been read yet
        self._host_out[i].copy_(out, non_blocking=True)
        self._host_ev[i].record()
        h = HostResult(self._host_out[i], self._host_ev[i])
        return h.result() if wait else h


class HostResult:
    """Losses of one Trainer.step_host(wait=False) call on their way to pinned host memory."""

    def __init__(self, buf, event):
        self._buf, self._event = buf, event

    def result(self):
        self._event.synchronize()
        return self._buf.clone()


def check_sharded_equals_unsharded(device, rays_per_rank=512, num_samples=64, hidden_proposal=256, hidden_nerf=1024,
                                   overlap=True, seed=0):
    """Run under torch.distributed (NCCL, one rank per GPU).  One training iteration on a ray-sharded batch (data
    parallel: sharded rays, all-reduced gradients and batch-coupled scalars) against the same iteration on the
    whole batch computed by this rank alone, same initial weights.  Deterministic sampling, so that shards and
    whole batch take the same samples, and a vanishing learning rate, so that the three sub-steps of both runs see
    the same weights.  Returns a dict of relative errors (losses; per-net gradient, Frobenius)."""
    from mipnerf360_b200.model import mipNeRF360
    from mipnerf360_b200.synthetic import generic_rays
    world, rank = dist.get_world_size(), dist.get_rank()
    B = rays_per_rank * world
    rays, pixels = generic_rays(B, 4242 + seed, device=device)  # the same global batch on every rank
    sl = slice(rank * rays_per_rank, (rank + 1) * rays_per_rank)
    models = []
    for _ in range(2):
        torch.manual_seed(seed)
        models.append(mipNeRF360(randomized=False, num_samples=num_samples, hidden_proposal=hidden_proposal,
                                 hidden_nerf=hidden_nerf, device=device))
    kw = dict(lr_init=1e-12, lr_final=1e-12, fused_zero_grad=False)  # the gradients are compared after the step
    dp = Trainer(models[0], overlap=overlap, **kw)
    solo = Trainer(models[1], data_parallel=False, **kw)
    out = {}
    lp_a = dp.prop_substep(Rays(*[r[sl] for r in rays]))
    lp_b = solo.prop_substep(rays)
    la_a, ps_a = dp.nerf_substep(Rays(*[r[sl] for r in rays]), pixels[sl])
    la_b, ps_b = solo.nerf_substep(rays, pixels)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    # the returned proposal loss is the local shard's share of the global sum (already divided by the global batch)
    lp_sum = lp_a.clone()
    dist.all_reduce(lp_sum)
    out["loss_prop"] = rel(lp_sum, lp_b)
    # loss_all = Loss_nerf (global, through the all-reduced squared error) + 0.01 * Loss_dist of the LOCAL shard:
    # the distortion term is a plain sum over rays (App. A9), so the global value is the sum of the shards' terms
    ln_a = 30 - ps_a
    dist_part = (la_a - ln_a).clone()
    dist.all_reduce(dist_part)
    out["loss_all"] = rel(ln_a + dist_part, la_b)
    out["psnr"] = rel(ps_a, ps_b)
    for name in ("prop", "nerf"):
        out["grad_" + name] = rel(dp.opt.groups[name]["grad"], solo.opt.groups[name]["grad"])
    out["world"] = world
    return out
