"""Host-side multi-GPU logic on CPU with gloo, world_size 2 (no kernels involved): ray partitioning and image
gather, the sharded photometric loss (sharded gradients == unsharded), additivity of the batch-coupled
quantities the ranks exchange (App. A1, A6), the flat parameter views, the LR schedule."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mip360_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mipnerf360_b200.render import gather_slabs, shard_bounds
        from mipnerf360_b200.train import sharded_loss_nerf
        g = torch.Generator().manual_seed(0)
        n = 37  # ragged: 19 + 18
        full = torch.rand(n, 5, generator=g)
        lo, hi = shard_bounds(n, rank, world)
        out = gather_slabs(full[lo:hi].clone(), n, world)
        ok_gather = torch.equal(out, full)
        # sharded Loss_nerf: gradient of the global loss w.r.t. the local rgb
        B = 16
        rgb = torch.rand(B * world, 3, generator=g)
        pix = torch.rand(B * world, 3, generator=g)
        ref_in = rgb.clone().requires_grad_(True)
        ref_loss, ref_psnr = O.Loss_nerf(ref_in, pix)
        ref_loss.backward()
        loc = rgb[rank * B:(rank + 1) * B].clone().requires_grad_(True)
        loss, psnr = sharded_loss_nerf(loc, pix[rank * B:(rank + 1) * B], world)
        loss.backward()
        ok_loss = torch.allclose(loss, ref_loss, rtol=1e-6) and torch.allclose(psnr, ref_psnr, rtol=1e-6)
        ok_grad = torch.allclose(loc.grad, ref_in.grad[rank * B:(rank + 1) * B], rtol=1e-5, atol=1e-8)
        # A6: bound totals are additive over ray shards; A1: the squared norm too
        N = 8
        tf = (torch.rand(B * world, N + 1, generator=g) * 0.3).cumsum(-1) + 0.1
        tc = (torch.rand(B * world, N + 1, generator=g) * 0.3).cumsum(-1) + 0.1
        wf = torch.rand(B * world, N, generator=g) * 0.1
        sl = slice(rank * B, (rank + 1) * B)
        tot = O.bounds_per_ray(tf[sl], wf[sl], tc[sl]).sum(0).double()
        dist.all_reduce(tot)
        ok_bounds = torch.allclose(tot.float(), O.bounds(tf, wf, tc)[0], rtol=1e-5)
        d = torch.randn(B * world, 3, generator=g)
        nsq = torch.tensor([O.frustum_norm_sq(tf[sl], d[sl])], dtype=torch.float64)
        dist.all_reduce(nsq)
        ok_norm = abs(float(nsq) - O.frustum_norm_sq(tf, d)) < 1e-9 * float(nsq)
        q.put((rank, ok_gather, ok_loss, ok_grad, ok_bounds, ok_norm))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_host_logic():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r


def test_shard_bounds_cover_everything():
    from mipnerf360_b200.render import shard_bounds
    for n in (0, 1, 7, 8, 9, 762048, 16252556):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= hi - lo <= (n + world - 1) // world for lo, hi in spans)


def test_lr_schedule_matches_reference_formula():
    """intern/scheduler.py:13-23 restated with numpy semantics."""
    import numpy as np
    from mipnerf360_b200.train import lr_at
    cfg = dict(lr_init=2e-3, lr_final=2e-5, max_steps=200000, lr_delay_steps=2500, lr_delay_mult=0.1)
    for step in (0, 1, 100, 2499, 2500, 2501, 100000, 200000, 250000):
        delay = cfg["lr_delay_mult"] + (1 - cfg["lr_delay_mult"]) * np.sin(0.5 * np.pi * np.clip(step / cfg["lr_delay_steps"], 0, 1))
        t = np.clip(step / cfg["max_steps"], 0, 1)
        ref = delay * np.exp(np.log(cfg["lr_init"]) * (1 - t) + np.log(cfg["lr_final"]) * t)
        assert math.isclose(lr_at(step, **cfg), ref, rel_tol=1e-12)
    assert lr_at(5, 1e-3, 1e-4, 10) == pytest.approx(1e-3 * (0.1 ** 0.5))


def test_constant_sampling_vectors_are_the_reference_formulas_and_cached():
    """ops.pdf_u_base / the level-0 s grid (ray.py:31-38,100): same torch ops as the reference, built once per device."""
    import torch
    from mipnerf360_b200 import ops
    eps = float(torch.finfo(torch.float32).eps)
    for m in (33, 65, 129):
        u = ops.pdf_u_base(m, True, "cpu")
        assert torch.equal(u, torch.arange(m) * (1 / m))                       # ray.py:31-33 before the jitter
        assert ops.pdf_u_base(m, True, torch.device("cpu")) is u               # cached
        d = ops.pdf_u_base(m, False, "cpu")
        assert torch.equal(d, torch.linspace(0.0, 1.0 - eps, m))               # ray.py:38
        assert d is not u and ops.pdf_u_base(m, False, "cpu") is d
    assert ops.jitter_scale(65) == float(torch.tensor(1 / 65 - eps, dtype=torch.float32))


def test_host_result_returns_a_private_copy():
    """train.HostResult (Trainer.step_host(wait=False)): result() hands out a copy, the pinned slot may be reused."""
    import torch
    from mipnerf360_b200.train import HostResult

    class _Ev:
        waited = 0

        def synchronize(self):
            self.waited += 1

    buf, ev = torch.tensor([1.0, 2.0, 3.0]), _Ev()
    h = HostResult(buf, ev)
    out = h.result()
    buf.zero_()
    assert ev.waited == 1 and out.tolist() == [1.0, 2.0, 3.0]
