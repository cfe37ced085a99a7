#!/usr/bin/env python
"""BASELINE.json configs[2..4] on one GPU (run under gpurun; writes JSON into gpurun_out/):

    python scripts/microbench.py micro   [--out f.json]    kernel sweep: cast+contract+IPE, resample, distortion,
                                                           compositing, interlevel at N = 32/64/128, 1M..64M rays
    python scripts/microbench.py render  [--out f.json]    full-image inference: 1008x756 LLFF-shaped (NDC) and
                                                           4946x3286 garden-shaped unbounded poses, rays/s
    python scripts/microbench.py visualize [--out f.json]  depth / normal pictures of a frame (pose.py:112-212) at both
                                                           frame sizes: Mpixel/s and fraction of HBM peak

Every number is CUDA-event time on the launching stream after 3 warm-up launches; inputs are far larger than L2.
Achieved GB/s uses the ALGORITHMIC bytes of SURVEY §8d / BASELINE.md §4, not the measured traffic.
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mipnerf360_b200 import ops  # noqa: E402
from mipnerf360_b200.intern.ray import Rays  # noqa: E402

DEV = torch.device("cuda")


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def timeit(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gen_rays(B):
    o = torch.randn(B, 3, device=DEV)
    d = torch.randn(B, 3, device=DEV)
    return Rays(o, d, d / d.norm(dim=-1, keepdim=True), torch.full((B, 1), 1e-3, device=DEV),
                torch.full((B, 1), 0.1, device=DEV), torch.full((B, 1), 10.0, device=DEV))


SWEEP_B = (1 << 20, 1 << 22, 1 << 24, 1 << 26)


def micro(out_path):
    pk = peak_hbm()
    rows = []
    free = torch.cuda.mem_get_info()[0]
    for N in (32, 64, 128):
        for B in SWEEP_B:
            knots = B * (N + 1) * 4
            if 8 * knots > 0.8 * free:
                continue
            torch.manual_seed(0)
            t = (torch.rand(B, N + 1, device=DEV) * 0.3).cumsum_(-1).add_(0.1)
            w = torch.rand(B, N, device=DEV).mul_(2.0 / N)
            jit = ops.draw_jitter(B, N + 1, DEV)

            def rec(name, ms, bytes_per_ray):
                gbs = bytes_per_ray * B / (ms * 1e-3) / 1e9
                rows.append(dict(kernel=name, rays=B, N=N, ms=ms, bytes_per_ray=bytes_per_ray, achieved_gbs=gbs,
                                 frac_of_hbm_peak=gbs / pk, rays_per_s=B / (ms * 1e-3)))
                print(f"{name:28s} N={N:3d} B={B:9d} {ms:9.3f} ms {gbs:8.1f} GB/s ({gbs / pk:.2f} of {pk:.0f})", flush=True)

            # K4 resample (blur + pdf + cdf + inverse cdf), randomized with supplied jitter
            ms = timeit(lambda: ops.resample(t, w, True, 0.01, jitter=jit))
            rec("resample", ms, 4 * (3 * N + 2) + 4 * (N + 1))
            # the model path: jitter drawn inside the kernel (Philox) and the contraction norm of the new knots accumulated
            dirs0 = torch.randn(B, 3, device=DEV)
            nsq = torch.zeros(1, device=DEV, dtype=torch.float64)
            ms = timeit(lambda: ops.resample(t, w, True, 0.01, directions=dirs0, norm_sq=nsq))
            rec("resample_rng_norm", ms, 4 * (3 * N + 2) + 12)
            # K0 level-0 sampling with in-kernel draw + norm
            near, far = torch.full((B, 1), 0.1, device=DEV), torch.full((B, 1), 10.0, device=DEV)
            ms = timeit(lambda: ops.level0_t_vals(near, far, N, True, directions=dirs0, norm_sq=nsq))
            rec("level0_rng_norm", ms, 4 * (N + 1) + 8 + 12)
            ms = timeit(lambda: ops.level0_t_vals(near, far, N, False))
            rec("level0_deterministic", ms, 4 * (N + 1) + 8)
            del dirs0, near, far
            # K5 distortion forward / backward
            s = t / t[:, -1:]
            ms = timeit(lambda: ops.distortion_per_ray(s, w))
            rec("distortion_fwd", ms, 4 * (2 * N + 1) + 4)
            g1 = torch.ones((), device=DEV)
            gw = torch.empty_like(w)
            ms = timeit(lambda: ops.call("mip360_distortion_bwd", s.data_ptr(), w.data_ptr(), B, N, g1.data_ptr(), gw.data_ptr()))
            rec("distortion_bwd", ms, 4 * (2 * N + 1) + 4 * N)
            del gw, s
            # K3 weights-only compositing and K6 per-ray bounds
            dirs = torch.randn(B, 3, device=DEV)
            ms = timeit(lambda: ops.density_to_weight(t, w, dirs))
            rec("density_to_weight_fwd", ms, 4 * N + 4 * (N + 1) + 12 + 4 * N)
            # K3 full compositing on the MLP head outputs (raw [B,N,4]) forward / backward
            if B * N * 16 * 3 < 0.5 * free:
                raw = torch.rand(B, N, 4, device=DEV)
                ms = timeit(lambda: ops.composite_heads(raw, t, dirs, -1.0, 0.001, False))
                rec("composite_fwd_heads", ms, 16 * N + 4 * (N + 1) + 12 + 20 + 4 * N)
                g_rgb, g_w = torch.rand(B, 3, device=DEV), torch.rand(B, N, device=DEV)
                g_raw = torch.empty_like(raw)
                ms = timeit(lambda: ops.call("mip360_composite_bwd", raw.data_ptr(), None, t.data_ptr(), dirs.data_ptr(), B, N,
                                             1, -1.0, 0.001, 0, g_rgb.data_ptr(), None, None, g_w.data_ptr(), None, None,
                                             g_raw.data_ptr(), None))
                rec("composite_bwd_heads", ms, 16 * N + 4 * (N + 1) + 12 + 12 + 4 * N + 16 * N)
                # the model path: heads' bias + Sigmoid applied here (fused-head MLP) and t_to_s in the same launch
                hb = torch.zeros(64, device=DEV)
                nr, fr = torch.full((B, 1), 0.1, device=DEV), torch.full((B, 1), 10.0, device=DEV)
                ms = timeit(lambda: ops.composite_heads(raw, t, dirs, -1.0, 0.001, False, near=nr, far=fr, head_bias=hb))
                rec("composite_fwd_logits_t_to_s", ms, 16 * N + 4 * (N + 1) + 12 + 8 + 20 + 4 * N + 8 * (N + 1))
                ms = timeit(lambda: ops.call("mip360_composite_bwd", raw.data_ptr(), None, t.data_ptr(), dirs.data_ptr(), B, N,
                                             2, -1.0, 0.001, 0, g_rgb.data_ptr(), None, None, g_w.data_ptr(), None, None,
                                             g_raw.data_ptr(), hb.data_ptr()))
                rec("composite_bwd_logits", ms, 16 * N + 4 * (N + 1) + 12 + 12 + 4 * N + 16 * N)
                del hb, nr, fr
                del raw, g_rgb, g_w, g_raw
            t2 = (torch.rand(B, N + 1, device=DEV) * 0.3).cumsum_(-1).add_(0.1)
            ms = timeit(lambda: ops.bounds_per_ray(t, w, t2))
            rec("bounds_per_ray", ms, 4 * (3 * N + 2) + 4 * N)
            ms = timeit(lambda: ops.bounds_batch_total(t, w, t2))
            rec("bounds_batch_total", ms, 4 * (3 * N + 2))
            tot = torch.rand(N, device=DEV, dtype=torch.float64)
            ms = timeit(lambda: ops.interlevel_loss(w, bound_total=tot))
            rec("interlevel_fwd", ms, 4 * N)
            gw2 = torch.empty_like(w)
            ms = timeit(lambda: ops.call("mip360_interlevel_bwd", w.data_ptr(), None, tot.data_ptr(), B, N, 0, float(B),
                                         g1.data_ptr(), gw2.data_ptr()))
            rec("interlevel_bwd", ms, 8 * N)
            del gw2, tot
            del t2, jit
            # K1 fused cast -> Gaussian -> contract -> IPE, bf16 [N,64] rows (the model path variant)
            if B * N * 128 < 0.45 * free:
                rays = gen_rays(B)
                vd = ops.viewdir_enc(rays.viewdirs)
                nsq = ops.cast_ipe(t[:1024], rays.origins[:1024], rays.directions[:1024], rays.radii[:1024], vd[:1024],
                                   want_x=True)["norm_sq"]
                ms = timeit(lambda: ops.cast_ipe(t, rays.origins, rays.directions, rays.radii, vd, norm_sq=nsq, want_x=True))
                rec("cast_ipe_bf16rows", ms, 48 + 4 * (N + 1) + 128 * N)
                ms = timeit(lambda: ops.frustum_norm_sq(t.data_ptr(), t.data_ptr() + 4, N + 1, rays.directions, B, N))
                rec("frustum_norm_sq", ms, 4 * (N + 1) + 12)
                del rays, vd
            if B * N * 168 < 0.45 * free:
                rays = gen_rays(B)
                ms = timeit(lambda: ops.cast_ipe(t, rays.origins, rays.directions, rays.radii, norm_sq=nsq, want_enc=True))
                rec("cast_ipe_fp32enc", ms, 48 + 4 * (N + 1) + 168 * N)
                del rays
            del t, w, dirs
            torch.cuda.empty_cache()
    json.dump(dict(hbm_peak_gbs=pk, rows=rows), open(out_path, "w"), indent=1)


# ---- synthetic cameras ------------------------------------------------------------
def pinhole_rays(h, w, focal, c2w, near, far, ndc):
    """dataset.py:113-134 pinhole rays on the device; optional NDC as ray.py:59-79 with radii from NDC neighbours."""
    x, y = torch.meshgrid(torch.arange(w, device=DEV, dtype=torch.float32), torch.arange(h, device=DEV, dtype=torch.float32),
                          indexing="xy")
    cam = torch.stack([(x - w * 0.5 + 0.5) / focal, -(y - h * 0.5 + 0.5) / focal, -torch.ones_like(x)], -1)
    d = (cam[..., None, :] * c2w[:3, :3]).sum(-1)
    o = c2w[:3, 3].expand_as(d)
    if ndc:
        tt = -(1.0 + o[..., 2]) / d[..., 2]
        o = o + tt[..., None] * d
        o0 = -focal / (w / 2) * o[..., 0] / o[..., 2]
        o1 = -focal / (h / 2) * o[..., 1] / o[..., 2]
        o2 = 1 + 2 / o[..., 2]
        d0 = -focal / (w / 2) * (d[..., 0] / d[..., 2] - o[..., 0] / o[..., 2])
        d1 = -focal / (h / 2) * (d[..., 1] / d[..., 2] - o[..., 1] / o[..., 2])
        d2 = -2 / o[..., 2]
        o, d = torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)
    v = d / d.norm(dim=-1, keepdim=True)
    dx = (d[:-1] - d[1:]).norm(dim=-1)
    dx = torch.cat([dx, dx[-2:-1]], 0)
    radii = dx[..., None] * 2 / math.sqrt(12)
    flat = lambda a: a.reshape(-1, a.shape[-1]).contiguous()
    n = h * w
    return Rays(flat(o), flat(d), flat(v), flat(radii), torch.full((n, 1), near, device=DEV), torch.full((n, 1), far, device=DEV))


def look_at(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    eye, target, up = (torch.tensor(v, dtype=torch.float32, device=DEV) for v in (eye, target, up))
    z = eye - target
    z = z / z.norm()
    x = torch.linalg.cross(up, z)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    c2w = torch.eye(4, device=DEV)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, eye
    return c2w


def render(out_path, chunks):
    """Under torchrun every rank renders its slab of the image (ray partition) and the result is all-gathered."""
    import torch.distributed as dist
    global DEV
    from mipnerf360_b200.model import mipNeRF360
    from mipnerf360_b200.render import render_image_distributed
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        DEV = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=DEV)
    torch.manual_seed(0)
    model = mipNeRF360(randomized=False, num_samples=64, device=DEV)
    res = []
    cases = [("llff_1008x756_ndc", 756, 1008, 0.82 * 1008, torch.eye(4, device=DEV), 0.05, 1.0, True),
             ("garden_4946x3286_unbounded", 3286, 4946, 0.8 * 4946, look_at((4.0, 0.0, 1.5)), 0.2, 1e3, False)]
    for name, h, w, focal, c2w, near, far, ndc in cases:
        # rays come from the device generator (mip360_generate_rays): nothing is built or uploaded by the host
        rays = ops.generate_rays(c2w.to(DEV), h, w, focal, near, far, ndc=ndc)
        n = h * w
        render_image_distributed(model, Rays(*[r[: 4 * chunks] for r in rays]), 1, min(n, 4 * chunks), chunks)  # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rgb, d, a = render_image_distributed(model, rays, h, w, chunks)
        e1.record()
        torch.cuda.synchronize()
        ms_t = torch.tensor([e0.elapsed_time(e1)], device=DEV, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        ms = float(ms_t)
        flop = n * 64 * 2 * (58 * 256 + 3 * 256 * 256 + 256 + 58 * 1024 + 7 * 1024 * 1024 + 4 * 1024)
        row = dict(case=name, n_gpus=world, rays=n, chunks=chunks, ms=ms, rays_per_s=n / (ms * 1e-3), tflops=flop / (ms * 1e-3) / 1e12,
                   finite=bool(torch.isfinite(rgb).all()), mean_acc=float(a.mean()))
        if rank == 0:
            print(row, flush=True)
        res.append(row)
        del rays, rgb, d, a
        torch.cuda.empty_cache()
    if rank == 0:
        json.dump(dict(rows=res), open(out_path, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


def visualize(out_path):
    """§8f rank 4: frame post-processing.  Algorithmic bytes per pixel: depth 4 + acc 4 read, 3 written (uint8 picture);
    the scaling statistics read depth twice (8 B); automatic planes read depth + acc in each of the 35 passes."""
    pk = peak_hbm()
    rows = []
    for h, w in ((756, 1008), (3286, 4946)):
        n = h * w
        yy, xx = torch.meshgrid(torch.arange(h, device=DEV, dtype=torch.float32),
                                torch.arange(w, device=DEV, dtype=torch.float32), indexing="ij")
        depth = 3 + torch.sin(xx / 40) * torch.cos(yy / 55) + 0.02 * torch.randn(h, w, device=DEV)
        acc = torch.rand(h, w, device=DEV)
        lut = torch.rand(256, 3, device=DEV)
        stats = ops.normals_scaling(depth)

        def rec(name, ms, bytes_per_px):
            gbs = bytes_per_px * n / ms / 1e6
            rows.append(dict(kernel=name, h=h, w=w, ms=ms, mpix_s=n / ms / 1e3, bytes_per_pixel=bytes_per_px, gbs=gbs,
                             frac_hbm=gbs / pk))
            print(f"{name:28s} {h}x{w}  {ms:8.4f} ms  {n / ms / 1e3:10.1f} Mpix/s  {gbs:8.1f} GB/s  {gbs / pk:.3f}")

        rec("normals_scaling", timeit(lambda: ops.normals_scaling(depth)), 8)
        rec("visualize_normals_u8", timeit(lambda: ops.visualize_normals(depth, acc, stats=stats, as_uint8=True)), 11)
        rec("visualize_normals_f32", timeit(lambda: ops.visualize_normals(depth, acc, stats=stats)), 20)
        rec("visualize_depth_u8", timeit(lambda: ops.visualize_depth(depth, acc, 2.0, 4.5, lut=lut, as_uint8=True)), 11)
        rec("visualize_depth_sinebow_u8", timeit(lambda: ops.visualize_depth(depth, acc, 2.0, 4.5, modulus=0.25, as_uint8=True)), 11)
        rec("depth_range_auto_frac0", timeit(lambda: ops.depth_range(depth, acc, None, None, 0.0)), 4)
        rec("depth_range_auto_frac0.05", timeit(lambda: ops.depth_range(depth, acc, None, None, 0.05)), 8 * 34 + 4)
        rgb = torch.rand(h, w, 3, device=DEV)
        rec("to8b_rgb", timeit(lambda: ops.to8b(rgb)), 15)
    json.dump(dict(peak_hbm_gbs=pk, rows=rows), open(out_path, "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["micro", "render", "visualize"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--chunks", type=int, default=65536)
    ap.add_argument("--rays", default=None, help="micro: comma-separated ray counts instead of the 1M..64M sweep")
    a = ap.parse_args()
    if a.rays:
        SWEEP_B = tuple(int(x) for x in a.rays.split(","))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = a.out or os.path.join(ROOT, "gpurun_out", f"{a.mode}.json")
    t0 = time.time()
    micro(out) if a.mode == "micro" else render(out, a.chunks) if a.mode == "render" else visualize(out)
    print(f"wrote {out} in {time.time() - t0:.1f} s")
