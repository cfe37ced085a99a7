#!/usr/bin/env python
"""bench.py — train rays/s of one reference training iteration (train.py:51-82) on synthetic rays.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference ...                     # the reference algorithm on the host cores

Workload (BASELINE.json configs[1]): 16384 rays per GPU per iteration, 64 samples per ray, default config.py
architecture (prop 58-256x4-1, nerf 58-1024x8-{1,3}), random-init weights, randomized sampling, bf16 MLP.
One "step" = 2 proposal sub-steps + 1 NeRF sub-step, each with its AdamW update, exactly the reference's
iteration.  N > 1: one process per GPU (torchrun), every rank holds its own 16384-ray shard (weak scaling),
gradients and the reference's batch-coupled scalars are all-reduced over NCCL.

Prints ONE JSON line (see the contract in the task statement); timing is CUDA events on the launching stream,
max over ranks; inputs are far larger than L2 (2 GB activations per layer), so no explicit L2 flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_SAMPLES = 64
PROP_FLOP_PER_SAMPLE = 2 * (58 * 256 + 3 * 256 * 256 + 256)
NERF_FLOP_PER_SAMPLE = 2 * (58 * 1024 + 7 * 1024 * 1024 + 4 * 1024)
# one reference iteration: 3 fwd of both nets + 2 prop bwd + 1 nerf bwd (bwd = 2 x fwd)   
ITER_FLOP_PER_SAMPLE = 3 * (PROP_FLOP_PER_SAMPLE + NERF_FLOP_PER_SAMPLE) + 4 * PROP_FLOP_PER_SAMPLE + 2 * NERF_FLOP_PER_SAMPLE


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def synth_rays(B, seed, device=None, pin=False):
    """SURVEY §8d 'Generic' synthetic rays: origins, directions ~ N(0,I), radii 1e-3, near 0.1, far 10."""
    from mipnerf360_b200.intern.ray import Rays
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(B, 3, generator=g)
    d = torch.randn(B, 3, generator=g)
    rays = Rays(o, d, d / d.norm(dim=-1, keepdim=True), torch.full((B, 1), 1e-3), torch.full((B, 1), 0.1),
                torch.full((B, 1), 10.0))
    pixels = torch.rand(B, 3, generator=g)
    if pin:
        rays = Rays(*[r.pin_memory() for r in rays])
        pixels = pixels.pin_memory()
    if device is not None:
        rays = Rays(*[r.to(device) for r in rays])
        pixels = pixels.to(device)
    return rays, pixels


# -------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) for the same iteration
# -------------------------------------------------------------------------------------------------
def cpu_iteration(O, params, opt, rays, pixels, N):
    """train.py:53-82 with the oracle's functions (fp32, torch CPU), AdamW included."""
    names_p = [k for k in params if k.startswith("prop_net")]
    names_n = [k for k in params if k.startswith("nerf_net")]
    for _ in range(2):
        t_hat, w_hat = O.prop_forward(params, rays, N, True)
        with torch.no_grad():
            out = O.nerf_forward(params, rays, t_hat, w_hat, True)
        lp = O.Loss_prop(out[3], out[4], t_hat, w_hat)
        opt.zero_grad()
        for k, g in zip(names_p, torch.autograd.grad(lp, [params[k] for k in names_p])):
            params[k].grad = g
        opt.step()
    with torch.no_grad():
        t_hat, w_hat = O.prop_forward(params, rays, N, True)
    rgb, _, _, _, w, s = O.nerf_forward(params, rays, t_hat, w_hat, True)
    ln, _ = O.Loss_nerf(rgb, pixels)
    la = ln + 0.01 * O.loss_dist(s, w)
    opt.zero_grad()
    for k, g in zip(names_n, torch.autograd.grad(la, [params[k] for k in names_n])):
        params[k].grad = g
    opt.step()
    return float(la.detach())


def run_cpu(sample_rays, steps, warmup):
    from oracle import mip360_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.init_state_dict(seed=0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=2e-3, weight_decay=1e-5)
    rays, pixels = synth_rays(sample_rays, 0)
    for _ in range(warmup):
        cpu_iteration(O, params, opt, rays, pixels, N_SAMPLES)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_iteration(O, params, opt, rays, pixels, N_SAMPLES)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(value=sample_rays / dt, ms_per_step=dt * 1e3, cores=cores, threads=torch.get_num_threads())


# -------------------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 100 ms from before the warm-up; `window()` keeps the samples whose timestamp falls
    inside the timed region (the recipe's clocks line of B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       os.environ.get("MIP360_SMI_MS", "100"), "-i", str(index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t_begin, t_end):
        import datetime
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), float(c[3]), c[4:8]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if t_begin <= r[0] <= t_end] or rows[-3:]
        if inside:
            sm = sorted(r[1] for r in inside)
            reasons = set()
            for r in inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in inside), power_w_max=max(r[3] for r in inside),
                       samples=len(inside), reasons=sorted(reasons))
        return out


# -------------------------------------------------------------------------------------------------
# per-kernel accounting (algorithmic work per launch, SURVEY §8d / BASELINE.md §4)
# -------------------------------------------------------------------------------------------------
def kernel_work(name, a):
    """(kind, units) for a profiled call: kind 'tensor' -> FLOP, 'hbm' -> bytes.  `a` = the int arguments."""
    if name == "mip360_linear_fwd":       # M, N, K, act, n_valid
        return "tensor", 2.0 * a[0] * a[1] * a[2]
    if name == "mip360_linear_dgrad":     # M, N, K, act
        return "tensor", 2.0 * a[0] * a[1] * a[2]
    if name == "mip360_linear_wgrad":     # M, N, K
        return "tensor", 2.0 * a[0] * a[1] * a[2]
    if name == "mip360_cast_ipe":         # t_stride, B, N, mode, add_origins : bf16 [N,64] output variant
        B, N = a[1], a[2]
        return "hbm", B * (48.0 + 4 * (N + 1) + 128 * N)
    if name == "mip360_resample":         # B, N, blur
        B, N = a[0], a[1]
        return "hbm", B * (4.0 * (3 * N + 2) + 4 * (N + 1))
    if name == "mip360_composite_fwd":    # B, N, head_mode, white
        B, N = a[0], a[1]
        return "hbm", B * (16.0 * N + 4 * (N + 1) + 12 + 20 + 4 * N)
    if name == "mip360_composite_bwd":
        B, N = a[0], a[1]
        return "hbm", B * (16.0 * N + 4 * (N + 1) + 12 + 16 + 4 * N + 16 * N)
    if name in ("mip360_density_to_weight_fwd", "mip360_density_to_weight_bwd"):
        B, N = a[0], a[1]
        return "hbm", B * (4.0 * N + 4 * (N + 1) + 12 + 4 * N + (4 * N if name.endswith("bwd") else 0))
    if name == "mip360_distortion_fwd":
        return "hbm", a[0] * (4.0 * (2 * a[1] + 1) + 4)
    if name == "mip360_distortion_bwd":
        return "hbm", a[0] * (4.0 * (2 * a[1] + 1) + 4 * a[1])
    if name == "mip360_bounds_per_ray":
        return "hbm", a[0] * (4.0 * (3 * a[1] + 2) + 4 * a[1])
    return "hbm", 0.0


def summarise_profile(prof, steps, pk):
    agg = {}
    for name, ints, e0, e1 in prof:
        key = (name, ints)
        ms = e0.elapsed_time(e1)
        d = agg.setdefault(key, [0.0, 0])
        d[0] += ms
        d[1] += 1
    total = sum(v[0] for v in agg.values())
    rows = []
    for (name, ints), (ms, cnt) in agg.items():
        kind, units = kernel_work(name, ints)
        avg = ms / cnt
        if kind == "tensor":
            ach = units / (avg * 1e-3) / 1e12
            rows.append(dict(kernel=name, args=list(ints), launches_per_step=cnt / steps, avg_ms=avg, share=ms / total,
                             bound="tensor", achieved=ach, peak=pk["tf_sustained"], unit="TFLOP/s",
                             frac=ach / pk["tf_sustained"]))
        else:
            ach = units / (avg * 1e-3) / 1e9
            rows.append(dict(kernel=name, args=list(ints), launches_per_step=cnt / steps, avg_ms=avg, share=ms / total,
                             bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"]))
    rows.sort(key=lambda r: -r["share"])
    return rows, total / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=16384, help="rays per GPU per iteration")
    ap.add_argument("--cpu-sample-rays", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel roofline table (JSON) here")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"train iteration (2 prop + 1 nerf sub-steps, AdamW), {args.rays} rays/GPU x {N_SAMPLES} samples, " \
               "default config.py widths, randomized, bf16 MLP"

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_cpu(args.cpu_sample_rays, args.steps, min(args.warmup, 1))
        sample = f"{args.cpu_sample_rays} rays of the same iteration per step (oracle port of the reference, fp32 torch CPU)"
        print(json.dumps({
            "impl": "reference", "metric": "train rays/s", "value": r["value"], "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL_DEBUG=VERSION/INFO print there)
        dist.init_process_group("nccl", device_id=dev)
    from mipnerf360_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):  # fresh checkout on a box with nvcc: build once (rank 0), never fall back
        if rank == 0:
            from mipnerf360_b200 import build
            build.build()
        if world > 1:
            dist.barrier()
    from mipnerf360_b200.model import mipNeRF360
    from mipnerf360_b200.train import Trainer

    torch.manual_seed(0)
    model = mipNeRF360(randomized=True, num_samples=N_SAMPLES, device=dev)  # identical init on every rank (seed 0)
    trainer = Trainer(model)
    rays, pixels = synth_rays(args.rays, 1000 + rank, device=dev)
    torch.manual_seed(1234 + rank)  # sampling draws differ per rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        trainer.step(rays, pixels)
    barrier()
    # timed region: EXACTLY K steps, CUDA events on the launching stream, barrier + synchronize on both sides
    t_begin = time.time()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        trainer.step(rays, pixels)
    e1.record()
    barrier()
    t_end = time.time()
    launches = _lib.launch_count()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    clk = clocks.stop(t_begin, t_end) if clocks else None
    # the same K steps again with every kernel launch bracketed by CUDA events (per-kernel durations for the
    # roofline; ~300 extra event records per step, so this pass is kept out of `value`)
    _lib.PROFILE = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        trainer.step(rays, pixels)
    p1.record()
    barrier()
    prof, _lib.PROFILE = _lib.PROFILE, None
    ms_step_instrumented = p0.elapsed_time(p1) / args.steps

    # end to end through the public API: pinned host rays/pixels in, losses out, every step
    rays_h, pixels_h = synth_rays(args.rays, 1000 + rank, pin=True)
    for _ in range(2):
        trainer.step_host(rays_h, pixels_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        trainer.step_host(rays_h, pixels_h)
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = sum(r.numel() * 4 for r in rays_h) + pixels_h.numel() * 4

    if rank == 0:
        pk = peaks()
        rows, kernel_ms = summarise_profile(prof, args.steps, pk)
        if args.kernel_table:
            json.dump(dict(ms_per_step=ms_step_instrumented, kernel_ms_per_step=kernel_ms, kernels=rows), open(args.kernel_table, "w"),
                      indent=1)
        top = rows[0]
        # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel/shape from the committed ncu --set full capture
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f'{top["kernel"]}{tuple(top["args"][:3])}')
        roof = dict(bound=top["bound"], achieved=top["achieved"], peak=top["peak"], unit=top["unit"], frac=top["frac"],
                    traffic=traffic, kernel=f'{top["kernel"]}{tuple(top["args"])}', share_of_step=top["share"],
                    peak_source=pk["src"] + " (sustained bf16)" if top["bound"] == "tensor" else pk["src"])
        total_rays = args.rays * world
        out = {
            "metric": "train rays/s", "value": total_rays / (ms_step * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "global_rays_per_step": total_rays, "parallelism": f"ray-sharded dp{world}",
                       "l2": "per-layer activations (2.1 GB) exceed L2; no flush needed"},
            "clocks": clk,
            "e2e": {"value": total_rays / float(e2e_s), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 12},
            "gpu_launches": launches,
            "roofline": roof,
            "ms_per_step_with_kernel_events": ms_step_instrumented,
            "step_tflops": ITER_FLOP_PER_SAMPLE * N_SAMPLES * args.rays / (ms_step * 1e-3) / 1e12,
            "step_tensor_frac": ITER_FLOP_PER_SAMPLE * N_SAMPLES * args.rays / (ms_step * 1e-3) / 1e12 / pk["tf_sustained"],
        }
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu(args.cpu_sample_rays, 2, 1)
            out["cpu_baseline"] = {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": "port",
                                   "sample": f"{args.cpu_sample_rays} rays of the same iteration, 2 timed steps, "
                                             f"oracle port of the reference (fp32 torch CPU, {r['threads']} threads)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
