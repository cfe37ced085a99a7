#!/usr/bin/env python
"""bench.py — train rays/s of one reference training iteration (train.py:51-82) on synthetic rays, plus render rays/s.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference ...                     # the reference algorithm on the host cores

Workload (BASELINE.json configs[1]): ONE 16384-ray batch per iteration, 64 samples per ray, default config.py
architecture (prop 58-256x4-1, nerf 58-1024x8-{1,3}), random-init weights, randomized sampling, bf16 MLP.
One "step" = 2 proposal sub-steps + 1 NeRF sub-step, each with its AdamW update, exactly the reference's
iteration.  N > 1: one process per GPU (torchrun); the 16384-ray batch is SHARDED over the ranks (16384/N rays each,
"scaling": "strong", as configs[1] / SURVEY §8e specify), gradients are all-reduced per layer bucket over NCCL and the
reference's batch-coupled scalars are all-reduced too; the sharded step is checked against the unsharded one before
timing.  The weak-scaling number (16384 rays per GPU) is reported next to it under "weak_scaling".

The same line carries "render": BASELINE.json configs[2] and [3] (1008x756 LLFF-shaped NDC frame, 4946x3286
garden-shaped unbounded frame) rendered through render.render_frame — device ray generation, chunk loop, to8b,
image gather at N > 1 and the device->host copy all inside the timed region — at chunks 65536 and 4096.

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
    from mipnerf360_b200.synthetic import generic_rays
    return generic_rays(B, seed, device=device, pin=pin)


# -------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) for the same iteration
# -------------------------------------------------------------------------------------------------
def cpu_iteration(O, params, opt, rays, pixels, N):
    """train.py:53-82 with the oracle's functions (fp32, torch CPU), AdamW included."""
    names_p = [k for k in params if k.startswith("prop_net")]
    names_n = [k for k in params if k.startswith("nerf_net")]
    for _ in range(2):
        t_hat, w_hat = O.prop_forward(params, rays, N, True)
        out = O.nerf_forward(params, rays, t_hat, w_hat, True)  # the reference builds this graph, then detaches (train.py:55-57)
        lp = O.Loss_prop(out[3].detach(), out[4].detach(), t_hat, w_hat)
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


def run_cpu(sample_rays, steps, warmup, literal=False):
    """The reference's algorithm (oracle port, fp32 torch CPU, all host threads) on a bounded sample of the workload.
    literal=True keeps the reference's per-sample autograd-Jacobian loop (parameterization.py:77-79) instead of its closed
    form: the algorithm exactly as the reference executes it, ~0.17 ms per sample and level."""
    from oracle import mip360_oracle as O
    O.LITERAL_CONTRACT = bool(literal)
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
    O.LITERAL_CONTRACT = False
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
    """(kind, units) for a profiled call: kind 'tensor' -> FLOP, 'hbm' -> bytes.  `a` = the int arguments of the C call."""
    if name == "mip360_linear_fwd":       # M, N, K, act, n_valid
        return "tensor", 2.0 * a[0] * a[1] * a[2]
    if name == "mip360_linear_fwd_head":  # M, N, K, act : trunk layer + the 4 head columns in its epilogue
        return "tensor", 2.0 * a[0] * a[1] * a[2] + 2.0 * a[0] * a[1] * 4
    if name == "mip360_mlp_fwd_fused_narrow":  # M, layers, n_valid (, 1 = activations saved): the whole proposal MLP
        return "tensor", float(a[0]) * PROP_FLOP_PER_SAMPLE
    if name == "mip360_head_bwd":         # M, N, act, ldw : trunk output read, its gradient written, 16 B/row of g
        return "hbm", a[0] * (4.0 * a[1] + 16)
    if name == "mip360_linear_dgrad":     # M, N, K, act
        return "tensor", 2.0 * a[0] * a[1] * a[2]
    if name == "mip360_linear_wgrad":     # M, N, K
        return "tensor", 2.0 * a[0] * a[1] * a[2]
    if name == "mip360_cast_ipe_x":       # t_stride, vd_dim, B, N, mode, flags, x_cols : bf16 [N, x_cols] output variant
        B, N, cols = a[2], a[3], a[6]
        return "hbm", B * (48.0 + 4 * (N + 1) + 2 * cols * N)
    if name == "mip360_level0_sample":    # (use_rng, stream,) B, N : writes the knots (reads 20 B/ray)
        B, N = a[-2], a[-1]
        return "hbm", B * (4.0 * (N + 1) + 8 + 12)
    if name == "mip360_resample_sample":  # (use_rng, stream,) B, N, blur : bins + weights in, knots out
        B, N = a[-3], a[-2]
        return "hbm", B * (4.0 * (3 * N + 2) + 12)
    if name == "mip360_composite_fwd_s":  # B, N, head_mode, white : + s_vals and t_shift rows
        B, N = a[0], a[1]
        return "hbm", B * (16.0 * N + 4 * (N + 1) + 12 + 20 + 4 * N + 8 * (N + 1) + 8)
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
    if name == "mip360_bounds":           # B, N : two knot rows + fine weights in, N totals out
        return "hbm", a[0] * (4.0 * (3 * a[1] + 2))
    if name in ("mip360_interlevel_fwd", "mip360_interlevel_bwd"):
        return "hbm", a[0] * a[1] * (8.0 if name.endswith("bwd") else 4.0)
    if name == "mip360_adamw_pack":       # entries, tiles, ... : p, g, m, v read, p, m, v, g written, bf16 W and W^T written
        return "hbm", a[1] * 1024 * (16.0 + 16 + 4)
    return "hbm", 0.0


def profile_key(name, ints):
    """The int arguments that identify a launch shape (drops per-call values: optimiser step, RNG stream id)."""
    if name == "mip360_adamw_pack":
        return ints[:2]
    if name == "mip360_level0_sample":
        return ints[-2:]
    if name == "mip360_resample_sample":
        return ints[-3:]
    return ints


def summarise_profile(prof, steps, pk):
    agg = {}
    for name, ints, e0, e1 in prof:
        ints = profile_key(name, ints)
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


GLOBAL_RAYS = 16384


def workload_config(n_gpus, global_rays, scaling):
    """The `config` object of the JSON line: identical in both arms (the reference arm runs a bounded sample of it)."""
    per = global_rays // n_gpus if scaling == "strong" else global_rays
    return {"workload": f"train iteration (2 prop + 1 nerf sub-steps, AdamW) on one {global_rays if scaling == 'strong' else per * n_gpus}-ray "
                        f"batch x {N_SAMPLES} samples, default config.py widths, randomized, bf16 MLP (BASELINE configs[1])",
            "global_rays_per_step": global_rays if scaling == "strong" else per * n_gpus, "rays_per_gpu": per,
            "parallelism": f"ray-sharded dp{n_gpus}",
            "l2": "activations per layer exceed L2 (>= 268 MB per layer at 2048 rays/GPU); no flush needed"}


def time_train(trainer, rays, pixels, steps, warmup, barrier, dev, world):
    """W warm-up steps, then EXACTLY K steps between CUDA events with a barrier + synchronize on both sides; max over ranks."""
    import torch.distributed as dist
    for _ in range(warmup):
        trainer.step(rays, pixels)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        trainer.step(rays, pixels)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / steps


def bench_render(model, dev, world, barrier, pk, chunk_sizes, cases):
    """configs[2]/[3]: whole frames through render.render_frame.  Host wall clock between barriers (the call ends with the
    device->host copy of the frame), max over ranks."""
    import torch.distributed as dist
    from mipnerf360_b200.render import render_frame
    rows = []
    flop_per_ray = N_SAMPLES * (PROP_FLOP_PER_SAMPLE + NERF_FLOP_PER_SAMPLE)
    for case in cases:
        h, w = case["height"], case["width"]
        # 128 = the reference CLI's default chunk (config.py:49): launch-latency territory, only on frames of <= 1 M rays
        sizes = list(chunk_sizes) + ([128] if h * w <= (1 << 20) and 128 not in chunk_sizes else [])
        for chunks in sizes:
            warm_h = max(3, min(h, (8 * chunks * world + w - 1) // w))  # a few chunks per rank
            render_frame(model, case["c2w"], warm_h, w, case["focal"], case["near"], case["far"], case["ndc"], chunks)
            barrier()
            t0 = time.perf_counter()
            rgb8, d, a = render_frame(model, case["c2w"], h, w, case["focal"], case["near"], case["far"], case["ndc"], chunks)
            barrier()
            sec = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(sec, op=dist.ReduceOp.MAX)
            sec = float(sec)
            n = h * w
            rows.append(dict(case=case["name"], rays=n, chunks=chunks, n_gpus=world, seconds=sec, rays_per_s=n / sec,
                             tensor_frac=n * flop_per_ray / sec / 1e12 / (pk["tf_sustained"] * world),
                             d2h_bytes=int(rgb8.nbytes + d.nbytes + a.nbytes), finite=bool((a == a).all()),
                             mean_acc=float(a.mean())))
            del rgb8, d, a
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=GLOBAL_RAYS, help="rays of the (global) batch per iteration")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: shard the --rays batch over the ranks (strong, configs[1]) or give every rank --rays rays")
    ap.add_argument("--cpu-sample-rays", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the additional weak-scaling measurement")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the captured iteration")
    ap.add_argument("--render-chunks", default="65536,4096")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel roofline table (JSON) here")
    args = ap.parse_args()

    if args.impl == "b200" and args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` without a launcher: start one process per GPU ourselves
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29541"),
                                   os.path.abspath(__file__)] + sys.argv[1:])

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        warm = args.warmup
        r = run_cpu(args.cpu_sample_rays, args.steps, warm)
        sample = (f"{args.cpu_sample_rays} rays of the same iteration per step; oracle port of the reference (fp32 torch CPU, "
                  f"{r['threads']} threads): /root/reference is Python and absent on the GPU box")
        print(json.dumps({
            "impl": "reference", "metric": "train rays/s", "value": r["value"], "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": args.scaling if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus, args.rays, args.scaling),
            "cpu_baseline": {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }), flush=True)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)  # NCCL_DEBUG is left as the caller set it
    from mipnerf360_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):  # fresh checkout on a box with nvcc: build once (rank 0), never fall back
        if rank == 0:
            from mipnerf360_b200 import build
            build.build()
        if world > 1:
            dist.barrier()
    from mipnerf360_b200.model import mipNeRF360
    from mipnerf360_b200.synthetic import garden_case, llff_case
    from mipnerf360_b200.train import Trainer, check_sharded_equals_unsharded

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    assert args.rays % world == 0, "--rays must be divisible by the number of GPUs"
    rays_per_gpu = args.rays // world if args.scaling == "strong" else args.rays

    # data-parallel correctness gate: the sharded step must equal the unsharded one (losses, gradients)
    dp_check = None
    if world > 1:
        dp_check = check_sharded_equals_unsharded(dev, rays_per_rank=max(256, 2048 // world))
        bad = {k: v for k, v in dp_check.items() if k != "world" and not v < 5e-3}
        assert not bad, f"sharded step differs from the unsharded step: {dp_check}"
        torch.cuda.empty_cache()

    torch.manual_seed(0)
    model = mipNeRF360(randomized=True, num_samples=N_SAMPLES, device=dev)  # identical init on every rank (seed 0)
    trainer = Trainer(model, graph=not args.no_graph)
    # the global batch is drawn once (seed 1000) and sharded; weak scaling gives rank r the batch of seed 1000 + r
    if args.scaling == "strong":
        g_rays, g_pixels = synth_rays(args.rays, 1000)
        sl = slice(rank * rays_per_gpu, (rank + 1) * rays_per_gpu)
        from mipnerf360_b200.intern.ray import Rays
        rays = Rays(*[r[sl].contiguous().to(dev) for r in g_rays])
        pixels = g_pixels[sl].contiguous().to(dev)
        rays_h = Rays(*[r[sl].contiguous().pin_memory() for r in g_rays])
        pixels_h = g_pixels[sl].contiguous().pin_memory()
    else:
        rays, pixels = synth_rays(rays_per_gpu, 1000 + rank, device=dev)
        rays_h, pixels_h = synth_rays(rays_per_gpu, 1000 + rank, pin=True)
    torch.manual_seed(1234 + rank)  # sampling draws differ per rank

    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(warmup):
        trainer.step(rays, pixels)
    barrier()
    # timed region: EXACTLY K steps, CUDA events on the launching stream, barrier + synchronize on both sides
    t_begin = time.time()
    _lib.reset_launch_count()
    replayed0 = trainer.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        trainer.step(rays, pixels)
    e1.record()
    barrier()
    t_end = time.time()
    launches = _lib.launch_count() + trainer.replayed_launches - replayed0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    clk = clocks.stop(t_begin, t_end) if clocks else None

    # end to end through the public API: pinned host rays/pixels in, losses out, every step
    # Every step copies its batch from pinned host memory and its three losses are read back on the host; the read of step
    # i is taken after step i+1 has been enqueued (step_host(wait=False)), so the GPU does not idle across the read-back.
    for _ in range(2):
        trainer.step_host(rays_h, pixels_h)
    barrier()
    t0 = time.perf_counter()
    pending, losses_seen = None, 0
    for _ in range(args.steps):
        h = trainer.step_host(rays_h, pixels_h, wait=False)
        if pending is not None:
            losses_seen += int(bool(torch.isfinite(pending.result()).all()))
        pending = h
    losses_seen += int(bool(torch.isfinite(pending.result()).all()))  # the last step's losses are on the host: the job is done
    e2e_wall = time.perf_counter() - t0
    barrier()
    e2e_s = torch.tensor([e2e_wall / args.steps], device=dev, dtype=torch.float64)  # max over ranks below
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = sum(r.numel() * 4 for r in rays_h) + pixels_h.numel() * 4

    # the same K steps again with every kernel launch bracketed by CUDA events (per-kernel durations for the
    # roofline; ~300 extra event records per step and one C call per GEMM instead of one per MLP, so this pass
    # is kept out of `value`)
    _lib.PROFILE = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        trainer.step(rays, pixels)
    p1.record()
    barrier()
    prof, _lib.PROFILE = _lib.PROFILE, None
    ms_step_instrumented = p0.elapsed_time(p1) / args.steps

    # N > 1: the weak-scaling figure (every rank a full 16384-ray batch) next to the strong-scaling headline
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak:
        w_rays, w_pixels = synth_rays(args.rays, 1000 + rank, device=dev)
        w_ms = time_train(trainer, w_rays, w_pixels, args.steps, 2, barrier, dev, world)
        weak = {"value": args.rays * world / (w_ms * 1e-3), "unit": "rays/s", "ms_per_step": w_ms, "rays_per_gpu": args.rays,
                "global_rays_per_step": args.rays * world}
        del w_rays, w_pixels

    pk = peaks()
    render_rows = None
    if not args.no_render:
        del trainer
        torch.cuda.empty_cache()
        model.eval()
        for net in (model, model.prop_net, model.nerf_net):
            net.randomized = False  # deterministic frames (the reference's eval() leaves the flags on, App. A7)
        chunk_sizes = [int(c) for c in args.render_chunks.split(",") if c]
        render_rows = bench_render(model, dev, world, barrier, pk, chunk_sizes, [llff_case(), garden_case()])

    if rank == 0:
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
                    peak_source=pk["src"] + " (sustained bf16)" if top["bound"] == "tensor" else pk["src"],
                    source="per-launch CUDA events of a second, instrumented pass over the same K steps (one C call per GEMM "
                           "instead of one per MLP: same kernels, same launch order; ms_per_step_with_kernel_events)")
        total_rays = rays_per_gpu * world
        flop_step = ITER_FLOP_PER_SAMPLE * N_SAMPLES * total_rays
        out = {
            "metric": "train rays/s", "value": total_rays / (ms_step * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, args.rays, args.scaling),
            "clocks": clk,
            "e2e": {"value": total_rays / float(e2e_s), "unit": "rays/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 12, "losses_read": losses_seen,
                    "readback": "every step's losses are read on the host, the read of step i after step i+1 was enqueued"},
            "gpu_launches": launches,
            "launch_mode": "eager" if args.no_graph else "one CUDA graph per iteration (kernels counted at capture)",
            "roofline": roof,
            "ms_per_step_with_kernel_events": ms_step_instrumented,
            "step_tflops": flop_step / (ms_step * 1e-3) / 1e12,
            "step_tensor_frac": flop_step / (ms_step * 1e-3) / 1e12 / (pk["tf_sustained"] * world),
        }
        if weak is not None:
            out["weak_scaling"] = weak
        if dp_check is not None:
            out["dp_check"] = dp_check
        if render_rows is not None:
            out["render"] = render_rows
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu(args.cpu_sample_rays, 2, 1)
            out["cpu_baseline"] = {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": "port",
                                   "sample": f"{args.cpu_sample_rays} rays of the same iteration, 2 timed steps, "
                                             f"oracle port of the reference (fp32 torch CPU, {r['threads']} threads)"}
            # the algorithm exactly as the reference executes it (per-sample autograd Jacobians): one step
            # of 16 rays = 6 levels x 1024 jacobian() calls, extrapolated linearly (the loop is O(rays x samples))
            lit = run_cpu(16, 1, 0, literal=True)
            out["cpu_baseline"]["literal_jacobian_loop"] = {
                "value": lit["value"], "unit": "rays/s", "sample": "16 rays, 1 step, per-sample jacobian() loop kept",
                "note": "what /root/reference does on this host; its own GPU path is launch-bound the same way (SURVEY A13)"}
        print(json.dumps(out), flush=True)
    if world > 1:
        # captured graphs hold NCCL work: drop them and drain the device before the communicator goes away, and do not
        # let a stuck teardown keep a finished benchmark alive
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
