"""torchrun probe: wall-clock per-iteration time of Trainer.step (device batch) vs Trainer.step_host (pinned host batch,
pipelined read-back) on the sharded 16384-ray batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from mipnerf360_b200.model import mipNeRF360
from mipnerf360_b200.synthetic import generic_rays
from mipnerf360_b200.train import Trainer

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = 16384 // world
torch.manual_seed(0)
model = mipNeRF360(randomized=True, device=dev)
tr = Trainer(model, graph=True)
rays_d, pix_d = generic_rays(B, 1000 + rank, device=dev)
rays_h, pix_h = generic_rays(B, 1000 + rank, pin=True)
for _ in range(5):
    tr.step(rays_d, pix_d)
torch.cuda.synchronize()
def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
K = 20
for rep in range(2):
    sync(); t0 = time.perf_counter()
    for _ in range(K): tr.step(rays_d, pix_d)
    torch.cuda.synchronize(); a = (time.perf_counter() - t0) / K * 1e3
    sync(); t0 = time.perf_counter()
    for _ in range(K): tr.step_host(rays_h, pix_h)
    torch.cuda.synchronize(); b = (time.perf_counter() - t0) / K * 1e3
    sync(); t0 = time.perf_counter()
    pend = None
    for _ in range(K):
        h = tr.step_host(rays_h, pix_h, wait=False)
        if pend is not None: pend.result()
        pend = h
    pend.result(); c = (time.perf_counter() - t0) / K * 1e3
    sync(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): tr.step(rays_d, pix_d)
    e1.record(); torch.cuda.synchronize(); d = e0.elapsed_time(e1) / K
    print(f"rank {rank} rep {rep}: step(dev) wall {a:.3f}  step_host blocking {b:.3f}  step_host pipelined {c:.3f}  step(dev) events {d:.3f} ms", flush=True)
if world > 1:
    dist.barrier(); torch.cuda.synchronize(); os._exit(0)
