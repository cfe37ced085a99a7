"""Is the clean timed region slower because it runs first (power controller settling) or because of how it launches?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mipnerf360_b200 import _lib
from mipnerf360_b200.model import mipNeRF360
from mipnerf360_b200.train import Trainer

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = mipNeRF360(randomized=True, num_samples=64, device=dev)
trainer = Trainer(model)
rays, pixels = bench.synth_rays(16384, 1000, device=dev)


def timed(k, profile=False):
    if profile:
        _lib.PROFILE = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        trainer.step(rays, pixels)
    e1.record()
    torch.cuda.synchronize()
    _lib.PROFILE = None
    return e0.elapsed_time(e1) / k


for _ in range(6):
    trainer.step(rays, pixels)
t0 = time.time()
for rnd in range(3):
    for smi in (None, "100", None, "1000"):
        sampler = None
        if smi:
            os.environ["MIP360_SMI_MS"] = smi
            sampler = bench.ClockSampler(0)
            time.sleep(0.3)
        r = [timed(5) for _ in range(3)]
        print(f"t={time.time() - t0:5.1f}s  smi={smi}  " + "  ".join(f"{x:.2f}" for x in r), flush=True)
        if sampler:
            sampler.stop(0, 1e12)
