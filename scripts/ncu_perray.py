"""Tiny driver for ncu captures of the per-ray kernels (4M rays, N = 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mipnerf360_b200 import ops
B, N, dev = 1 << 22, 64, "cuda"
t = (torch.rand(B, N + 1, device=dev) * 0.3).cumsum_(-1).add_(0.1)
w = torch.rand(B, N, device=dev).mul_(2.0 / N)
jit = ops.draw_jitter(B, N + 1, dev)
t2 = (torch.rand(B, N + 1, device=dev) * 0.3).cumsum_(-1).add_(0.1)
for _ in range(3):
    ops.resample(t, w, True, 0.01, jitter=jit)
    ops.bounds_per_ray(t, w, t2)
    ops.distortion_per_ray(t, w)
torch.cuda.synchronize()
