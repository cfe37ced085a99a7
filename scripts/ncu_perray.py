"""Tiny driver for ncu captures of the per-ray kernels (4M rays, N = 64): each kernel is launched twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mipnerf360_b200 import ops
B, N, dev = 1 << 22, 64, "cuda"
t = (torch.rand(B, N + 1, device=dev) * 0.3).cumsum_(-1).add_(0.1)
w = torch.rand(B, N, device=dev).mul_(2.0 / N)
jit = ops.draw_jitter(B, N + 1, dev)
t2 = (torch.rand(B, N + 1, device=dev) * 0.3).cumsum_(-1).add_(0.1)
dirs = torch.randn(B, 3, device=dev)
raw = torch.rand(B, N, 4, device=dev)
g_rgb, g_w = torch.rand(B, 3, device=dev), torch.rand(B, N, device=dev)
g_raw = torch.empty_like(raw)
for _ in range(2):
    ops.resample(t, w, True, 0.01, jitter=jit)
    ops.bounds_per_ray(t, w, t2)
    ops.composite_heads(raw, t, dirs, -1.0, 0.001, False)
    ops.call("mip360_composite_bwd", raw.data_ptr(), None, t.data_ptr(), dirs.data_ptr(), B, N, 1, -1.0, 0.001, 0,
             g_rgb.data_ptr(), None, g_w.data_ptr(), None, None, g_raw.data_ptr())
    ops.frustum_norm_sq(t.data_ptr(), t.data_ptr() + 4, N + 1, dirs, B, N)
torch.cuda.synchronize()
