"""Tiny driver for ncu captures of the per-ray / memory-bound kernels (4M rays, N = 64; 1M-row MLP tensors): each kernel
is launched twice.  usage (under gpurun): bash scripts/gpu_ncu_perray.sh TAG"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mipnerf360_b200 import ops
from mipnerf360_b200.synthetic import garden_case
B, N, dev = 1 << 22, 64, "cuda"
t = (torch.rand(B, N + 1, device=dev) * 0.3).cumsum_(-1).add_(0.1)
w = torch.rand(B, N, device=dev).mul_(2.0 / N)
t2 = (torch.rand(B, N + 1, device=dev) * 0.3).cumsum_(-1).add_(0.1)
dirs = torch.randn(B, 3, device=dev)
near, far = torch.full((B, 1), 0.1, device=dev), torch.full((B, 1), 10.0, device=dev)
raw = torch.randn(B, N, 4, device=dev)
hb = torch.zeros(64, device=dev)
g_rgb, g_w = torch.rand(B, 3, device=dev), torch.rand(B, N, device=dev)
g_raw = torch.empty_like(raw)
s = t / t[:, -1:]
g1 = torch.ones((), device=dev)
gw = torch.empty_like(w)
Bx = 1 << 20  # the fused encoder writes 8 KB per ray
o = torch.randn(Bx, 3, device=dev)
vd = ops.viewdir_enc(dirs[:Bx] / dirs[:Bx].norm(dim=-1, keepdim=True))
rad = torch.full((Bx, 1), 1e-3, device=dev)
case = garden_case(2048, 2048)
c2w = case["c2w"].to(dev)
# 1M-row MLP tensors for the one-pass head backward
M, H = 1 << 20, 1024
y8 = torch.rand(M, H, device=dev).bfloat16()
gz = torch.randn(M, 4, device=dev) * 1e-3
w4 = torch.randn(H, 4, device=dev) / 32
dWh, dbh = torch.zeros(64, H, device=dev), torch.zeros(64, device=dev)
for _ in range(2):
    nsq = torch.zeros(1, device=dev, dtype=torch.float64)
    t0 = ops.level0_t_vals(near, far, N, True, directions=dirs, norm_sq=nsq)
    nsq2 = torch.zeros(1, device=dev, dtype=torch.float64)
    ops.resample(t, w, True, 0.01, directions=dirs, norm_sq=nsq2)
    ops.bounds_batch_total(t, w, t2)
    ops.composite_heads(raw, t, dirs, -1.0, 0.001, False, near=near, far=far, head_bias=hb)
    ops.call("mip360_composite_bwd", raw.data_ptr(), None, t.data_ptr(), dirs.data_ptr(), B, N, 2, -1.0, 0.001, 0,
             g_rgb.data_ptr(), None, None, g_w.data_ptr(), None, None, g_raw.data_ptr(), hb.data_ptr())
    wts = ops.density_to_weight(t, w, dirs)
    ops.frustum_norm_sq(t.data_ptr(), t.data_ptr() + 4, N + 1, dirs, B, N)
    ops.distortion_per_ray(s, w)
    ops.call("mip360_distortion_bwd", s.data_ptr(), w.data_ptr(), B, N, g1.data_ptr(), gw.data_ptr())
    tot = torch.rand(N, device=dev, dtype=torch.float64)
    ops.interlevel_loss(w, bound_total=tot)
    ops.cast_ipe(t[:Bx], o, dirs[:Bx], rad, vd, norm_sq=nsq, want_x=True)
    ops.generate_rays(c2w, 2048, 2048, case["focal"], case["near"], case["far"])
    ops.head_bwd(gz, w4, y8, 2, dWh, dbh)
torch.cuda.synchronize()
