#!/usr/bin/env python
"""Time the tcgen05 GEMM entry points in isolation (CUDA events, inputs >> L2).  usage: python scripts/gemm_bench.py [M]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mipnerf360_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = "cuda"


def timeit(fn, iters=10, warm=3):
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


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


for (N, K) in [(1024, 1024), (256, 256), (1024, 64), (64, 1024)]:
    x, W, b = bf(M, K), bf(N, K) / K ** 0.5, torch.randn(N, device=dev)
    dY, Wt, y = bf(M, N), bf(K, N), bf(M, K).abs()
    fl = 2.0 * M * N * K / 1e12
    for name, fn in [
        ("fwd relu", lambda: ops.linear_fwd(x, W, b, 1)),
        ("fwd sigmoid", lambda: ops.linear_fwd(x, W, b, 2)),
        ("dgrad relu", lambda: ops.linear_dgrad(dY, Wt, y, 1)),
        ("wgrad +db", lambda: ops.linear_wgrad(dY, x)),
        ("wgrad no db", lambda: ops.linear_wgrad(dY, x, want_db=False)),
        ("torch matmul (cuBLAS) fwd", lambda: torch.matmul(x, W.T)),
        ("torch matmul (cuBLAS) wgrad", lambda: torch.matmul(dY.T, x)),
    ]:
        ms = timeit(fn)
        print(f"M={M} N={N:5d} K={K:5d} {name:28s} {ms:8.4f} ms  {fl / (ms * 1e-3):8.1f} TFLOP/s", flush=True)
