#!/usr/bin/env python
"""Sustained-load behaviour of the big GEMMs: time per launch, SM clock and board power sampled through NVML while a
loop of 200 identical launches is in flight (the GPU is power-capped, so energy per launch decides the speed)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pynvml  # noqa: E402
import torch  # noqa: E402

from mipnerf360_b200 import _lib, ops  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
M, N, K, dev = 1 << 20, 1024, 1024, "cuda"
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)  # noqa: E731
x, W, b = bf(M, K), bf(N, K) / 32, torch.randn(N, device=dev)
dY, Wt, y = bf(M, N), bf(K, N) / 32, bf(M, K).abs()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
dW = torch.zeros(N, K, device=dev)
db = torch.zeros(N, device=dev)
outc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
dWc = torch.empty(N, K, device=dev, dtype=torch.bfloat16)
dYt = dY.T.contiguous()


def run(name, fn, iters=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    clk, pw = [], []
    while not e1.query():
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000)
        time.sleep(0.02)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    clk, pw = clk[len(clk) // 3:] or [0], pw[len(pw) // 3:] or [0]
    print(f"{name:34s} {ms:7.4f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TF/s  sm {sum(clk) / len(clk):6.0f} MHz  {sum(pw) / len(pw):6.0f} W"
          f"  {ms * sum(pw) / len(pw) / 1e3:6.3f} J/launch", flush=True)
    time.sleep(1.0)


for rep in range(2):
    run("fwd relu (pair)", lambda: ops.linear_fwd(x, W, b, 1, out=out) if False else ops.call(
        "mip360_linear_fwd", x.data_ptr(), W.data_ptr(), b.data_ptr(), M, N, K, 1, out.data_ptr(), None, 0))
    run("dgrad relu (pair)", lambda: ops.call("mip360_linear_dgrad", dY.data_ptr(), Wt.data_ptr(), y.data_ptr(), M, N, K, 1,
                                              out.data_ptr()))
    run("wgrad +db (pair)", lambda: ops.call("mip360_linear_wgrad", dY.data_ptr(), x.data_ptr(), M, N, K, dW.data_ptr(),
                                             db.data_ptr()))
    run("wgrad no db (pair)", lambda: ops.call("mip360_linear_wgrad", dY.data_ptr(), x.data_ptr(), M, N, K, dW.data_ptr(), None))
    _lib.set_option(_lib.OPT_CTA_PAIR, False)
    run("fwd relu (single CTA)", lambda: ops.call("mip360_linear_fwd", x.data_ptr(), W.data_ptr(), b.data_ptr(), M, N, K, 1,
                                                  out.data_ptr(), None, 0))
    run("wgrad no db (single CTA)", lambda: ops.call("mip360_linear_wgrad", dY.data_ptr(), x.data_ptr(), M, N, K, dW.data_ptr(),
                                                     None))
    _lib.set_option(_lib.OPT_CTA_PAIR, True)
    run("cuBLAS fwd  x @ W.T", lambda: torch.matmul(x, W.T, out=outc))
    run("cuBLAS wgrad dY.T @ x", lambda: torch.matmul(dY.T, x, out=dWc))
    run("cuBLAS wgrad dYt @ x (K-major A)", lambda: torch.matmul(dYt, x, out=dWc))
