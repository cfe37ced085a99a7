"""A/B of the packed GEMM epilogues under sustained load, alternating on the same box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mipnerf360_b200 import _lib, ops
M, dev = 1 << 20, "cuda"
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
def bench(N, K):
    x, W, b = bf(M, K), bf(N, K) / K ** 0.5, torch.randn(N, device=dev)
    dY, Wt, y = bf(M, N), bf(K, N) / N ** 0.5, bf(M, K).abs()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    outk = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    fns = {"fwd relu": lambda: ops.call("mip360_linear_fwd", x.data_ptr(), W.data_ptr(), b.data_ptr(), M, N, K, 1, out.data_ptr(), None, 0),
           "fwd sigmoid": lambda: ops.call("mip360_linear_fwd", x.data_ptr(), W.data_ptr(), b.data_ptr(), M, N, K, 2, out.data_ptr(), None, 0),
           "dgrad relu": lambda: ops.call("mip360_linear_dgrad", dY.data_ptr(), Wt.data_ptr(), y.data_ptr(), M, N, K, 1, outk.data_ptr())}
    for name, fn in fns.items():
        res = {}
        for rep in range(2):
            for packed in (1, 0):
                _lib.set_option(_lib.OPT_PACKED_EPILOGUE, bool(packed))
                for _ in range(5): fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(150): fn()
                e1.record(); torch.cuda.synchronize()
                res.setdefault(packed, []).append(e0.elapsed_time(e1) / 150)
                time.sleep(0.5)
        p, s = min(res[1]), min(res[0])
        print(f"N={N:5d} K={K:5d} {name:12s} packed {p:.4f} ms  scalar {s:.4f} ms  ({100 * (p / s - 1):+.1f} %)", flush=True)
    _lib.set_option(_lib.OPT_PACKED_EPILOGUE, True)
for N, K in ((1024, 1024), (256, 256), (1024, 64)):
    bench(N, K)
