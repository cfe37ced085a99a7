#!/usr/bin/env python
"""A/B of the NeRF head: separate 64-column head GEMM vs. the head folded into the last trunk GEMM's epilogue
(interleaved on the same box, CUDA events, 1M rows = 16384 rays x 64 samples)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mipnerf360_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda")
    M, H = 1 << 20, 1024
    torch.manual_seed(0)
    x = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
    W = (torch.randn(H, H, device=dev) / 32).bfloat16()
    b = torch.randn(H, device=dev) * 0.1
    Wh = torch.zeros(64, H, device=dev)
    Wh[:4] = torch.randn(4, H, device=dev) / 32
    Whb = Wh.bfloat16()
    bh = torch.zeros(64, device=dev)
    w4 = Whb[:4].float().t().contiguous()
    out = torch.zeros(M, 4, device=dev)

    def separate():
        y, _ = ops.linear_fwd(x, W, b, 2)
        ops.linear_fwd(y, Whb, bh, 2, out_f32_cols=4, want_bf16=False)

    def fused_train():
        out.zero_()
        ops.linear_fwd_head(x, W, b, 2, w4, out, want_bf16=True)

    def fused_infer():
        out.zero_()
        ops.linear_fwd_head(x, W, b, 2, w4, out, want_bf16=False)

    def trunk_only():
        ops.linear_fwd(x, W, b, 2)

    y8, _ = ops.linear_fwd(x, W, b, 2)
    gz = torch.randn(M, 4, device=dev) * 1e-3
    Wht = Whb.t().contiguous()
    dWh, dbh = torch.zeros(64, H, device=dev), torch.zeros(64, device=dev)

    def bwd_separate():
        dzh = ops.head_grad_pack(gz, None, 0)
        ops.linear_wgrad(dzh, y8, dW=dWh, db=dbh)
        ops.linear_dgrad(dzh, Wht, y8, 2)

    def bwd_fused():
        ops.head_bwd(gz, w4, y8, 2, dWh, dbh)

    variants = dict(separate=separate, fused_train=fused_train, fused_infer=fused_infer, trunk_only=trunk_only,
                    bwd_separate=bwd_separate, bwd_fused=bwd_fused)
    for f in variants.values():
        for _ in range(3):
            f()
    torch.cuda.synchronize()
    times = {k: [] for k in variants}
    for rep in range(8):
        for k, f in variants.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                f()
            e1.record()
            torch.cuda.synchronize()
            times[k].append(e0.elapsed_time(e1) / 10)
    res = {k: dict(ms_median=sorted(v)[len(v) // 2], ms_min=min(v)) for k, v in times.items()}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "head_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
