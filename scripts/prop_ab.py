#!/usr/bin/env python
"""Same-box interleaved A/B of the proposal MLP forward (1M rows): layer-by-layer GEMM launches vs. the layer-fused kernel,
inference (nothing saved) and training (activations saved)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mipnerf360_b200 import _lib, mlp as MLP  # noqa: E402
from mipnerf360_b200.model import prop_net  # noqa: E402


def main():
    dev = torch.device("cuda")
    torch.manual_seed(0)
    net = prop_net(randomized=False, num_samples=64, device=dev)
    M = 1 << 20
    x = (torch.randn(M, 64, device=dev) * 0.7).bfloat16()

    def run(fused, grad):
        _lib.set_option(_lib.OPT_FUSED_NARROW, fused)
        if grad:
            MLP.mlp_apply(net._packed, x)
        else:
            with torch.no_grad():
                MLP.mlp_apply(net._packed, x)

    variants = {"layers_infer": (False, False), "fused_infer": (True, False), "layers_train": (False, True),
                "fused_train": (True, True)}
    for a in variants.values():
        for _ in range(3):
            run(*a)
    torch.cuda.synchronize()
    times = {k: [] for k in variants}
    for _ in range(8):
        for k, a in variants.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run(*a)
            e1.record()
            torch.cuda.synchronize()
            times[k].append(e0.elapsed_time(e1) / 10)
    _lib.set_option(_lib.OPT_FUSED_NARROW, True)
    res = {k: dict(ms_median=sorted(v)[len(v) // 2], ms_min=min(v)) for k, v in times.items()}
    flop = 2.0 * M * (58 * 256 + 3 * 256 * 256 + 256)
    for k in res:
        res[k]["tflops_algorithmic"] = flop / (res[k]["ms_median"] * 1e-3) / 1e12
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "prop_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
