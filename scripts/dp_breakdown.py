#!/usr/bin/env python
"""Where the data-parallel step spends its time (run under torchrun on N GPUs; writes gpurun_out/dp_breakdown_nN.json).

Times the strong-scaling iteration (16384 / N rays per rank) in four variants: complete; without the gradient
all-reduce; without the scalar collectives (contraction norm, proposal bounds, squared error); without both (= a
single GPU at that batch size).  The last three compute wrong results on purpose — they only locate the cost."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from mipnerf360_b200.model import mipNeRF360
    from mipnerf360_b200.synthetic import generic_rays
    from mipnerf360_b200.train import Trainer
    per = 16384 // world
    rays, pixels = generic_rays(per, 1000 + rank, device=dev)
    rows = []
    for graph in (True, False):
        for overlap in (True, False):
            for skip in ((), ("grads",), ("scalars",), ("grads", "scalars")):
                if not overlap and "grads" in skip:
                    continue
                torch.manual_seed(0)
                model = mipNeRF360(randomized=True, num_samples=64, device=dev)
                tr = Trainer(model, graph=graph, overlap=overlap)
                tr.debug_skip = set(skip)
                for _ in range(5):
                    tr.step(rays, pixels)
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    tr.step(rays, pixels)
                e1.record()
                dist.barrier()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                rows.append(dict(graph=graph, overlap=overlap, skip=list(skip), ms_per_step=float(ms)))
                if rank == 0:
                    print(rows[-1], flush=True)
                del tr, model
                torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(dict(world=world, rays_per_rank=per, rows=rows), open(os.path.join(ROOT, "gpurun_out", f"dp_breakdown_n{world}.json"), "w"), indent=1)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
