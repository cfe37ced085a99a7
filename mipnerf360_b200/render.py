"""Ray-partitioned rendering (model.render_image, model.py:254-274, across the GPUs of one box).

Rays of an image are independent, so each rank renders one contiguous slab with the chunk loop of
`mipNeRF360.render_image` and the only collective is the final gather of rgb/dist/acc (20 B/ray).  As in the
reference, the batch-global contraction norm (App. A1) is taken per chunk, so a sharded render equals a
single-GPU render with the same chunk boundaries.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from mipnerf360_b200.intern.ray import namedtuple_map


def shard_bounds(n, rank, world):
    """Contiguous slab [lo, hi) of n items for `rank`: ceil(n/world) items per rank, the last ones ragged/empty."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def render_rays(model, rays, chunks):
    """Chunk loop of model.render_image on device-resident or host rays -> (rgb [n,3], dist [n], acc [n]) on device."""
    n = rays[0].shape[0]
    dev = next(model.parameters()).device
    rgbs, dists, accs = [], [], []
    with torch.no_grad():
        for i in range(0, n, chunks):
            chunk = namedtuple_map(lambda r: r[i:i + chunks].to(dev, non_blocking=True), rays)
            rgb, d, a = model(chunk)
            rgbs.append(rgb); dists.append(d); accs.append(a)
    if not rgbs:
        return (torch.empty(0, 3, device=dev), torch.empty(0, device=dev), torch.empty(0, device=dev))
    return torch.cat(rgbs), torch.cat(dists), torch.cat(accs)


def gather_slabs(local, n, world, group=None):
    """all_gather of per-rank slabs (padded to ceil(n/world) rows) back into one [n, ...] tensor."""
    per = (n + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat(out)[:n]


def render_image_distributed(model, rays, height, width, chunks=4096):
    """Every rank passes the full ray set (host tensors are fine) and receives the full image:
    (rgb [h,w,3] float, dist [h,w], acc [h,w]) on the device.  world_size 1 degenerates to render_rays."""
    import mipnerf360_b200.model as M
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    n = rays[0].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    mine = namedtuple_map(lambda r: r[lo:hi], rays)
    prev, M.SYNC_BATCH_STATS = M.SYNC_BATCH_STATS, False  # chunks are independent: no cross-rank norm exchange
    try:
        rgb, d, a = render_rays(model, mine, chunks)
    finally:
        M.SYNC_BATCH_STATS = prev
    if world > 1:
        packed = gather_slabs(torch.cat([rgb, d[:, None], a[:, None]], dim=1), n, world)
        rgb, d, a = packed[:, :3], packed[:, 3], packed[:, 4]
    return rgb.reshape(height, width, 3), d.reshape(height, width), a.reshape(height, width)
