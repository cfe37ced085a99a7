"""Ray-partitioned rendering (model.render_image, model.py:254-274, across the GPUs of one box).

Rays of an image are independent, so each rank renders one contiguous slab with the chunk loop of
`mipNeRF360.render_image` and the only collective is the final gather of the image (3 B/ray of uint8 colour plus
8 B/ray of distance and accumulation).  As in the reference, the batch-global contraction norm (App. A1) is taken
per chunk, so a sharded render equals a single-GPU render with the same chunk boundaries: slabs are multiples of
`chunks` rays (`shard_bounds(..., align=chunks)`).

`render_frame` is the whole of test.py:37-45 / video.py:33-39 for one camera on the device: pinhole (or LLFF-NDC) ray
generation per chunk (dataset.py:109-145, 364-387), the chunk loop, `to8b`, gather, one device->host copy.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from mipnerf360_b200 import ops
from mipnerf360_b200.intern.ray import namedtuple_map


def shard_bounds(n, rank, world, align=1):
    """Contiguous slab [lo, hi) of n items for `rank`: ceil(n/world) items per rank rounded up to a multiple of
    `align`, the last ranks ragged/empty."""
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class ChunkForward:
    """model(chunk) for a fixed chunk size as ONE CUDA-graph replay: the ~35 kernel launches of an inference forward
    and the Python between them collapse into one launch, which is what a render loop with small chunks (the reference's
    defaults are 4096, model.py:254, and 128, config.py:49) spends its time on.  Inputs are copied into static buffers;
    randomized sampling keeps working because the in-kernel generator reads its replay counter from device memory.
    Chunks of another size (the ragged last one) run eagerly.  Falls back to eager execution if capture is refused."""

    def __init__(self, model, chunk_rays):
        self.model, self.n = model, int(chunk_rays)
        self.graph = None
        self.failed = False
        self.warm = False

    def _capture(self, chunk):
        from mipnerf360_b200.intern.ray import Rays
        self.static_in = Rays(*[r.detach().clone() for r in chunk])
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g, capture_error_mode="thread_local"):
            ops.rng_advance(self.static_in.origins.device)
            self.static_out = self.model(self.static_in)
        self.graph = g

    def __call__(self, chunk):
        if chunk[0].shape[0] != self.n or self.failed or not chunk[0].is_cuda:
            with torch.no_grad():
                return self.model(chunk)
        if not self.warm:
            # the first full chunk runs eagerly (and counts): lazy kernel attributes, operand refresh, allocator pools
            self.warm = True
            with torch.no_grad():
                return self.model(chunk)
        if self.graph is None:
            try:
                self._capture(chunk)
            except RuntimeError:
                self.failed = True
                torch.cuda.synchronize()
                with torch.no_grad():
                    return self.model(chunk)
        for dst, src in zip(self.static_in, chunk):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


GRAPH_MAX_CHUNK = 16384  # rays: above this a chunk is GPU-bound (>= 15 ms of GEMMs) and a private graph pool only costs memory


def _chunk_runner(model, n, chunks, graph):
    """A callable chunk -> (rgb, dist, acc): graph replays when the loop is long enough to repay the capture and the
    chunks are small enough to be launch-bound."""
    if graph and n >= 4 * chunks and chunks <= GRAPH_MAX_CHUNK and torch.cuda.is_available():
        return ChunkForward(model, chunks)

    def eager(chunk):
        with torch.no_grad():
            return model(chunk)
    return eager


def render_rays(model, rays, chunks, graph=True):
    """Chunk loop of model.render_image on device-resident or host rays -> (rgb [n,3], dist [n], acc [n]) on device."""
    n = rays[0].shape[0]
    dev = next(model.parameters()).device
    rgb = torch.empty((n, 3), device=dev)
    d = torch.empty((n,), device=dev)
    a = torch.empty((n,), device=dev)
    run = _chunk_runner(model, n, chunks, graph)
    for i in range(0, n, chunks):
        chunk = namedtuple_map(lambda r: r[i:i + chunks].to(dev, non_blocking=True), rays)
        rgb[i:i + chunks], d[i:i + chunks], a[i:i + chunks] = run(chunk)
    return rgb, d, a


def gather_slabs(local, n, world, group=None, per=None):
    """all_gather of per-rank slabs (padded to `per` = ceil(n/world) rows) back into one [n, ...] tensor."""
    per = per or (n + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if local.is_cuda:
        out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=group)
    else:  # gloo (CPU tests of the host logic)
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.cat(parts)
    return out[:n]


def render_image_distributed(model, rays, height, width, chunks=4096, group=None, graph=True):
    """Every rank passes the full ray set (host tensors are fine) and receives the full image:
    (rgb [h,w,3] float, dist [h,w], acc [h,w]) on the device.  world_size 1 degenerates to render_rays."""
    world, rank = _world(group)
    n = rays[0].shape[0]
    lo, hi = shard_bounds(n, rank, world, align=chunks)
    mine = namedtuple_map(lambda r: r[lo:hi], rays)
    rgb, d, a = render_rays(model, mine, chunks, graph)  # chunks are independent: model() issues no collective
    if world > 1:
        per = shard_bounds(n, 0, world, align=chunks)[1]
        packed = gather_slabs(torch.cat([rgb, d[:, None], a[:, None]], dim=1), n, world, group, per=per)
        rgb, d, a = packed[:, :3], packed[:, 3], packed[:, 4]
    return rgb.reshape(height, width, 3), d.reshape(height, width), a.reshape(height, width)


def render_frame(model, cam_to_world, height, width, focal, near, far, ndc=False, chunks=4096, group=None,
                 to_host=True, graph=True):
    """One camera -> (uint8 [h,w,3], dist [h,w], acc [h,w]) like model.render_image (model.py:254-274), with the
    rays generated on the device chunk by chunk.  Under torch.distributed every rank renders its slab and
    receives the whole frame.  to_host: return NumPy arrays (one D2H copy per output), else device tensors."""
    world, rank = _world(group)
    dev = next(model.parameters()).device
    c2w = ops.f32c(torch.as_tensor(cam_to_world, dtype=torch.float32).to(dev))
    n = height * width
    lo, hi = shard_bounds(n, rank, world, align=chunks)
    m = hi - lo
    rgb = torch.empty((m, 3), device=dev)
    d = torch.empty((m,), device=dev)
    a = torch.empty((m,), device=dev)
    torch.cuda.nvtx.range_push("mip360/render_frame")
    run = _chunk_runner(model, m, chunks, graph)
    with torch.no_grad():
        for i in range(0, m, chunks):
            c = min(chunks, m - i)
            chunk = ops.generate_rays(c2w, height, width, focal, near, far, ndc=ndc, ray_begin=lo + i, ray_count=c)
            rgb[i:i + c], d[i:i + c], a[i:i + c] = run(chunk)
    torch.cuda.nvtx.range_pop()
    rgb8 = ops.to8b(rgb)  # the picture leaves the device (and crosses NVLink) as 3 B/pixel
    if world > 1:
        per = shard_bounds(n, 0, world, align=chunks)[1]
        rgb8 = gather_slabs(rgb8, n, world, group, per=per)
        da = gather_slabs(torch.stack([d, a], dim=1), n, world, group, per=per)
        d, a = da[:, 0], da[:, 1]
    rgb8, d, a = rgb8.reshape(height, width, 3), d.reshape(height, width), a.reshape(height, width)
    if to_host:
        return rgb8.cpu().numpy(), d.cpu().numpy(), a.cpu().numpy()
    return rgb8, d, a
