"""Multi-GPU plumbing for the path: ONE broadcast of the flattened scene, rays sharded by image
tile, no further collectives.

The reference shards the image into independent tile jobs over threads
(renderer/kernel/rendering/generic/genericframerenderer.cpp:365-379, tile size 32x32 from
renderer/modeling/frame/frame.cpp:1331-1345); here the same tiles are dealt to ranks, one process
per GPU, with ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def hilbert_index(order: int, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Index of (x, y) along a Hilbert curve covering a 2^order x 2^order grid."""
    x = x.astype(np.int64).copy()
    y = y.astype(np.int64).copy()
    d = np.zeros_like(x)
    s = 1 << (order - 1) if order > 0 else 0
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate the quadrant
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return d


def tile_grid(width: int, height: int, tile: int = 32) -> List[Tuple[int, int, int, int]]:
    """Tiles (x0, y0, x1, y1) of a frame in Hilbert order."""
    tx = (width + tile - 1) // tile
    ty = (height + tile - 1) // tile
    gx, gy = np.meshgrid(np.arange(tx), np.arange(ty), indexing="xy")
    gx, gy = gx.reshape(-1), gy.reshape(-1)
    order = max(1, int(np.ceil(np.log2(max(tx, ty, 2)))))
    rank = np.argsort(hilbert_index(order, gx, gy), kind="stable")
    return [(int(gx[i]) * tile, int(gy[i]) * tile, min(width, (int(gx[i]) + 1) * tile), min(height, (int(gy[i]) + 1) * tile)) for i in rank]


def tile_shard(width: int, height: int, world: int, rank: int, tile: int = 32) -> np.ndarray:
    """Pixel indices (y * width + x) owned by ``rank``: Hilbert-ordered tiles dealt round-robin."""
    tiles = tile_grid(width, height, tile)[rank::world]
    out = []
    for x0, y0, x1, y1 in tiles:
        ys, xs = np.meshgrid(np.arange(y0, y1), np.arange(x0, x1), indexing="ij")
        out.append((ys * width + xs).reshape(-1))
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)


def tile_ids_shard(width: int, height: int, world: int, rank: int, tile: int = 32) -> np.ndarray:
    """Row-major tile indices (ty * tiles_x + tx) owned by ``rank``, in Hilbert order: the tile list
    ``asgpu_path_stream_render`` takes.  Same deal as ``tile_shard``."""
    tx = (width + tile - 1) // tile
    return np.array([(y0 // tile) * tx + (x0 // tile) for x0, y0, _, _ in tile_grid(width, height, tile)[rank::world]], dtype=np.uint32)


def broadcast_bytes(buf: Optional["torch.Tensor"], src: int = 0, device=None, group=None) -> "torch.Tensor":
    """Broadcast a uint8 tensor whose size only ``src`` knows (two collectives: size, payload).
    Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    if device is None:
        device = buf.device if buf is not None else torch.device("cpu")
    size = torch.tensor([buf.numel() if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(size, src, group=group)
    if rank != src:
        buf = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src, group=group)
    return buf


def replicate_scene(ctx, src: int = 0, device: Optional[int] = None, group=None):
    """The single scene broadcast: ``ctx`` is the TraceContext on ``src`` (None elsewhere);
    returns a TraceContext on every rank."""
    import torch
    import torch.distributed as dist
    from .intersector import TraceContext
    rank = dist.get_rank(group)
    dev = torch.device("cuda", device if device is not None else torch.cuda.current_device())
    blob = ctx.blob_tensor() if rank == src else None
    blob = broadcast_bytes(blob, src, dev, group)
    return ctx if rank == src else TraceContext.from_blob(blob, adopt=True)
