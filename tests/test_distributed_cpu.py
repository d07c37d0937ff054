"""Host-side logic of the N > 1 path on CPU: tile sharding and the scene-blob broadcast, with a
world_size-2 gloo process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from appleseed_b200 import distributed as D


def test_hilbert_is_a_bijection_with_unit_steps():
    n = 16
    x, y = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    d = D.hilbert_index(4, x.reshape(-1), y.reshape(-1))
    assert sorted(d) == list(range(n * n))
    order = np.argsort(d)
    px, py = x.reshape(-1)[order], y.reshape(-1)[order]
    assert (np.abs(np.diff(px)) + np.abs(np.diff(py)) == 1).all()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_tile_shards_partition_the_frame(world):
    w, h = 1920, 1080
    shards = [D.tile_shard(w, h, world, r) for r in range(world)]
    allpix = np.concatenate(shards)
    assert len(allpix) == w * h and len(np.unique(allpix)) == w * h
    sizes = [len(s) for s in shards]
    assert (max(sizes) - min(sizes)) / (w * h / world) < 0.02     # round-robin over tiles: near-perfect balance


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tile_id_shards_cover_the_same_pixels_as_the_pixel_shards(world):
    w, h, tile = 1920, 1080, 32
    tx = (w + tile - 1) // tile
    seen = []
    for r in range(world):
        ids = D.tile_ids_shard(w, h, world, r, tile)
        seen.append(ids)
        pix = []
        for t in ids.tolist():
            x0, y0 = (t % tx) * tile, (t // tx) * tile
            ys, xs = np.meshgrid(np.arange(y0, min(h, y0 + tile)), np.arange(x0, min(w, x0 + tile)), indexing="ij")
            pix.append((ys * w + xs).reshape(-1))
        assert np.array_equal(np.concatenate(pix), D.tile_shard(w, h, world, r, tile))
    every = np.concatenate(seen)
    assert len(every) == tx * ((h + tile - 1) // tile) == len(np.unique(every))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        payload = torch.arange(100003, dtype=torch.int64).to(torch.uint8) if rank == 0 else None
        got = D.broadcast_bytes(payload, src=0, device=torch.device("cpu"))
        ok = got.numel() == 100003 and bool((got == torch.arange(100003, dtype=torch.int64).to(torch.uint8)).all())
        # every rank traces only its own shard; shards are disjoint and need no exchange
        mine = D.tile_shard(256, 128, world, rank)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(mine)], dtype=torch.int64))
        total = int(sum(c.item() for c in counts))
        out[rank] = (ok, total)
    finally:
        dist.destroy_process_group()


def test_blob_broadcast_and_sharding_world_size_2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r][0] for r in range(world))
    assert all(out[r][1] == 256 * 128 for r in range(world))
