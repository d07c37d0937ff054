"""Wavefront queues and the path stream (include/asgpu.h, "Wavefront ray queues") on the GPU.

* every wavefront the stream traces (captured rays + results) matches the oracle on the identical
  ray set under the north-star parity rule (tests/parity.py);
* the queue bookkeeping is consistent: one probe per surface hit, one bounce per surface hit while
  depth < max_bounces, accumulators = recount from the captured results;
* size-independent properties: the image does not depend on the queue capacity (batching), on the
  split of tiles between renderers (tile sharding = the multi-GPU decomposition), nor on the
  kernel (exact vs wide)."""
import numpy as np
import pytest

import cases
import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200 import scenes, wavefront
    from appleseed_b200.intersector import TraceContext
    desc = scenes.scene_c3(48, 3)
    lo, hi = scenes.scene_bbox(desc)
    centre = 0.5 * (lo + hi)
    eye = centre + np.array([0.3, 1.6, 1.1]) * np.linalg.norm(hi - lo) * 0.45
    lights = np.array([[lo[0], hi[1] + 2.0, lo[2]], [hi[0], hi[1] + 2.0, lo[2]], [lo[0], hi[1] + 2.0, hi[2]], [hi[0], hi[1] + 2.0, hi[2]]])
    cfg = dict(width=80, height=56, spp=3, camera_to_world=wavefront.look_at(eye, centre), lights=lights, max_bounces=3,
               tile_size=16, seed=11, offset_eps=1.0e-6 * float(np.linalg.norm(hi - lo)))
    ctx = TraceContext(desc, device=0)
    return desc, ctx, wavefront, cfg


def render(wavefront, ctx, cfg, capacity, tiles=None, capture=0, **over):
    ps = wavefront.PathStream(ctx, wavefront.PathStreamConfig(**{**cfg, **over}), queue_capacity=capacity)
    if capture:
        ps.capture(capture)
    ps.render(tiles)
    img, stats = ps.image(), ps.stats()
    caps = ps.captured() if capture else []
    ps.close()
    return img, stats, caps


def test_every_wavefront_matches_the_oracle(setup, orc):
    desc, ctx, wavefront, cfg = setup
    o = orc.scene(desc)
    img, stats, caps = render(wavefront, ctx, cfg, 1 << 20, capture=1 << 22)
    assert len(caps) == 2 * (cfg["max_bounces"] + 1)          # one batch: (closest, probe) per depth
    spp, w = cfg["spp"], cfg["width"]
    recount = np.zeros_like(img)
    prev_hits = None
    total = {"closest": 0, "probe": 0}
    for c in caps:
        n = len(c.rays)
        total[c.kind] += n
        assert n > 0
        pixel = c.path_ids // spp
        if c.kind == "closest":
            ref = o.trace(c.rays, threads=4)
            s = parity.compare_hits(o, c.rays, c.results, ref)
            assert s["identity_equal"] == n - s["tie_exempt"]
            if c.depth == 0:
                assert n == cfg["width"] * cfg["height"] * spp
                assert len(np.unique(c.path_ids)) == n
                assert np.all(c.rays.flags == 1) and np.allclose(np.linalg.norm(c.rays.dir, axis=1), 1.0, atol=1e-12)
            else:
                assert n == prev_hits                         # one bounce per surface hit of the previous depth
            hit = c.results["prim_type"] == 2
            prev_hits = int(hit.sum())
            np.add.at(recount[..., 0].reshape(-1), pixel[hit], 1)
            np.add.at(recount[..., 2].reshape(-1), pixel[~hit], 1)
            h = c.results[hit]
            ident = (h["primitive_index"].astype(np.uint64) * 2654435761 + h["object_instance_index"].astype(np.uint64) * 0x9E3779B1
                     + h["assembly_instance"].astype(np.uint64) * 0x85EBCA6B + h["tri_slot"].astype(np.uint64)) & 0xFFFFFFFF
            np.add.at(recount[..., 3].reshape(-1), pixel[hit], ident.astype(np.uint32))
        else:
            assert n == prev_hits                             # one shadow probe per path vertex
            ref = o.trace_probe(c.rays, threads=4)
            parity.compare_probes(o, c.rays, c.results, ref)
            assert np.all(c.rays.flags == 4)
            np.add.at(recount[..., 1].reshape(-1), pixel[c.results == 0], 1)
    assert np.array_equal(img, recount)
    assert stats["camera_rays"] + stats["bounce_rays"] == total["closest"]
    assert stats["probe_rays"] == total["probe"] == stats["surface_hits"]
    assert stats["surface_hits"] == int(img[..., 0].sum()) and stats["escaped"] == int(img[..., 2].sum())
    assert stats["unoccluded"] == int(img[..., 1].sum())
    assert stats["wavefronts"] == cfg["max_bounces"] + 1
    assert stats["kernel_launches"] == 1 + 4 * (cfg["max_bounces"] + 1)


def test_bounce_rays_start_on_the_surface_they_hit(setup):
    desc, ctx, wavefront, cfg = setup
    _, _, caps = render(wavefront, ctx, cfg, 1 << 20, capture=1 << 22)
    closest = [c for c in caps if c.kind == "closest"]
    probes = [c for c in caps if c.kind == "probe"]
    for parent, child, probe in zip(closest[:-1], closest[1:], probes[:-1]):
        hit = parent.results["prim_type"] == 2
        pts = parent.rays.org[hit] + parent.results["t"][hit][:, None] * parent.rays.dir[hit]
        by_path = dict(zip(parent.path_ids[hit].tolist(), range(int(hit.sum()))))
        idx = np.array([by_path[p] for p in child.path_ids.tolist()])
        d = np.linalg.norm(child.rays.org - pts[idx], axis=1)
        assert np.allclose(d, cfg["offset_eps"], rtol=1e-6, atol=1e-12)
        assert np.allclose(np.linalg.norm(child.rays.dir, axis=1), 1.0, atol=1e-12)
        # Cosine-weighted about the offset normal: every direction leaves the surface.
        nrm = (child.rays.org - pts[idx]) / d[:, None]
        assert np.all(np.einsum("ij,ij->i", nrm, child.rays.dir) >= -1e-12)
        # The probe of the same vertex starts at the same point and stops just short of its light.
        pidx = np.array([by_path[p] for p in probe.path_ids.tolist()])
        assert np.allclose(np.linalg.norm(probe.rays.org - pts[pidx], axis=1), cfg["offset_eps"], rtol=1e-6, atol=1e-12)
        reach = probe.rays.org + probe.rays.dir * (probe.rays.tmax / (1.0 - 1.0e-6))[:, None]
        dist = np.min(np.linalg.norm(reach[:, None, :] - np.asarray(cfg["lights"])[None], axis=2), axis=1)
        assert np.all(dist < 1e-9)


def test_image_is_independent_of_batching_sharding_and_kernel(setup):
    desc, ctx, wavefront, cfg = setup
    full, stats, _ = render(wavefront, ctx, cfg, 1 << 20)
    per_tile = cfg["tile_size"] ** 2 * cfg["spp"]
    small, stats_small, _ = render(wavefront, ctx, cfg, 3 * per_tile)           # many batches
    assert np.array_equal(full, small)
    assert stats_small["wavefronts"] > stats["wavefronts"]
    for k in ("camera_rays", "bounce_rays", "probe_rays", "surface_hits", "escaped", "unoccluded"):
        assert stats[k] == stats_small[k], k
    # Tile sharding (the multi-GPU decomposition): shards render disjoint pixels that add up to the frame.
    from appleseed_b200.distributed import tile_ids_shard
    parts = [render(wavefront, ctx, cfg, 1 << 20, tiles=tile_ids_shard(cfg["width"], cfg["height"], 3, r, cfg["tile_size"]))[0] for r in range(3)]
    assert np.array_equal(full, parts[0] + parts[1] + parts[2])
    touched = [(p.reshape(-1, 4).sum(axis=1) > 0) for p in parts]
    assert not np.any(touched[0] & touched[1]) and not np.any(touched[1] & touched[2])
    # The exact kernels (reference visit order) produce the same image.
    exact, _, _ = render(wavefront, ctx, cfg, 1 << 20, exact=True)
    assert np.array_equal(full, exact)


def test_queue_api_round_trip(setup, orc):
    import torch
    desc, ctx, wavefront, cfg = setup
    from appleseed_b200.intersector import HIT_BYTES, hits_from_tensor
    _, rays, probes = cases.case_c3()
    rays = rays.slice(0, 5000)
    q = wavefront.RayQueue(ctx, 8192)
    assert len(q) == 0
    q.push(rays.slice(0, 2000))
    q.push(rays.slice(2000, 5000))
    assert len(q) == 5000
    hits = torch.empty(q.capacity * HIT_BYTES, dtype=torch.uint8, device="cuda:0")
    q.trace(hits)
    torch.cuda.synchronize()
    got = hits_from_tensor(hits, 5000)
    from appleseed_b200.intersector import Intersector
    assert got.tobytes() == Intersector(ctx).trace(rays).tobytes()
    with pytest.raises(Exception, match="overflow"):
        q.push(rays)
    q.reset()
    assert len(q) == 0
    q.close()


def test_path_stream_rejects_bad_input(setup):
    desc, ctx, wavefront, cfg = setup
    from appleseed_b200.intersector import AsgpuError
    with pytest.raises(AsgpuError, match="capacity"):
        wavefront.PathStream(ctx, wavefront.PathStreamConfig(**cfg), queue_capacity=10)
    with pytest.raises(AsgpuError, match="light_count"):
        wavefront.PathStream(ctx, wavefront.PathStreamConfig(**{**cfg, "lights": np.zeros((9, 3))}), queue_capacity=1 << 16)
    ps = wavefront.PathStream(ctx, wavefront.PathStreamConfig(**cfg), queue_capacity=1 << 16)
    with pytest.raises(AsgpuError, match="range"):
        ps.render([ps.tile_count])
    ps.close()


def test_stream_with_parent_shading_points(setup, orc):
    """ASGPU_STREAM_PARENTS: child rays start at the hit point and carry the refined parent record."""
    desc, ctx, wavefront, cfg = setup
    o = orc.scene(desc)
    img, stats, caps = render(wavefront, ctx, cfg, 1 << 20, capture=1 << 22, parents=True)
    closest = [c for c in caps if c.kind == "closest"]
    probes = [c for c in caps if c.kind == "probe"]
    assert np.all(closest[0].parents["assembly_instance"] == 0xFFFFFFFF)       # camera rays have no parent
    for d, (parent, probe) in enumerate(zip(closest, probes)):
        # The wavefront's own results, parents honoured.
        ref = o.trace_parents(parent.rays, parent.parents, threads=4)
        parity.compare_hits(o, parent.rays, parent.results, ref)
        assert int(((parent.results["prim_type"] == 2) & (parent.results["t"] < 1e-9)).sum()) == 0     # no self-intersection
        # What the children carry = refine_and_offset of these hits, bit for bit.
        hit = parent.results["prim_type"] == 2
        refined = o.refine_offset(parent.rays, parent.results, threads=4)
        by_path = dict(zip(parent.path_ids.tolist(), range(len(parent.path_ids))))
        children = [probe] + ([closest[d + 1]] if d + 1 < len(closest) else [])
        for child in children:
            idx = np.array([by_path[p] for p in child.path_ids.tolist()])
            assert child.parents.tobytes() == refined[idx].tobytes()
            pts = parent.rays.org[idx] + parent.results["t"][idx][:, None] * parent.rays.dir[idx]
            assert np.allclose(child.rays.org, pts, rtol=0, atol=1e-12)                                # no epsilon offset
        pref = o.trace_probe_parents(probe.rays, probe.parents, threads=4)
        parity.compare_probes(o, probe.rays, probe.results, pref)
    # generate + per depth: closest trace, shade (refine_and_offset inside), probe trace, accumulate.
    assert stats["kernel_launches"] == 1 + 4 * (cfg["max_bounces"] + 1)
    # Same image from the exact kernels.
    exact, _, _ = render(wavefront, ctx, cfg, 1 << 20, parents=True, exact=True)
    assert np.array_equal(img, exact)


def test_refine_in_a_kernel_of_its_own_gives_the_same_stream(setup, monkeypatch):
    """ASGPU_FUSE_REFINE=0: refine_offset_kernel + shade_kernel instead of the fused shade kernel --
    same image, same captured parent records, one more launch per depth."""
    desc, ctx, wavefront, cfg = setup
    img, stats, caps = render(wavefront, ctx, cfg, 1 << 20, capture=1 << 22, parents=True)
    monkeypatch.setenv("ASGPU_FUSE_REFINE", "0")
    img2, stats2, caps2 = render(wavefront, ctx, cfg, 1 << 20, capture=1 << 22, parents=True)
    assert np.array_equal(img, img2)
    assert stats2["kernel_launches"] == stats["kernel_launches"] + cfg["max_bounces"] + 1
    assert len(caps) == len(caps2)
    for a, b in zip(caps, caps2):
        # Queue order is decided by atomics: compare ray by ray through the path ids.
        ia, ib = np.argsort(a.path_ids, kind="stable"), np.argsort(b.path_ids, kind="stable")
        assert a.kind == b.kind and np.array_equal(a.path_ids[ia], b.path_ids[ib])
        assert a.parents[ia].tobytes() == b.parents[ib].tobytes() and a.results[ia].tobytes() == b.results[ib].tobytes()


def test_queue_count_beyond_capacity_is_clamped(setup):
    """A producer that counts past the capacity (enqueue_slot keeps adding on overflow) must not make
    the trace kernels read or write beyond the queue's arrays: every consumer clamps the count."""
    import ctypes as C
    import torch
    desc, ctx, wavefront, cfg = setup
    from appleseed_b200 import _lib
    from appleseed_b200.intersector import HIT_BYTES, Intersector, hits_from_tensor
    from appleseed_b200.scene import CRays
    _, rays, probes = cases.case_c3()
    cap = 4096
    rays, probes = rays.slice(0, cap), probes.slice(0, cap)
    for exact in (False, True):
        q = wavefront.RayQueue(ctx, cap)
        q.push(rays)
        cr, ids, count = CRays(), C.c_void_p(), C.c_void_p()
        assert q.lib.asgpu_queue_device_arrays(q.handle, C.byref(cr), C.byref(ids), C.byref(count)) == 0
        # Overwrite the device-side count with capacity + 100000.
        from cuda.bindings import runtime as cudart
        big = np.array([cap + 100000], dtype=np.uint64)
        err, = cudart.cudaMemcpy(count.value, big.ctypes.data, 8, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice)
        assert int(err) == 0
        guard = 64 * HIT_BYTES
        hits = torch.full((cap * HIT_BYTES + guard,), 0xAB, dtype=torch.uint8, device="cuda:0")
        q.trace(hits, exact=exact)
        torch.cuda.synchronize()
        assert bool((hits[cap * HIT_BYTES:] == 0xAB).all()), "the trace wrote past the queue capacity"
        assert hits_from_tensor(hits, cap).tobytes() == Intersector(ctx).trace(rays, exact=exact).tobytes()
        assert len(q) == cap            # asgpu_queue_count clamps too
        occ = torch.full((cap + 64,), 0xAB, dtype=torch.uint8, device="cuda:0")
        q.trace_probe(occ, exact=exact)
        torch.cuda.synchronize()
        assert bool((occ[cap:] == 0xAB).all())
        q.close()


def test_queue_rejects_sort_flag(setup):
    import torch
    desc, ctx, wavefront, cfg = setup
    from appleseed_b200 import _lib
    from appleseed_b200.intersector import HIT_BYTES
    q = wavefront.RayQueue(ctx, 1024)
    hits = torch.empty(1024 * HIT_BYTES, dtype=torch.uint8, device="cuda:0")
    rc = q.lib.asgpu_trace_queue(ctx.handle, q.handle, hits.data_ptr(), _lib.TRACE_SORT, None)
    assert rc == -3 and "SORT" in _lib.last_error()
    q.close()


@pytest.fixture(scope="module")
def moving_setup():
    """The path stream over a DEFORMING mesh (C4-style, msc = 3) under two assembly instances, with a
    shutter interval: every ray of a path carries the path's time."""
    from appleseed_b200 import scenes, wavefront
    from appleseed_b200.intersector import TraceContext
    desc = scenes.scene_c4(40, 3)
    lo, hi = scenes.scene_bbox(desc)
    centre = 0.5 * (lo + hi)
    diag = float(np.linalg.norm(hi - lo))
    eye = centre + np.array([0.25, 1.3, 0.9]) * diag * 0.6
    lights = np.array([[lo[0], hi[1] + 2.0, lo[2]], [hi[0], hi[1] + 2.0, hi[2]]])
    cfg = dict(width=64, height=48, spp=4, camera_to_world=wavefront.look_at(eye, centre), lights=lights, max_bounces=2,
               tile_size=16, seed=3, offset_eps=1.0e-6 * diag, shutter_open=0.25, shutter_close=1.5, parents=True)
    return desc, TraceContext(desc, device=0), wavefront, cfg


def test_stream_over_moving_triangles_carries_ray_time(moving_setup, asref):
    desc, ctx, wavefront, cfg = moving_setup
    r = asref.scene(desc)
    img, stats, caps = render(wavefront, ctx, cfg, 1 << 20, capture=1 << 22)
    closest = [c for c in caps if c.kind == "closest"]
    probes = [c for c in caps if c.kind == "probe"]
    assert len(closest) == cfg["max_bounces"] + 1 and stats["surface_hits"] > 1000
    cam = closest[0]
    tn, ta = cam.rays.time_normalized, cam.rays.time_absolute
    assert tn.min() >= 0.0 and tn.max() < 1.0 and len(np.unique(tn)) > 0.9 * len(tn)
    one = np.float32(1.0)
    expect = (one - tn) * np.float32(cfg["shutter_open"]) + tn * np.float32(cfg["shutter_close"])     # foundation::lerp in float
    assert np.array_equal(ta, expect.astype(np.float32))
    time_of_path = dict(zip(cam.path_ids.tolist(), tn.tolist()))
    for d, (parent, probe) in enumerate(zip(closest, probes)):
        # Children inherit their path's time (pathtracer.h:764).
        for c in (parent, probe):
            assert np.array_equal(c.rays.time_normalized, np.array([time_of_path[p] for p in c.path_ids.tolist()], dtype=np.float32))
        ref = r.trace_parents(parent.rays, parent.parents, threads=4)
        same = parent.results["tri_slot"] == ref["tri_slot"]
        assert np.array_equal(parent.results["prim_type"], ref["prim_type"]) and same.mean() > 0.999
        assert parent.results[same].tobytes() == ref[same].tobytes()
        assert int(((parent.results["prim_type"] == 2) & (parent.results["t"] < 1e-9)).sum()) == 0     # no self-intersection
        refined = r.refine_offset(parent.rays, parent.results, threads=4)
        by_path = dict(zip(parent.path_ids.tolist(), range(len(parent.path_ids))))
        idx = np.array([by_path[p] for p in probe.path_ids.tolist()])
        assert probe.parents.tobytes() == refined[idx].tobytes()
        pref = r.trace_probe_parents(probe.rays, probe.parents, threads=4)
        assert (probe.results == pref).mean() > 0.999
    # A closed shutter freezes the mesh at time 0: a different image.
    still, _, _ = render(wavefront, ctx, cfg, 1 << 20, shutter_open=0.0, shutter_close=0.0)
    assert not np.array_equal(still, img)
    exact, _, _ = render(wavefront, ctx, cfg, 1 << 20, exact=True)
    assert np.array_equal(img, exact)


def test_stream_profile_times_every_trace_launch(setup):
    desc, ctx, wavefront, cfg = setup
    ps = wavefront.PathStream(ctx, wavefront.PathStreamConfig(**cfg), queue_capacity=1 << 20)
    ps.set_profiling(True)
    ps.render()
    prof = ps.profile()
    assert prof["closest_launches"] == cfg["max_bounces"] + 1 == prof["probe_launches"]
    assert prof["closest_ms"] > 0 and prof["probe_ms"] > 0 and prof["stage_ms"] > 0
    ps.clear()
    assert ps.profile()["closest_launches"] == 0
    ps.close()


def test_tile_read_back_equals_the_frame(setup):
    """asgpu_path_stream_read_tiles: the pixels of a rank's own tiles (what a tile-sharded renderer
    reads back) scatter into exactly the frame asgpu_path_stream_read_image returns; edge tiles
    that stick out of the image (height 56 = 3.5 tiles of 16) are zero-padded; bad lists are refused."""
    desc, ctx, wavefront, cfg = setup
    ps = wavefront.PathStream(ctx, wavefront.PathStreamConfig(**cfg), queue_capacity=1 << 18)
    n_tiles = ps.tile_count
    mine = np.arange(n_tiles, dtype=np.uint32)[1::2][::-1].copy()          # a shard, in an order of its own
    ps.render(mine)
    frame = ps.image()
    px = ps.image_tiles(mine)
    assert px.shape == (len(mine), cfg["tile_size"], cfg["tile_size"], 4)
    assert np.array_equal(ps.scatter_tiles(mine, px), frame)
    assert int(px.astype(np.uint64).sum()) == int(frame.astype(np.uint64).sum()) > 0
    others = np.arange(n_tiles, dtype=np.uint32)[0::2]
    assert not ps.image_tiles(others).any()                                # tiles nobody rendered
    assert ps.image_tiles(np.zeros(0, dtype=np.uint32)).shape[0] == 0
    with pytest.raises(Exception):
        ps.image_tiles(np.array([n_tiles], dtype=np.uint32))
    ps.close()
