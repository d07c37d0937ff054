"""Parity tests proper: the CUDA kernels, called through the C ABI (libasgpu.so), against the
CPU oracle on identical seeded inputs.

* EXACT kernels: hit records and probe results byte-identical to the oracle; traversal counters
  identical (same visit order as the reference).
* WIDE kernels: the north-star rule (tests/parity.py): identity exact unless the two nearest
  candidates tie within 1e-6 relative, t within 1e-5 relative, barycentrics within 1e-5 absolute.
Golden vectors generated from the reference's own headers (tests/golden) are checked too, so the
GPU results are tied to the reference and not just to the restatement."""
import hashlib
import os

import numpy as np
import pytest

import cases
import kat
import parity

pytestmark = pytest.mark.gpu

GOLDEN = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_hits.npz")))


@pytest.fixture(scope="module")
def engine():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200 import intersector
    return intersector


def make(engine, desc, **kw):
    ctx = engine.TraceContext(desc, device=0, **kw)
    return ctx, engine.Intersector(ctx)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_exact_kernels_bit_identical_to_oracle_and_golden(engine, orc, name):
    desc, rays, probes = cases.CASES[name]()
    o = orc.scene(desc)
    ctx, isect = make(engine, desc)
    ref, cref = o.trace(rays, threads=4, counters=True)
    ctx.counters(reset=True)
    got = isect.trace(rays, exact=True, counters=True)
    assert got.tobytes() == ref.tobytes()
    c = ctx.counters(reset=True)
    for k in ("rays", "assembly_nodes_visited", "instances_visited", "triangle_nodes_visited", "triangles_tested", "hits"):
        assert c[k] == cref[k], k
    assert c["kernel_launches"] >= 1
    pref = o.trace_probe(probes, threads=4)
    pgot = isect.trace_probe(probes, exact=True)
    assert np.array_equal(pgot, pref)
    # Golden vectors from the reference's own headers.
    assert hashlib.sha256(got.tobytes()).hexdigest() == str(GOLDEN[name + "_hits_sha256"])
    assert hashlib.sha256(pgot.tobytes()).hexdigest() == str(GOLDEN[name + "_probe_sha256"])


@pytest.mark.parametrize("name", list(cases.CASES))
def test_wide_kernels_meet_parity_rule(engine, orc, name):
    desc, rays, probes = cases.CASES[name]()
    o = orc.scene(desc)
    ctx, isect = make(engine, desc)
    stats = parity.compare_hits(o, rays, isect.trace(rays), o.trace(rays, threads=4))
    assert stats["identity_equal"] + stats["tie_exempt"] == len(rays)
    parity.compare_probes(o, probes, isect.trace_probe(probes), o.trace_probe(probes, threads=4))


def test_device_buffers_match_host_buffers(engine, orc):
    import torch
    desc, rays, probes = cases.case_c3()
    ctx, isect = make(engine, desc)
    dr = engine.DeviceRays.from_host(rays, "cuda:0")
    hits = torch.empty(len(rays) * engine.HIT_BYTES, dtype=torch.uint8, device="cuda:0")
    isect.trace_device(dr, hits)
    torch.cuda.synchronize()
    assert engine.hits_from_tensor(hits, len(rays)).tobytes() == isect.trace(rays).tobytes()
    dp = engine.DeviceRays.from_host(probes, "cuda:0")
    occ = torch.empty(len(probes), dtype=torch.uint8, device="cuda:0")
    isect.trace_probe_device(dp, occ)
    torch.cuda.synchronize()
    assert np.array_equal(occ.cpu().numpy(), isect.trace_probe(probes))


def test_reference_format_trees_from_the_oracle_flatten_to_the_same_results(engine, orc):
    """asgpu_scene_create consumes appleseed's own tree arrays: feed it the trees built by the
    checker (standing in for a live TriangleTree / AssemblyTree) instead of the product builder."""
    import ctypes as C
    from appleseed_b200 import _lib
    from oracle.oracle import AssemblyTreeView, TriangleTreeView
    desc, rays, _ = cases.case_mixed()
    o = orc.scene(desc)
    views = []
    for i in range(o.tree_count):
        v = TriangleTreeView()
        orc._get_tt(o.handle, i, C.byref(v))
        w = _lib.TriangleTreeView()
        for f, _t in _lib.TriangleTreeView._fields_:
            setattr(w, f, getattr(v, f))
        views.append(w)
    av = AssemblyTreeView()
    orc._get_at(o.handle, C.byref(av))
    n_items = int(av.item_count)
    inst = np.frombuffer((C.c_uint32 * n_items).from_address(av.item_assembly_instance), dtype=np.uint32)
    tree = np.frombuffer((C.c_uint32 * n_items).from_address(av.item_tree), dtype=np.uint32)
    items = (_lib.AssemblyItem * max(1, n_items))()
    for k in range(n_items):
        ai = desc.assembly_instances[int(inst[k])]
        items[k].parent_to_local[:] = ai.parent_to_local.reshape(-1).tolist()
        items[k].assembly_instance = int(inst[k])
        items[k].triangle_tree = int(tree[k])
        items[k].vis_flags = ai.vis_flags & 0xFFFFFFFF
    top = _lib.AssemblyTreeView()
    top.nodes = av.nodes
    top.items = C.cast(items, C.POINTER(_lib.AssemblyItem))
    top.node_count = av.node_count
    top.item_count = n_items
    ctx = engine.TraceContext.from_tree_views(views, top)
    isect = engine.Intersector(ctx)
    ref = o.trace(rays, threads=4)
    assert isect.trace(rays, exact=True).tobytes() == ref.tobytes()
    parity.compare_hits(o, rays, isect.trace(rays), ref)


def test_reference_known_answers_on_gpu(engine):
    # test_tracer.cpp:421-456, 954-981, 1016-1059; test_intersector.cpp:116-147; ray.h:49-53.
    _, isect = make(engine, kat.tracer_scene([2.0, 4.0]))
    for exact in (True, False):
        h = isect.trace(kat.x_ray(), exact=exact)
        assert h["t"][0] == 2.0 and h["assembly_instance"][0] == 0 and h["prim_type"][0] == 2
        assert isect.trace_probe(kat.x_ray(), exact=exact)[0] == 1
        from appleseed_b200.scene import RayBatch
        between = RayBatch(np.array([[2.0 + 1e-9, 0.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]), 0.0, 2.0 * (1.0 - 1.0e-6))
        assert isect.trace_probe(between, exact=exact)[0] == 0
        edge = RayBatch(np.zeros((2, 3)), np.array([[1.0, 0, 0], [1.0, 0, 0]]), np.array([2.0, 0.0]), np.array([10.0, 2.0]))
        assert list(isect.trace(edge, exact=exact)["prim_type"]) == [2, 0]       # tmin inclusive, tmax exclusive
    _, isect = make(engine, kat.tracer_scene([2.0], scale=0.5))
    for exact in (True, False):
        assert abs(isect.trace(kat.x_ray(), exact=exact)["t"][0] - 1.0) <= 1e-15
    _, isect = make(engine, kat.empty_bbox_scene())
    for exact in (True, False):
        h = isect.trace(kat.empty_bbox_ray(), exact=exact)
        assert h["prim_type"][0] == 0 and h["t"][0] == 2.0 and h["assembly_instance"][0] == 0xFFFFFFFF
        assert isect.trace_probe(kat.empty_bbox_ray(), exact=exact)[0] == 0


def test_empty_and_ragged_batches(engine, orc):
    desc, rays, _ = cases.case_c2(32, 5000)
    o = orc.scene(desc)
    ctx, isect = make(engine, desc)
    assert len(isect.trace(rays.slice(0, 0))) == 0 and len(isect.trace_probe(rays.slice(0, 0))) == 0
    for n in (1, 31, 33, 1000):                     # not multiples of the 32-ray queue chunk
        sub = rays.slice(0, n)
        assert isect.trace(sub, exact=True).tobytes() == o.trace(sub).tobytes()
        parity.compare_hits(o, sub, isect.trace(sub), o.trace(sub))


def test_host_pipeline_spans_several_chunks(engine, orc):
    # More rays than one staging chunk (2^20) so the three-stream pipeline wraps around.
    from appleseed_b200 import scenes
    desc = scenes.scene_c2(64)
    lo, hi = scenes.scene_bbox(desc)
    rays = scenes.uniform_sphere_rays(2_500_000, lo - 0.1, hi + np.array([0.1, 0.5, 0.1]), 41)
    o = orc.scene(desc)
    ctx, isect = make(engine, desc)
    ref = o.trace(rays, threads=8)
    assert isect.trace(rays, exact=True).tobytes() == ref.tobytes()
    parity.compare_hits(o, rays, isect.trace(rays), ref)
    assert np.array_equal(isect.trace_probe(rays, exact=True), o.trace_probe(rays, threads=8))


def test_blob_export_import_round_trip(engine, orc):
    desc, rays, probes = cases.case_c4(2)
    ctx, isect = make(engine, desc)
    blob = ctx.blob_tensor()
    assert blob.numel() == ctx.blob_size == ctx.info()["blob_bytes"]
    ctx2 = engine.TraceContext.from_blob(blob, adopt=True)
    isect2 = engine.Intersector(ctx2)
    assert isect2.trace(rays).tobytes() == isect.trace(rays).tobytes()
    assert isect2.trace(rays, exact=True).tobytes() == isect.trace(rays, exact=True).tobytes()
    assert np.array_equal(isect2.trace_probe(probes), isect.trace_probe(probes))
    assert ctx2.info() == ctx.info()


def test_corrupted_blob_is_refused_on_import(engine):
    """The import path runs the flattener's structural checks on the header and the small tables (a
    truncated or corrupted broadcast payload must not reach the kernels)."""
    import torch
    desc, rays, _ = cases.case_c4(1)
    ctx, isect = make(engine, desc)
    good = ctx.blob_tensor()
    torch.cuda.synchronize()
    header = good[:256].cpu().numpy().copy()
    trees_off = int(header.view(np.uint64)[5])              # BlobHeader::trees (after 8 uint32 + total_bytes)
    assert 0 < trees_off < good.numel()
    for field_offset in (0, 24, 40, 48, 56):                # TreeDesc::bnodes, tris, keys, wnodes, wtris
        bad = good.clone()
        huge = torch.tensor(np.array([good.numel() - 8], dtype=np.uint64).view(np.uint8), device=bad.device)
        bad[trees_off + field_offset: trees_off + field_offset + 8] = huge
        with pytest.raises(engine.AsgpuError, match="validation"):
            engine.TraceContext.from_blob(bad, adopt=True)
    with pytest.raises(engine.AsgpuError):
        engine.TraceContext.from_blob(good[: good.numel() // 2].clone(), adopt=True)      # truncated
    ok = engine.TraceContext.from_blob(good, adopt=True)
    assert engine.Intersector(ok).trace(rays).tobytes() == isect.trace(rays).tobytes()


def test_wide_only_and_exact_only_scenes(engine, orc):
    from appleseed_b200 import _lib
    desc, rays, _ = cases.case_c3()
    o = orc.scene(desc)
    ref = o.trace(rays, threads=4)
    ctx, isect = make(engine, desc, flags=_lib.SCENE_WIDE)
    parity.compare_hits(o, rays, isect.trace(rays), ref)
    with pytest.raises(engine.AsgpuError, match="exact layout"):
        isect.trace(rays, exact=True)
    ctx, isect = make(engine, desc, flags=_lib.SCENE_EXACT)
    assert isect.trace(rays, exact=True).tobytes() == ref.tobytes()
    with pytest.raises(engine.AsgpuError, match="wide layout"):
        isect.trace(rays)


def test_axis_parallel_and_negative_tmin_rays(engine, orc):
    from appleseed_b200 import scenes
    from appleseed_b200.scene import RayBatch
    desc = scenes.scene_c2(40)
    o = orc.scene(desc)
    _, isect = make(engine, desc)
    rng = np.random.default_rng(3)
    n = 20000
    org = np.stack([rng.uniform(-1.2, 1.2, n), rng.uniform(-0.5, 1.0, n), rng.uniform(-1.2, 1.2, n)], 1)
    axes = np.array([[0, -1, 0], [0, 1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1], [0, -1, -0.0], [-0.0, -1, 0]], dtype=np.float64)
    d = axes[rng.integers(0, len(axes), n)]
    org[::3, 0] = np.round(org[::3, 0] * 20) / 20
    org[::3, 2] = np.round(org[::3, 2] * 20) / 20
    rays = RayBatch(org, d, rng.choice([0.0, -0.7, -3.0], n), rng.choice([scenes.DBL_MAX, 2.0], n))
    ref = o.trace(rays, threads=4)
    assert isect.trace(rays, exact=True).tobytes() == ref.tobytes()
    parity.compare_hits(o, rays, isect.trace(rays), ref)
    parity.compare_probes(o, rays, isect.trace_probe(rays), o.trace_probe(rays, threads=4))


def test_full_size_c2_properties(engine, orc):
    """BASELINE config C2 at full size (999 698 triangles): oracle comparison on a sample,
    size-independent properties on everything."""
    from appleseed_b200 import scenes
    desc = scenes.scene_c2(707)
    ctx, isect = make(engine, desc)
    assert ctx.info()["triangle_count"] == 999_698
    lo, hi = scenes.scene_bbox(desc)
    rays = scenes.uniform_sphere_rays(4_000_000, lo - 0.1, hi + np.array([0.1, 0.6, 0.1]), 5)
    wide = isect.trace(rays)
    exact = isect.trace(rays, exact=True)
    # (1) the two independent kernels agree on hit / miss and distance everywhere
    assert np.array_equal(wide["prim_type"], exact["prim_type"])
    hit = exact["prim_type"] == 2
    assert np.allclose(wide["t"][hit], exact["t"][hit], rtol=1e-5, atol=0.0)
    assert (wide["tri_slot"] == exact["tri_slot"]).mean() > 0.9999
    # (2) hits lie inside the ray interval and on the reported triangle (barycentric bounds)
    assert (exact["t"][hit] >= rays.tmin[hit]).all() and (exact["t"][hit] < rays.tmax[hit]).all()
    assert (exact["u"][hit] >= 0).all() and (exact["v"][hit] >= 0).all() and (exact["u"][hit] + exact["v"][hit] <= 1 + 1e-6).all()
    # (3) probes agree with closest hit: occluded <=> a hit exists in the interval
    occ = isect.trace_probe(rays)
    assert np.array_equal(occ.astype(bool), hit)
    # (4) shortening the interval to just before the hit turns every probe into a miss
    short = rays.take(np.nonzero(hit)[0][:500_000])
    short.tmax = exact["t"][hit][:500_000] * (1.0 - 1e-9)
    assert isect.trace_probe(short).sum() == 0
    # (5) oracle on a sample
    idx = np.arange(0, len(rays), 16)            # 250 000 rays
    o = orc.scene(desc)
    sub = rays.take(idx)
    ref = o.trace(sub, threads=8)
    assert exact[idx].tobytes() == ref.tobytes()
    parity.compare_hits(o, sub, wide[idx], ref)


@pytest.mark.parametrize("msc", [1, 2, 3])
def test_moving_triangles_at_scale_time_slices_change_nothing(engine, orc, msc, monkeypatch):
    """C4-like scene (180 000 moving triangles): the time-sliced wide boxes are an acceleration
    only -- results equal those of the all-motion boxes, of the exact kernels, and of the oracle on
    a sample; rays exactly on slice boundaries and at both ends of the time axis included."""
    from appleseed_b200 import scenes
    desc = scenes.scene_c4(300, msc)
    lo, hi = scenes.scene_bbox(desc)
    rays = scenes.uniform_sphere_rays(600_000, lo - 0.05, hi + 0.05, 17 + msc, time=True)
    t = rays.time_normalized
    t[:4096] = (np.arange(4096) % 17).astype(np.float32) / np.float32(16.0)       # slice boundaries k / 16
    t[:4096] = np.minimum(t[:4096], np.float32(1.0) - np.float32(2.0 ** -24))
    t[4096:8192] = np.nextafter(t[:4096], np.float32(0.0))
    rays.time_absolute = rays.time_normalized = t
    ctx, isect = make(engine, desc)
    sliced = isect.trace(rays)
    exact = isect.trace(rays, exact=True)
    assert np.array_equal(sliced["prim_type"], exact["prim_type"])
    hit = exact["prim_type"] == 2
    assert hit.sum() > 100_000
    assert np.allclose(sliced["t"][hit], exact["t"][hit], rtol=1e-5, atol=0.0)
    assert (sliced["tri_slot"] == exact["tri_slot"]).mean() > 0.9999
    occ = isect.trace_probe(rays)
    assert np.array_equal(occ, isect.trace_probe(rays, exact=True))
    monkeypatch.setenv("ASGPU_TIME_SLICES", "0")
    ctx0, isect0 = make(engine, desc)
    assert ctx0.info()["wide_node_bytes"] < ctx.info()["wide_node_bytes"]
    plain = isect0.trace(rays)
    same = plain.tobytes() == sliced.tobytes()
    if not same:            # only exact-t ties may be resolved differently (different visit sets)
        diff = np.nonzero((plain["tri_slot"] != sliced["tri_slot"]) | (plain["t"] != sliced["t"]))[0]
        assert len(diff) < 20 and np.allclose(plain["t"][diff], sliced["t"][diff], rtol=1e-6, atol=0.0)
    assert np.array_equal(isect0.trace_probe(rays), occ)
    idx = np.arange(0, len(rays), 25)
    o = orc.scene(desc)
    sub = rays.take(idx)
    ref = o.trace(sub, threads=8)
    assert exact[idx].tobytes() == ref.tobytes()
    parity.compare_hits(o, sub, sliced[idx], ref)
    parity.compare_probes(o, sub, occ[idx], o.trace_probe(sub, threads=8))


def test_full_size_c3_properties(engine, orc):
    """BASELINE config C3 at full size (9 999 392 triangles x 64 assembly instances): the two
    independent kernels agree everywhere, probes agree with closest hit, oracle on a sample."""
    from appleseed_b200 import scenes
    desc = scenes.scene_c3(2236, 8)
    ctx, isect = make(engine, desc)
    info = ctx.info()
    assert info["triangle_count"] == 9_999_392 and info["instance_count"] == 64
    lo, hi = scenes.scene_bbox(desc)
    ext = hi - lo
    rays = scenes.uniform_sphere_rays(2_000_000, lo - 0.02 * ext, hi + 0.02 * ext, 23)
    wide = isect.trace(rays)
    exact = isect.trace(rays, exact=True)
    assert np.array_equal(wide["prim_type"], exact["prim_type"])
    hit = exact["prim_type"] == 2
    assert 0.1 < hit.mean() < 0.9
    assert np.allclose(wide["t"][hit], exact["t"][hit], rtol=1e-5, atol=0.0)
    same = (wide["tri_slot"] == exact["tri_slot"]) & (wide["assembly_instance"] == exact["assembly_instance"])
    assert same.mean() > 0.9999
    assert len(np.unique(exact["assembly_instance"][hit])) == 64             # every instance is reachable
    occ = isect.trace_probe(rays)
    assert np.array_equal(occ.astype(bool), hit)
    assert np.array_equal(isect.trace_probe(rays, exact=True), occ)
    # Parent shading points at scale: refined points sit within a few ulps of the hit point.
    par = isect.refine_and_offset(rays, exact)
    assert np.all(par["assembly_instance"][hit] == exact["assembly_instance"][hit]) and np.all(par["assembly_instance"][~hit] == 0xFFFFFFFF)
    # Oracle on every 8th ray: 250 000 closest-hit rays, as many shadow probes from their hit points
    # (Tracer::trace_between's tmax), and the parent records of the hits.
    idx = np.arange(0, len(rays), 8)
    o = orc.scene(desc)
    sub = rays.take(idx)
    ref = o.trace(sub, threads=16)
    assert exact[idx].tobytes() == ref.tobytes()
    stats = parity.compare_hits(o, sub, wide[idx], ref)
    assert stats["rays"] == 250_000
    assert par[idx].tobytes() == o.refine_offset(sub, ref, threads=16).tobytes()
    import bench
    probes = bench.shadow_rays_from(desc, sub, ref, 3)
    pref = o.trace_probe(probes, threads=16)
    assert 0.05 < pref.mean() < 0.95
    assert np.array_equal(isect.trace_probe(probes, exact=True), pref)
    parity.compare_probes(o, probes, isect.trace_probe(probes), pref)

    # A C5 frame on the same scene (480 x 270 x 1 spp, parents carried): every wavefront of the path
    # stream against the oracle's trace_parents / trace_probe_parents on the identical rays.
    from appleseed_b200 import wavefront
    args = type("A", (), {"width": 480, "height": 270, "spp": 1, "no_parents": False})()
    cfg = bench.c5_config(args, desc)
    ps = wavefront.PathStream(ctx, wavefront.PathStreamConfig(**cfg), queue_capacity=1 << 20)
    ps.capture(1 << 23)
    ps.render()
    caps, st = ps.captured(), ps.stats()
    ps.close()
    assert st["camera_rays"] == 480 * 270 and st["surface_hits"] > 50_000 and len(caps) == 8
    for c in caps:
        if c.kind == "closest":
            cref = o.trace_parents(c.rays, c.parents, threads=16)
            s5 = parity.compare_hits(o, c.rays, c.results, cref)
            assert s5["identity_equal"] >= s5["rays"] - s5["tie_exempt"]
            assert int(((c.results["prim_type"] == 2) & (c.results["t"] < 1e-9)).sum()) == 0
        else:
            parity.compare_probes(o, c.rays, c.results, o.trace_probe_parents(c.rays, c.parents, threads=16))


@pytest.mark.parametrize("msc", [1, 3])
def test_full_size_c4_properties(engine, orc, msc):
    """BASELINE config C4 at full size (2 000 000 moving triangles, 2 and 4 poses): the two
    independent kernels agree everywhere, probes agree with closest hit, oracle on a sample."""
    from appleseed_b200 import scenes
    desc = scenes.scene_c4(1000, msc)
    ctx, isect = make(engine, desc)
    info = ctx.info()
    assert info["triangle_count"] == 2_000_000 == info["moving_triangle_count"]
    lo, hi = scenes.scene_bbox(desc)
    ext = hi - lo
    rays = scenes.uniform_sphere_rays(2_000_000, lo - 0.02 * ext, hi + 0.02 * ext, 31 + msc, time=True)
    wide = isect.trace(rays)
    exact = isect.trace(rays, exact=True)
    assert np.array_equal(wide["prim_type"], exact["prim_type"])
    hit = exact["prim_type"] == 2
    assert 0.1 < hit.mean() < 0.9
    assert np.allclose(wide["t"][hit], exact["t"][hit], rtol=1e-5, atol=0.0)
    assert (wide["tri_slot"] == exact["tri_slot"]).mean() > 0.9999
    assert set(np.unique(exact["motion_segment"][hit])) == set(range(msc))        # every pose interval is used
    occ = isect.trace_probe(rays)
    assert np.array_equal(occ.astype(bool), hit)
    assert np.array_equal(isect.trace_probe(rays, exact=True), occ)
    idx = np.arange(0, len(rays), 10)
    o = orc.scene(desc)
    sub = rays.take(idx)
    ref = o.trace(sub, threads=16)
    assert exact[idx].tobytes() == ref.tobytes()
    parity.compare_hits(o, sub, wide[idx], ref)
    parity.compare_probes(o, sub, occ[idx], o.trace_probe(sub, threads=16))
    # Parent records and support planes of moving hits at scale.
    par = isect.refine_and_offset(sub, ref)
    assert par.tobytes() == o.refine_offset(sub, ref, threads=16).tobytes()
    assert isect.support_planes(sub, ref).tobytes() == o.support_planes(sub, ref, threads=16).tobytes()
    ctx.close()


@pytest.mark.parametrize("name", ["c3", "mixed"])
def test_extreme_and_non_finite_rays(engine, orc, name):
    desc, rays, _ = cases.CASES[name]()
    o = orc.scene(desc)
    ctx, isect = make(engine, desc)
    ok = cases.extreme_rays(rays.slice(0, 6000), cases.EXTREME_OK, 5)
    ref = o.trace(ok, threads=4)
    assert isect.trace(ok, exact=True).tobytes() == ref.tobytes()
    parity.compare_hits(o, ok, isect.trace(ok), ref)
    pref = o.trace_probe(ok, threads=4)
    assert np.array_equal(isect.trace_probe(ok, exact=True), pref)
    parity.compare_probes(o, ok, isect.trace_probe(ok), pref)
    # No meaning as rays (NaN, infinite direction, tmin = -inf): the exact kernels still repeat the
    # reference bit for bit, the wide kernels terminate.
    bad = cases.extreme_rays(rays.slice(0, 6000), cases.EXTREME_UNDEFINED, 6)
    assert isect.trace(bad, exact=True).tobytes() == o.trace(bad, threads=4).tobytes()
    assert np.array_equal(isect.trace_probe(bad, exact=True), o.trace_probe(bad, threads=4))
    assert len(isect.trace(bad)) == len(bad) and len(isect.trace_probe(bad)) == len(bad)
    ctx.close()
