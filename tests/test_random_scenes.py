"""Differential tests on random scenes (tests/cases.py::random_scene): arbitrary affine instances
including handedness swaps, coincident geometry and instances (exact-t ties), big leaves, mixed
visibility, moving meshes with 1-3 segments, rays with random intervals / flags / unnormalised
directions.

CPU tier: restatement == reference headers (byte for byte); the host simulation of the product's
flattener + traversal code meets the parity rule.  GPU tier: the kernels through the C ABI."""
import numpy as np
import pytest

import cases
import parity

SEEDS = list(range(40))
GPU_SEEDS = list(range(16))


@pytest.mark.parametrize("seed", SEEDS)
def test_restatement_equals_reference_headers(orc, asref, seed):
    desc, rays = cases.random_scene(seed)
    o, r = orc.scene(desc), asref.scene(desc)
    a, ca = o.trace(rays, threads=2, counters=True)
    b, cb = r.trace(rays, threads=2, counters=True)
    assert a.tobytes() == b.tobytes()
    for k in ("rays", "instances_visited", "triangles_tested", "hits"):      # the header build counts these
        assert ca[k] == cb[k], k
    assert np.array_equal(o.trace_probe(rays, threads=2), r.trace_probe(rays, threads=2))
    assert (a["prim_type"] == 2).sum() > 5


@pytest.fixture(scope="module")
def sim():
    from hostsim import hostsim
    return hostsim.load()


@pytest.mark.parametrize("seed", SEEDS)
def test_host_simulation_of_product_code(sim, orc, seed):
    from hostsim import hostsim
    desc, rays = cases.random_scene(seed)
    o = orc.scene(desc)
    s = hostsim.SimScene(sim, desc)
    ref = o.trace(rays, threads=2)
    assert s.trace(rays, wide=False)[0].tobytes() == ref.tobytes()
    parity.compare_hits(o, rays, s.trace(rays, wide=True)[0], ref)
    pref = o.trace_probe(rays, threads=2)
    assert np.array_equal(s.trace_probe(rays, wide=False)[0], pref)
    parity.compare_probes(o, rays, s.trace_probe(rays, wide=True)[0], pref)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", GPU_SEEDS)
def test_kernels_on_random_scenes(orc, seed):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200.intersector import Intersector, TraceContext
    desc, rays = cases.random_scene(seed, n_rays=20000)
    o = orc.scene(desc)
    isect = Intersector(TraceContext(desc, device=0))
    ref, cref = o.trace(rays, threads=4, counters=True)
    isect.ctx.counters(reset=True)
    assert isect.trace(rays, exact=True, counters=True).tobytes() == ref.tobytes()
    c = isect.ctx.counters(reset=True)
    for k in ("rays", "assembly_nodes_visited", "instances_visited", "triangle_nodes_visited", "triangles_tested", "hits"):
        assert c[k] == cref[k], k
    parity.compare_hits(o, rays, isect.trace(rays), ref)
    parity.compare_hits(o, rays, isect.trace(rays, sort=True), ref)      # coincident geometry: ties may resolve differently
    pref = o.trace_probe(rays, threads=4)
    assert np.array_equal(isect.trace_probe(rays, exact=True), pref)
    parity.compare_probes(o, rays, isect.trace_probe(rays), pref)


def _static_rays(rays):
    from appleseed_b200.scene import RayBatch
    return RayBatch(rays.org, rays.dir, rays.tmin, rays.tmax, flags=rays.flags)


@pytest.mark.parametrize("seed", SEEDS[:20])
def test_refine_and_parents_on_random_static_scenes(orc, asref, seed):
    """refine_and_offset with arbitrary affine (also handedness-swapping) object and assembly
    instances and unnormalised ray directions: restatement == reference headers."""
    desc, rays = cases.random_scene(seed, moving=False)
    rays = _static_rays(rays)
    o, r = orc.scene(desc), asref.scene(desc)
    hits = o.trace(rays, threads=2)
    pa, pb = o.refine_offset(rays, hits, threads=2), r.refine_offset(rays, hits, threads=2)
    assert pa.tobytes() == pb.tobytes()
    h = hits["prim_type"] == 2
    child = rays.take(np.nonzero(h)[0])
    child.org = child.org + hits["t"][h][:, None] * child.dir          # from the hit point, onwards and backwards
    child.dir[1::2] *= -1.0
    child.tmin = np.zeros(len(child))
    assert o.trace_parents(child, pa[h], threads=2).tobytes() == r.trace_parents(child, pa[h], threads=2).tobytes()
    assert np.array_equal(o.trace_probe_parents(child, pa[h], threads=2), r.trace_probe_parents(child, pa[h], threads=2))


@pytest.mark.parametrize("seed", SEEDS[:20])
def test_host_build_of_refine_and_parent_code(sim, orc, seed):
    """The product's refine_core.h and parent_origin, compiled for the host, against the oracle."""
    from hostsim import hostsim
    desc, rays = cases.random_scene(seed, moving=False)
    rays = _static_rays(rays)
    o = orc.scene(desc)
    s = hostsim.SimScene(sim, desc)
    hits = o.trace(rays, threads=2)
    par = s.refine_offset(rays, hits)
    assert par.tobytes() == o.refine_offset(rays, hits, threads=2).tobytes()
    h = hits["prim_type"] == 2
    child = rays.take(np.nonzero(h)[0])
    child.org = child.org + hits["t"][h][:, None] * child.dir
    child.dir[1::2] *= -1.0
    child.tmin = np.zeros(len(child))
    ref = o.trace_parents(child, par[h], threads=2)
    assert s.trace_parents(child, par[h], wide=False).tobytes() == ref.tobytes()
    wide = s.trace_parents(child, par[h], wide=True)
    differs = np.nonzero(~(wide.view(np.uint8).reshape(len(ref), -1) == ref.view(np.uint8).reshape(len(ref), -1)).all(axis=1))[0]
    for k in ("t", "u", "v", "prim_type"):          # exact ties only (coincident triangles / instances)
        assert np.array_equal(wide[k][differs], ref[k][differs]), k
    pref = o.trace_probe_parents(child, par[h], threads=2)
    assert np.array_equal(s.trace_probe_parents(child, par[h], wide=False), pref)
    assert np.array_equal(s.trace_probe_parents(child, par[h], wide=True), pref)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", GPU_SEEDS[:10])
def test_refine_and_parents_kernels_on_random_static_scenes(orc, seed):
    from appleseed_b200.intersector import Intersector, TraceContext
    desc, rays = cases.random_scene(seed, n_rays=20000, moving=False)
    rays = _static_rays(rays)
    o = orc.scene(desc)
    isect = Intersector(TraceContext(desc, device=0))
    hits = o.trace(rays, threads=4)
    assert isect.trace(rays, exact=True).tobytes() == hits.tobytes()
    par = isect.refine_and_offset(rays, hits)
    assert par.tobytes() == o.refine_offset(rays, hits, threads=4).tobytes()
    h = hits["prim_type"] == 2
    child = rays.take(np.nonzero(h)[0])
    child.org = child.org + hits["t"][h][:, None] * child.dir
    child.dir[1::2] *= -1.0
    child.tmin = np.zeros(len(child))
    ref = o.trace_parents(child, par[h], threads=4)
    assert isect.trace_with_parents(child, par[h], exact=True).tobytes() == ref.tobytes()
    # orc_two_nearest knows nothing about parents, so the tie rule is applied by hand: wherever the
    # identity differs the two candidates must be an EXACT tie (duplicated triangles: same t, u, v).
    wide = isect.trace_with_parents(child, par[h])
    same = wide.view(np.uint8).reshape(len(ref), -1) == ref.view(np.uint8).reshape(len(ref), -1)
    differs = np.nonzero(~same.all(axis=1))[0]
    for k in ("t", "u", "v", "prim_type"):          # coincident instances tie exactly as well
        assert np.array_equal(wide[k][differs], ref[k][differs]), k
    assert len(differs) <= 0.02 * len(ref) + 5
    pref = o.trace_probe_parents(child, par[h], threads=4)
    assert np.array_equal(isect.trace_probe_with_parents(child, par[h], exact=True), pref)
    parity.compare_probes(o, child, isect.trace_probe_with_parents(child, par[h]), pref)
