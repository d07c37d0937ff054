"""Device-side ShadingPoint::refine_and_offset and the parent-shading-point origin rule
(asgpu_refine_and_offset, asgpu_trace_with_parents, asgpu_trace_probe_with_parents) against the
oracle: parent records byte-identical; exact kernels byte-identical, wide kernels under the
north-star rule (tests/parity.py)."""
import numpy as np
import pytest

import cases
import parity
from appleseed_b200 import scenes
from appleseed_b200.scene import RayBatch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from appleseed_b200 import intersector
    return intersector


def _static(rays):
    return RayBatch(rays.org, rays.dir, rays.tmin, rays.tmax, flags=rays.flags)


@pytest.mark.parametrize("name", ["cornell", "c2", "c3"])
def test_parents_match_the_oracle(engine, orc, name):
    desc, rays, _ = cases.CASES[name]()
    rays = _static(rays)
    o = orc.scene(desc)
    ctx = engine.TraceContext(desc, device=0)
    isect = engine.Intersector(ctx)
    hits = isect.trace(rays, exact=True)
    ref_hits = o.trace(rays, threads=4)
    assert hits.tobytes() == ref_hits.tobytes()
    par = isect.refine_and_offset(rays, hits)
    ref_par = o.refine_offset(rays, ref_hits, threads=4)
    assert par.tobytes() == ref_par.tobytes()

    h = hits["prim_type"] == 2
    mask, pts, nrm = scenes.hit_points_and_normals(desc, rays, hits)
    bounce = scenes.bounce_rays(pts, nrm, 5, offset=0.0)
    p = par[h]
    ref = o.trace_parents(bounce, p, threads=4)
    assert isect.trace_with_parents(bounce, p, exact=True).tobytes() == ref.tobytes()
    stats = parity.compare_hits(o, bounce, isect.trace_with_parents(bounce, p), ref)
    assert stats["identity_equal"] >= stats["rays"] - stats["tie_exempt"]
    assert int(((ref["prim_type"] == 2) & (ref["t"] < 1e-9)).sum()) == 0

    lights = pts.mean(axis=0, keepdims=True) + np.array([[0.3, 5.0, 0.2]])
    sh = scenes.shadow_rays(pts, lights, 9)
    pref = o.trace_probe_parents(sh, p, threads=4)
    assert np.array_equal(isect.trace_probe_with_parents(sh, p, exact=True), pref)
    parity.compare_probes(o, sh, isect.trace_probe_with_parents(sh, p), pref)

    # Rays whose parent sits in another instance (or nowhere) behave like plain rays.
    none = np.zeros(len(bounce), dtype=p.dtype)
    none["assembly_instance"] = 0xFFFFFFFF
    assert isect.trace_with_parents(bounce, none, exact=True).tobytes() == isect.trace(bounce, exact=True).tobytes()


def test_refine_needs_source_geometry(engine):
    desc, rays, _ = cases.case_c2()
    rays = _static(rays)
    ctx = engine.TraceContext(desc, device=0, source_geometry=False)
    isect = engine.Intersector(ctx)
    hits = isect.trace(rays)
    with pytest.raises(engine.AsgpuError, match="source geometry"):
        isect.refine_and_offset(rays, hits)


@pytest.mark.parametrize("name", ["c4_msc1", "c4_msc2", "c4_msc3", "mixed"])
def test_moving_triangles_support_planes_and_parents(engine, asref, name):
    """Moving hits: the support plane is the triangle interpolated at the ray time (a member of the
    leaf visitor, triangletree.h:232, triangletree.cpp:1468-1469, 1496-1497) and the geometric normal
    comes from the source vertices interpolated between the poses around that time
    (shadingpoint.cpp:186-256).  Checked against oracle/_ref, byte for byte."""
    desc, rays, _ = cases.CASES[name]()
    r = asref.scene(desc)
    ctx = engine.TraceContext(desc, device=0)
    isect = engine.Intersector(ctx)
    ref_hits, ref_planes = r.trace_planes(rays, threads=4)
    hits = isect.trace(rays, exact=True)
    assert hits.tobytes() == ref_hits.tobytes()
    assert isect.support_planes(rays, hits).tobytes() == ref_planes.tobytes()
    par = isect.refine_and_offset(rays, hits)
    assert par.tobytes() == r.refine_offset(rays, ref_hits, threads=4).tobytes()

    h = hits["prim_type"] == 2
    pts = rays.org[h] + hits["t"][h][:, None] * rays.dir[h]
    nrm = par["geo_normal"][h] / np.linalg.norm(par["geo_normal"][h], axis=1, keepdims=True)
    bounce = scenes.bounce_rays(pts, nrm, 5, offset=0.0)
    bounce.time_absolute, bounce.time_normalized = rays.time_absolute[h], rays.time_normalized[h]
    p = par[h]
    ref = r.trace_parents(bounce, p, threads=4)
    assert isect.trace_with_parents(bounce, p, exact=True).tobytes() == ref.tobytes()
    wide = isect.trace_with_parents(bounce, p)
    same = wide["tri_slot"] == ref["tri_slot"]
    assert same.mean() > 0.999 and np.allclose(wide["t"][same], ref["t"][same], rtol=1e-5, atol=0)


def test_support_planes_of_static_scenes(engine, asref):
    desc, rays, _ = cases.case_c3()
    r = asref.scene(desc)
    isect = engine.Intersector(engine.TraceContext(desc, device=0))
    ref_hits, ref_planes = r.trace_planes(rays, threads=4)
    hits = isect.trace(rays)
    assert hits.tobytes() == ref_hits.tobytes()
    planes = isect.support_planes(rays, hits)
    assert planes.tobytes() == ref_planes.tobytes()
    assert np.all(planes[hits["prim_type"] != 2] == 0.0)
