"""ShadingPoint::refine_and_offset and the parent-shading-point origin rule in the oracle
(oracle.cpp) against the reference's own refining.h / raytrianglemt.h / transform.h compiled into
oracle/_ref (ref_driver.cpp): byte-identical records and hit results, plus the properties the
reference relies on (points on opposite sides of the support plane, no self-intersection)."""
import numpy as np
import pytest

import cases
from appleseed_b200 import scenes
from appleseed_b200.scene import RayBatch


def _static(rays):
    return RayBatch(rays.org, rays.dir, rays.tmin, rays.tmax, flags=rays.flags)


@pytest.mark.parametrize("name", ["cornell", "c2", "c3"])
def test_refine_and_parents_match_reference_headers(orc, asref, name):
    desc, rays, _ = cases.CASES[name]()
    rays = _static(rays)
    o, r = orc.scene(desc), asref.scene(desc)
    hits = o.trace(rays, threads=4)
    pa, pb = o.refine_offset(rays, hits, threads=4), r.refine_offset(rays, hits, threads=4)
    assert pa.tobytes() == pb.tobytes()
    h = hits["prim_type"] == 2
    assert h.sum() > 100
    assert np.all(pa["assembly_instance"][h] == hits["assembly_instance"][h]) and np.all(pa["assembly_instance"][~h] == 0xFFFFFFFF)
    # Child rays from the (un-offset) world hit point, parent supplied.
    mask, pts, nrm = scenes.hit_points_and_normals(desc, rays, hits)
    bounce = scenes.bounce_rays(pts, nrm, 5, offset=0.0)
    par = pa[h]
    ha, hb = o.trace_parents(bounce, par, threads=4), r.trace_parents(bounce, par, threads=4)
    assert ha.tobytes() == hb.tobytes()
    lights = pts.mean(axis=0, keepdims=True) + np.array([[0.3, 5.0, 0.2]])
    sh = scenes.shadow_rays(pts, lights, 9)
    assert np.array_equal(o.trace_probe_parents(sh, par, threads=4), r.trace_probe_parents(sh, par, threads=4))
    # The point of it all: without the parent the bounce rays hit their own triangle at t ~ 0.
    plain = o.trace(bounce, threads=4)
    self_hits = lambda x: int(((x["prim_type"] == 2) & (x["t"] < 1e-9)).sum())
    assert self_hits(plain) > 0.2 * len(bounce) and self_hits(ha) == 0


def test_offset_points_straddle_the_surface(orc):
    desc, rays, _ = cases.case_c2()
    rays = _static(rays)
    o = orc.scene(desc)
    hits = o.trace(rays, threads=4)
    h = hits["prim_type"] == 2
    p = o.refine_offset(rays, hits, threads=4)[h]
    n = p["geo_normal"] / np.linalg.norm(p["geo_normal"], axis=1, keepdims=True)
    mid = rays.org[h] + hits["t"][h][:, None] * rays.dir[h]            # identity instance: refine space == world
    assert np.all(np.einsum("ij,ij->i", p["front"] - p["back"], n) > 0)      # front lies on the normal's side of back
    assert np.all(np.linalg.norm(0.5 * (p["front"] + p["back"]) - mid, axis=1) < 1e-12 * np.maximum(np.abs(mid).max(axis=1), 1.0))
    assert np.all(np.einsum("ij,ij->i", n, rays.dir[h]) <= 0)          # the normal faces the incoming ray
    scale = np.maximum(np.abs(mid).max(axis=1), 1e-3)
    assert np.all(np.linalg.norm(p["front"] - p["back"], axis=1) < 1e-9 * scale)
    # No parent -> plain trace.
    none = np.zeros(len(rays), dtype=p.dtype)
    none["assembly_instance"] = 0xFFFFFFFF
    assert o.trace_parents(rays, none, threads=4).tobytes() == hits.tobytes()
