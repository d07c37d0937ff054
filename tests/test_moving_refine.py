"""Moving triangles through ShadingPoint::m_triangle_support_plane and ShadingPoint::refine_and_offset.

The reference keeps the interpolated triangle of a moving hit in a MEMBER of the leaf visitor
(m_interpolated_triangle, triangletree.h:232, assigned at triangletree.cpp:1468-1469) and makes the
ShadingPoint's support plane from it (:1483-1499); refine_and_offset then interpolates the SOURCE
vertices between the two poses around the ray time (shadingpoint.cpp:186-256).  CPU tier:

* oracle/_ref: the plane the traversal itself stored == the plane recomputed from (leaf slot, ray time),
  which is what the ABI call asgpu_get_support_planes does;
* restatement (oracle.cpp) == oracle/_ref, byte for byte: planes, parent records, child rays;
* the host build of the product code (refine_core.h, traverse_core.h::hit_triangle) == oracle/_ref.
The GPU tier (tests/test_gpu_parents.py) repeats the last comparison on the kernels.
"""
import numpy as np
import pytest

import cases
from appleseed_b200 import scenes
from hostsim import hostsim

MOVING = ["c4_msc1", "c4_msc2", "c4_msc3", "mixed"]


@pytest.fixture(scope="module")
def sim():
    return hostsim.load()


@pytest.mark.parametrize("name", MOVING + ["c3", "cornell"])
def test_traversal_plane_equals_plane_from_slot_and_time(asref, name):
    desc, rays, _ = cases.CASES[name]()
    r = asref.scene(desc)
    hits, planes = r.trace_planes(rays, threads=4)
    assert hits.tobytes() == r.trace(rays, threads=4).tobytes()
    again = r.support_planes(rays, hits, threads=4)
    assert planes.tobytes() == again.tobytes()
    h = hits["prim_type"] == 2
    assert h.sum() > 100 and np.all(planes[~h] == 0.0)
    # The plane holds the hit point: |dot(n, p - v0)| is tiny next to the triangle.
    p = rays.org[h] + hits["t"][h][:, None] * rays.dir[h]
    if name != "mixed" and name != "c3":        # one identity assembly instance: world == assembly space
        n = np.cross(planes[h][:, 3:6], planes[h][:, 6:9])
        d = np.abs(np.einsum("ij,ij->i", n, p - planes[h][:, 0:3])) / np.maximum(np.linalg.norm(n, axis=1), 1e-300)
        assert d.max() < 1e-5


@pytest.mark.parametrize("name", MOVING)
def test_restatement_matches_reference_headers_on_moving_hits(orc, asref, name):
    desc, rays, _ = cases.CASES[name]()
    o, r = orc.scene(desc), asref.scene(desc)
    hits = r.trace(rays, threads=4)
    assert o.support_planes(rays, hits, threads=4).tobytes() == r.support_planes(rays, hits, threads=4).tobytes()
    pa, pb = o.refine_offset(rays, hits, threads=4), r.refine_offset(rays, hits, threads=4)
    assert pa.tobytes() == pb.tobytes()
    h = hits["prim_type"] == 2
    moving = h & (hits["motion_segment"] >= 0)
    assert moving.sum() > 100
    # Child rays that carry their parent and their parent's time: no self-intersection.
    pts = rays.org[h] + hits["t"][h][:, None] * rays.dir[h]
    nrm = pa["geo_normal"][h] / np.linalg.norm(pa["geo_normal"][h], axis=1, keepdims=True)
    bounce = scenes.bounce_rays(pts, nrm, 5, offset=0.0)
    bounce.time_absolute, bounce.time_normalized = rays.time_absolute[h], rays.time_normalized[h]
    ha, hb = o.trace_parents(bounce, pa[h], threads=4), r.trace_parents(bounce, pa[h], threads=4)
    assert ha.tobytes() == hb.tobytes()
    if name != "mixed":         # (instances of "mixed" overlap: a bounce may legitimately start inside another one)
        assert int(((ha["prim_type"] == 2) & (ha["t"] < 1e-9)).sum()) == 0


@pytest.mark.parametrize("name", MOVING)
def test_product_host_build_matches_reference_headers(sim, asref, name):
    desc, rays, _ = cases.CASES[name]()
    r = asref.scene(desc)
    s = hostsim.SimScene(sim, desc)
    hits = r.trace(rays, threads=4)
    assert s.support_planes(rays, hits).tobytes() == r.support_planes(rays, hits, threads=4).tobytes()
    assert s.refine_offset(rays, hits).tobytes() == r.refine_offset(rays, hits, threads=4).tobytes()


def test_points_straddle_the_moving_surface(orc):
    desc, rays, _ = cases.case_c4(3)
    o = orc.scene(desc)
    hits = o.trace(rays, threads=4)
    h = hits["prim_type"] == 2
    p = o.refine_offset(rays, hits, threads=4)[h]
    planes = o.support_planes(rays, hits, threads=4)[h]
    n = np.cross(planes[:, 3:6], planes[:, 6:9])
    side = lambda q: np.einsum("ij,ij->i", n, q - planes[:, 0:3])
    assert np.all(side(p["front"]) * side(p["back"]) < 0)          # opposite sides of the interpolated triangle's plane
    g = p["geo_normal"]
    cosine = np.einsum("ij,ij->i", g, n) / (np.linalg.norm(g, axis=1) * np.linalg.norm(n, axis=1))
    assert np.all(np.abs(cosine) > 1 - 1e-4)                        # source-vertex normal == leaf-triangle normal, up to float rounding
